"""In-kernel timeline of the fused lat0 + smooth0 launch of FeatureNet (csrc/tc_conv2.cu, PROD > 0) at the headline size
(3 x 512 x 640), for 4 / 6 / 8 computing producer warps and folded / unfolded taps, plus the CUDA-event time of the whole
FeatureNet call.  Writes gpurun_out/fused_lat_timeline.json."""
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "tests"))
from enerf_b200 import capi, packing  # noqa: E402
from enerf_b200 import config as bcfg  # noqa: E402
import stage_harness  # noqa: E402

S, H, W = 3, 512, 640
cfg = bcfg.make_cfg(volume_planes=[8, 8], render_if=[False, True])
sd, _ = stage_harness.make_case(64, 96, 2, cfg, seed=5)
src = (2 * torch.rand(S, 3, H, W) - 1).cuda()
ws = torch.empty(capi.feature_net_workspace_bytes(S, H, W) // 4, device="cuda")
f0 = torch.empty(S, H // 4, W // 4, 32, device="cuda")
f1 = torch.empty(S, H // 2, W // 2, 16, device="cuda")
f2 = torch.empty(S, H, W, 8, device="cuda")
out = {}
for rule in (2, 1):
    capi.tc_conv_fold_rule(rule)
    pk = packing.pack_feature_net(sd, torch.device("cuda"), tensor_cores=True)
    for fuse, ty in ((8, 0), (6, 0), (6, 11), (4, 0), (False, 0)):
        if ty and rule != 2:
            continue
        capi.tc_conv2_fuse_lateral(fuse)
        capi.tc_conv2_tune(ty=ty)          # (forces the tile height of EVERY tc_conv2 layer: compare the fused launch's stamps, not the total)
        buf = torch.zeros(3 * 16 * 8, dtype=torch.int64, device="cuda")
        for it in range(4):
            capi.tc_conv2_debug_lateral(buf if it == 3 else None)
            capi.feature_net(pk, src, f0, f1, f2, ws, tensor_cores=True)
            torch.cuda.synchronize()
        capi.tc_conv2_debug_lateral(None)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(20):
            capi.feature_net(pk, src, f0, f1, f2, ws, tensor_cores=True)
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 20
        key = f"rule{rule}.{'prod%d' % fuse if fuse else 'separate'}" + (f".ty{ty}" if ty else "")
        rep = {"feature_net_ms": ms}
        if fuse:
            t = buf.cpu().view(3, 16, 8)
            t0 = int(t[t > 0].min())
            rep.update({"producer": [[int(v) - t0 if v else None for v in t[0, k, :4]] for k in range(8)],
                        "mma": [[int(v) - t0 if v else None for v in t[1, k, :4]] for k in range(8)],
                        "epilogue": [[int(v) - t0 if v else None for v in t[2, k, :4]] for k in range(8)]})
        out[key] = rep
        print("==", key, "feature_net %.4f ms" % ms)
        if fuse:
            print("   producer [start, operand slot free, sources landed, tile written]", rep["producer"][:6])
            print("   mma      [start, acc free, tile landed, issued]                  ", rep["mma"][:6])
            print("   epilogue [start, acc full, acc released, stored]                 ", rep["epilogue"][:6])
capi.tc_conv_fold_rule(2)
capi.tc_conv2_fuse_lateral(True)
capi.tc_conv2_tune()
os.makedirs("gpurun_out", exist_ok=True)
json.dump(out, open("gpurun_out/fused_lat_timeline.json", "w"), indent=1)
