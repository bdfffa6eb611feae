"""In-kernel timeline of csrc/tc_conv2.cu (enerf_tc_conv2_debug) for a few headline layers: per tile of CTA 0, when the producer /
MMA warp 0 / epilogue reached their stamps (ns since the first).  Writes gpurun_out/conv2_timeline.json."""
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
from enerf_b200 import capi, packing  # noqa: E402
from tools_conv2_sweep import LAYERS  # noqa: E402

WANT = ("feat.conv0.1", "feat.smooth0", "reg0.conv0", "reg1.conv0", "reg1.conv2", "reg1.head9", "reg1.conv11")
RULE0_TOO = ("feat.conv0.1", "feat.smooth0", "reg1.head9")          # layers the level-1/2 fold rules add: also unfolded (rule 0)
out = {}
for name, kind, KD, KH, cin, cout, mode, relu, (D, H, W), rule in [(*L, f) for L in LAYERS for f in (2, 0)]:
    if name not in WANT or (rule == 0 and name not in RULE0_TOO):
        continue
    capi.tc_conv_fold_rule(rule)
    fold = kind == 0 and packing.tc_fold_kx(KD, KH, 1, cout, single=(mode == 3), head=(mode == 1), cin=cin)
    g = torch.Generator().manual_seed(1)
    x = torch.randn(D, H, W, cin, generator=g).cuda()
    skip = None
    if kind == 0:
        w = torch.randn(cout, cin, KD, KH, KH, generator=g) / (cin * KD * KH * KH) ** 0.5
        wp = packing.pack_tc_conv(packing._taps_cin_cout(w), fold_kx=fold).cuda()
        Do, Ho, Wo = D, H, W
    else:
        w = torch.randn(cin, cout, 3, 3, 3, generator=g) / (cin * 27 / 8) ** 0.5
        wp = packing.pack_tc_deconv(w.permute(2, 3, 4, 0, 1).reshape(27, cin, cout)).cuda()
        Do, Ho, Wo = 2 * D, 2 * H, 2 * W
        skip = torch.randn(Do, Ho, Wo, cout, generator=g).cuda()
    bias = torch.zeros(cout).cuda() if mode in (0, 2) else None
    o = torch.empty((Do, Ho, Wo) if mode == 3 else (Do, Ho, Wo, 8 if mode == 1 else cout)).cuda()
    o2 = torch.empty(Do, Ho, Wo).cuda() if mode == 1 else None
    buf = torch.zeros(3 * 16 * 8, dtype=torch.int64, device="cuda")
    for it in range(4):
        capi.tc_conv2_debug(buf if it == 3 else None)
        capi.tc_conv(kind, KD, KH, cout, mode, relu, x, wp, bias, skip, o, o2, out_cstride=(8 if mode == 1 else cout))
        torch.cuda.synchronize()
    capi.tc_conv2_debug(None)
    capi.tc_conv_fold_rule(2)
    t = buf.cpu().view(3, 16, 8)
    t0 = int(t[t > 0].min())
    plan = capi.tc_conv2_plan(kind, KD, KH, 1, cin, cout, mode, D, H, W, fold)
    rep = {"plan": {k: plan[k] for k in ("TZ", "TY", "n_mt", "kbc", "n_kb", "n_slots", "n_acc", "n_tiles", "N", "n_taps")},
           "producer": [[int(v) - t0 if v else None for v in t[0, k, :3]] for k in range(8)],
           "mma": [[int(v) - t0 if v else None for v in t[1, k, :4]] for k in range(8)],
           "epilogue": [[int(v) - t0 if v else None for v in t[2, k, :4]] for k in range(8)]}
    name = name + (".rule0" if rule == 0 else "")
    out[name] = rep
    print("==", name, rep["plan"])
    print("   producer [start, slot free, boxes issued]      ", rep["producer"][:6])
    print("   mma      [start, acc free, kblock landed, issued]", rep["mma"][:6])
    print("   epilogue [start, acc full, acc released, stored] ", rep["epilogue"][:6])
os.makedirs("gpurun_out", exist_ok=True)
json.dump(out, open("gpurun_out/conv2_timeline.json", "w"), indent=1)
