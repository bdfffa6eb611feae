"""Diagnostic: in-kernel %globaltimer timeline of the tensor-core ray kernels at the bench workload (CTA 0, first tiles).
  warp-specialised kernel (render_rays_ws.cu): role 0 = consumer, 1 / 2 = gather groups; stamps per tile, ns relative to the kernel's first stamp
  single-role kernel (render_rays_tc.cuh): two tiles of 14 phase stamps.
Writes gpurun_out/ray_timeline.json."""
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import bench  # noqa: E402
from enerf_b200 import capi  # noqa: E402

cfg, net, batch, wl = bench.build_problem("c2")
net = net.cuda()
for k in [k for k in batch if k.startswith("rays_")]:
    batch.pop(k)
gb = bench.to_dev(batch, torch.device("cuda"))
out = {}
for impl, name in ((2, "ws"), (1, "single_role")):
    capi.render_rays_tc_select(impl)
    buf = torch.zeros(3 * 8 * 16, dtype=torch.int64, device="cuda")
    with torch.no_grad():
        for it in range(4):
            capi.render_rays_debug(buf if it == 3 else None)
            net(gb)
            torch.cuda.synchronize()
    capi.render_rays_debug(None)
    t = buf.cpu().view(3, 8, 16)
    if impl == 2:
        t0 = int(t[t > 0].min())
        cons = ["tile start", "gather ready", "batch1 issued", "batch1 done", "E1 done", "G2 done", "E2 done", "G3 done", "E3 done", "G5 done", "E5 done", "tile end"]
        gath = ["start", "gathered (regs)", "buffer free", "rows stored"]
        rep = {"consumer": [{cons[i]: int(t[0, k, i]) - t0 for i in range(12)} for k in range(8)],
               "gather0": [{gath[i]: int(t[1, k, i]) - t0 for i in range(4)} for k in range(4)],
               "gather1": [{gath[i]: int(t[2, k, i]) - t0 for i in range(4)} for k in range(4)]}
        for k, v in rep.items():
            for row in v:
                print(name, k, row)
    else:
        flat = buf.cpu()[:32].tolist()
        names = ["start", "gather+A", "sync1", "G1 done", "E1 done", "G2 done", "E2 done", "G3 done", "E3 done", "sync5", "G5 done", "E5 done", "composite", "tile end"]
        rep = [{names[i]: flat[tile * 16 + i] - flat[tile * 16] for i in range(14)} for tile in range(2)]
        for row in rep:
            print(name, row)
    out[name] = rep
capi.render_rays_tc_select(0)
os.makedirs("gpurun_out", exist_ok=True)
json.dump(out, open("gpurun_out/ray_timeline.json", "w"), indent=1)
