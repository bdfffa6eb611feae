"""Diagnostic: phase timeline of the tcgen05 ray kernel (first two tiles of CTA 0) at the bench workload."""
import sys, torch
sys.path.insert(0, "/root/repo")
import bench
from enerf_b200 import capi, config as cfg_mod, synthetic
cfg, net, batch = bench.build_problem(cfg_mod, synthetic)
net = net.cuda()
gb = {k: v.cuda() for k, v in batch.items()}
buf = torch.zeros(32, dtype=torch.int64, device="cuda")
with torch.no_grad():
    for it in range(4):
        capi.render_rays_debug(buf if it == 3 else None)
        net(gb)
        torch.cuda.synchronize()
capi.render_rays_debug(None)
t = buf.cpu().tolist()
names = ["start", "gather+A", "sync1", "G1 done", "E1 done", "G2 done", "E2 done", "G3 done", "E3 done", "sync5", "G5 done", "E5 done", "composite", "tile end"]
for tile in range(2):
    base = t[tile * 16]
    print("tile", tile, {names[i]: t[tile * 16 + i] - base for i in range(14)})
