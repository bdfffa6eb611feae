"""Diagnostic: phase timeline (ns) of one CTA of the tcgen05 conv kernel for a few layer shapes."""
import sys, torch
sys.path.insert(0, "/root/repo")
from enerf_b200 import capi, packing

def run(name, KD, KH, cin, cout, dims, mode=0):
    D, H, W = dims
    x = torch.randn(D, H, W, cin, device="cuda")
    w = torch.randn(cout, cin, KD, KH, KH) / (cin * KD * KH * KH) ** 0.5
    wp = packing.pack_tc_conv(packing._taps_cin_cout(w), fold_kx=packing.tc_fold_kx(KD, KH, 1, cout)).cuda()
    b = torch.zeros(cout, device="cuda")
    out = torch.empty(D, H, W, cout, device="cuda")
    buf = torch.zeros(64, dtype=torch.int64, device="cuda")
    for it in range(3):
        capi.tc_conv_debug(buf if it == 2 else None)
        capi.tc_conv(0, KD, KH, cout, mode, 1, x, wp, b, None, out)
        torch.cuda.synchronize()
    capi.tc_conv_debug(None)
    t = buf.cpu().tolist()
    t0 = t[0]
    ev = {i: t[i] - t0 for i in range(64) if t[i]}
    print(name, "stages", cin // 8, {k: v for k, v in sorted(ev.items())})

run("toplayer 1x1 32->32 (3,128,160)", 1, 1, 32, 32, (3, 128, 160))
run("conv2.1 3x3 32->32 (3,128,160)", 1, 3, 32, 32, (3, 128, 160))
run("conv0-L1 3x3x3 16->8 (8,256,320)", 3, 3, 16, 8, (8, 256, 320))
run("head-like 3x3x3 8->8 (8,256,320)", 3, 3, 8, 8, (8, 256, 320))
run("conv4 3x3x3 32->32 (2,64,80)", 3, 3, 32, 32, (2, 64, 80))
