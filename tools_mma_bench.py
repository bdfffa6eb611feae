import sys
sys.path.insert(0, "/root/repo")
from enerf_b200 import capi
for layout in (0, 2, 6):
    for N in (16, 64, 256):
        for accs in (1, 2):
            if accs * N > 512: continue
            capi.tc_mma_bench(layout, N, 64, accs)
            a = capi.tc_mma_bench(layout, N, 256, accs)
            b = capi.tc_mma_bench(layout, N, 2304, accs)
            print(f"layout {layout} N {N:3d} accs {accs}: {(b - a) / 2048 * 1.965:7.1f} cycles/MMA  (256: {a} ns, 2304: {b} ns)")
