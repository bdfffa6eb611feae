"""Runs the tc_conv2 groundwork probes on the B200 and writes gpurun_out/tc_probe.json:

  swz  : enerf_tc_swz_selftest over Kf (swizzle span) x row offsets x base_offset modes.  B is a one-hot matrix, so
         D[m][n] = the element the tensor core read as (row m, k n): on a mismatch the report decodes WHICH
         (row, k) of A arrived there (A[r][k] = r + k/64 is unique per element).
  tma  : enerf_tma_box_bench -- aggregate GB/s and B/clk/SM of halo-box loads for C = 8/16/32 channels.
"""
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
from enerf_b200 import capi  # noqa: E402


def swz_report():
    out = []
    for Kf in (8, 16, 32):
        rows = 256
        # TF32-exact values (11 significant bits): (row % 32) * 32 + k  <= 1023
        A = ((torch.arange(rows, dtype=torch.float32)[:, None] % 32) * 32 + torch.arange(Kf, dtype=torch.float32)[None, :]).cuda().contiguous()
        N = max(16, Kf)
        B = torch.zeros(N, Kf)
        for k in range(Kf):
            B[k, k] = 1.0
        B = B.cuda()
        for row_off in (0, 1, 3, 8, 13, 34, 70, 105):
            for bo in (0, 1):
                D = torch.full((128, N), float("nan"), device="cuda")
                try:
                    capi.tc_swz_selftest(A, B, D, row_off, bo)
                    torch.cuda.synchronize()
                except Exception as e:  # noqa: BLE001
                    out.append({"Kf": Kf, "row_off": row_off, "bo_mode": bo, "error": str(e)})
                    continue
                got = D[:, :Kf].cpu()
                ref = A[row_off:row_off + 128].cpu()
                ok = bool(torch.equal(got, ref))
                rec = {"Kf": Kf, "row_off": row_off, "bo_mode": bo, "ok": ok}
                if not ok:
                    r = torch.floor(got / 32.0)
                    k = got - 32.0 * r
                    bad = (got != ref)
                    rec["n_bad"] = int(bad.sum())
                    # decoded source (row - expected row, k) for the first rows
                    rec["decoded_first_rows"] = [[(int(r[m, n] - ((row_off + m) % 32)) if got[m, n] == got[m, n] else None, int(k[m, n]) if got[m, n] == got[m, n] else None)
                                                  for n in range(0, Kf, 4)] for m in range(0, 16)]
                out.append(rec)
    return out


def tma_report():
    out = []
    sink = torch.zeros(1, device="cuda")
    clk = 1.965e9
    for C in (8, 16, 32):
        for name, shape, tile in (("2d_7x32", (3, 512, 640, C), (32, 7, 1)), ("3d_4x4x32", (8, 256, 320, C), (32, 4, 4))):
            x = torch.randn(shape, device="cuda")
            tx, ty, tz = tile
            box = (tx + 2) * (ty + 2) * ((tz + 2) if tz > 1 else 1) * C * 4
            for depth in (1, 2, 4):
                if depth * ((box + 1023) // 1024 * 1024) + 1024 > 220 * 1024:
                    continue
                for grid in (148, 296):
                    iters = 128
                    capi.tma_box_bench(x, tx, ty, tz, depth, 8, grid, sink)
                    torch.cuda.synchronize()
                    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                    e0.record()
                    capi.tma_box_bench(x, tx, ty, tz, depth, iters, grid, sink)
                    e1.record()
                    torch.cuda.synchronize()
                    ms = e0.elapsed_time(e1)
                    total = box * iters * grid
                    out.append({"C": C, "case": name, "box_bytes": box, "depth": depth, "grid": grid, "ms": ms, "GBps": total / ms / 1e6,
                                "B_per_clk_per_SM": total / (ms * 1e-3) / clk / 148, "us_per_box_per_cta": ms * 1e3 / iters})
    return out


if __name__ == "__main__":
    rep = {"swz": swz_report(), "tma": tma_report()}
    os.makedirs("gpurun_out", exist_ok=True)
    with open("gpurun_out/tc_probe.json", "w") as f:
        json.dump(rep, f, indent=1)
    for r in rep["swz"]:
        print({k: v for k, v in r.items() if k != "decoded_first_rows"})
        if "decoded_first_rows" in r:
            print("   decoded:", r["decoded_first_rows"][:10])
    for r in rep["tma"]:
        print(r)
