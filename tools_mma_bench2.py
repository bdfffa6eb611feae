"""tcgen05.mma issue-rate probe v2 (enerf_tc_mma_bench2): issuers per CTA x CTAs per SM.  cycles per MMA per SM =
max issuer time * clk / (n_mma * issuers * CTAs per SM)."""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
from enerf_b200 import capi  # noqa: E402

CLK = 1.965
rows = []
for layout in (0, 2, 4, 6):
    for N in (16, 32, 64):
        for issuers in (1, 2, 4):
            for ctas_per_sm, pad in ((1, 120 * 1024), (2, 0)):
                grid = 148 * ctas_per_sm
                capi.tc_mma_bench2(layout, N, 64, issuers, 1, grid, pad)
                a = capi.tc_mma_bench2(layout, N, 512, issuers, 1, grid, pad)[:, :issuers].max().item()
                b = capi.tc_mma_bench2(layout, N, 2560, issuers, 1, grid, pad)[:, :issuers].max().item()
                per_issuer = (b - a) / 2048 * CLK
                rec = {"layout": layout, "N": N, "issuers": issuers, "ctas_per_sm": ctas_per_sm, "cycles_per_mma_per_issuer": per_issuer,
                       "cycles_per_mma_per_sm": per_issuer / (issuers * ctas_per_sm)}
                rows.append(rec)
                print(rec, flush=True)
os.makedirs("gpurun_out", exist_ok=True)
json.dump(rows, open("gpurun_out/mma_bench2.json", "w"), indent=1)
