"""TMEM read latency with the tensor pipe idle / busy (enerf_tc_ldtm_bench): ns and cycles per {tcgen05.ld + wait::ld}."""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
from enerf_b200 import capi  # noqa: E402

rows = []
for cols in (8, 32):
    for mode in (0, 1, 2):
        for N in (16, 64):
            n_ld = 2000
            n_mma = 12000        # keeps the pipe busy for longer than the loads take
            capi.tc_ldtm_bench(mode, N, 100, 100, cols)
            t = capi.tc_ldtm_bench(mode, N, n_mma, n_ld, cols).double()
            ns = float(t.mean()) / n_ld
            rec = {"cols": cols, "mma_issuers": mode, "N": N, "ns_per_load": ns, "cycles_per_load": ns * 1.965}
            rows.append(rec)
            print(rec, flush=True)
os.makedirs("gpurun_out", exist_ok=True)
json.dump(rows, open("gpurun_out/ldtm_bench.json", "w"), indent=1)
