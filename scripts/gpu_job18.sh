#!/bin/bash
set +e
timeout 1500 python -m pytest tests -m gpu -x -q -k "render_rays or end_to_end or fullsize or composite or golden or network or human or selftest" 2>&1 | tail -8 > gpurun_out/pytest_gpu_n.log; cat gpurun_out/pytest_gpu_n.log
python bench.py --steps 20 --warmup 5 > gpurun_out/bench_r2_n.json 2> gpurun_out/bench_r2_n.err; echo "bench rc=$?"; tail -c 300 gpurun_out/bench_r2_n.err
