#!/bin/bash
set +e
timeout 1500 python -m pytest tests -m gpu -x -q -k "cost_volume or end_to_end or golden or composite or band" 2>&1 | tail -8 > gpurun_out/pytest_gpu_o.log; cat gpurun_out/pytest_gpu_o.log
python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-extras > gpurun_out/bench_r2_o.json 2> gpurun_out/bench_r2_o.err; echo "bench rc=$?"; tail -c 300 gpurun_out/bench_r2_o.err
