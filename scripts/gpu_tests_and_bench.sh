#!/bin/bash
set +e
timeout 2400 python -m pytest tests -m gpu -x -q 2>&1 | tail -8 > gpurun_out/pytest_gpu_final.log; cat gpurun_out/pytest_gpu_final.log
python bench.py --steps 20 --warmup 5 > gpurun_out/bench_final_c2.json 2> gpurun_out/bench_final_c2.err; echo "bench c2 rc=$?"; tail -c 200 gpurun_out/bench_final_c2.err
