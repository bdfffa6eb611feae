#!/bin/bash
set +e
timeout 2400 python -m pytest tests -m gpu -x -q -k "not tc_conv_layer" 2>&1 | tail -8 > gpurun_out/pytest_gpu_p.log; cat gpurun_out/pytest_gpu_p.log
python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-extras > gpurun_out/bench_r2_p.json 2> gpurun_out/bench_r2_p.err; echo "bench rc=$?"; tail -c 300 gpurun_out/bench_r2_p.err
