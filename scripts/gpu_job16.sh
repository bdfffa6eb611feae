#!/bin/bash
set +e
timeout 1500 python -m pytest tests -m gpu -x -q -k "not tc_conv_layer" 2>&1 | tail -8 > gpurun_out/pytest_gpu_l.log; cat gpurun_out/pytest_gpu_l.log
python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/bench_r2_l.json 2> gpurun_out/bench_r2_l.err; echo "bench rc=$?"; tail -c 300 gpurun_out/bench_r2_l.err
