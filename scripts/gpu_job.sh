#!/bin/bash
# one gpurun payload (runs on the GPU box from the repo root); logs under gpurun_out/
set +e
python tools_mma_bench2.py > gpurun_out/mma_bench2.log 2>&1; echo "mma rc=$?"
timeout 900 python -m pytest tests -m gpu -x -q -k "tc_conv_layer or swizzled or fused_lateral" 2>&1 | tail -15 > gpurun_out/pytest_gpu_c1.log; cat gpurun_out/pytest_gpu_c1.log
timeout 600 python tools_conv2_sweep.py quick > gpurun_out/conv2_sweep_quick.log 2>&1; tail -3 gpurun_out/conv2_sweep_quick.log
timeout 1500 python -m pytest tests -m gpu -x -q -k "not tc_conv_layer and not swizzled and not fused_lateral and not c4_1024 and not headline_512" 2>&1 | tail -15 > gpurun_out/pytest_gpu_b.log; cat gpurun_out/pytest_gpu_b.log
python bench.py --steps 20 --warmup 5 > gpurun_out/bench_r2_b.json 2> gpurun_out/bench_r2_b.err; echo "bench rc=$?"; tail -c 300 gpurun_out/bench_r2_b.err
python bench.py --workload c4 --steps 10 --warmup 3 > gpurun_out/bench_r2_c4.json 2> gpurun_out/bench_r2_c4.err; echo "c4 rc=$?"; tail -c 300 gpurun_out/bench_r2_c4.err
python bench.py --workload c5 --steps 10 --warmup 3 > gpurun_out/bench_r2_c5.json 2> gpurun_out/bench_r2_c5.err; echo "c5 rc=$?"; tail -c 300 gpurun_out/bench_r2_c5.err
