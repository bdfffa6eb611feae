#!/bin/bash
set +e
timeout 900 python -m pytest tests -m gpu -x -q -k "tc_conv_layer or swizzled or fused_lateral" 2>&1 | tail -8 > gpurun_out/pytest_gpu_e1.log; cat gpurun_out/pytest_gpu_e1.log
timeout 900 python tools_conv2_sweep.py > gpurun_out/conv2_sweep_c.log 2>&1; echo "sweep rc=$?"; tail -2 gpurun_out/conv2_sweep_c.log
cp gpurun_out/conv2_sweep.json gpurun_out/conv2_sweep_c.json
timeout 900 python tools_frame_ab.py > gpurun_out/frame_ab_c.log 2>&1; echo "frame_ab rc=$?"; cp gpurun_out/frame_ab.json gpurun_out/frame_ab_c.json
timeout 1800 python -m pytest tests -m gpu -x -q -k "not tc_conv_layer and not swizzled and not fused_lateral" 2>&1 | tail -8 > gpurun_out/pytest_gpu_e2.log; cat gpurun_out/pytest_gpu_e2.log
python bench.py --steps 20 --warmup 5 > gpurun_out/bench_r2_e.json 2> gpurun_out/bench_r2_e.err; echo "bench rc=$?"; tail -c 300 gpurun_out/bench_r2_e.err
