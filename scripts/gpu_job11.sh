#!/bin/bash
set +e
timeout 900 python -m pytest tests -m gpu -x -q -k "tc_conv_layer or swizzled or fused_lateral or cost_reg or feature_net" 2>&1 | tail -8 > gpurun_out/pytest_gpu_h1.log; cat gpurun_out/pytest_gpu_h1.log
timeout 900 python tools_frame_ab.py > gpurun_out/frame_ab_f.log 2>&1; echo "frame_ab rc=$?"; cp gpurun_out/frame_ab.json gpurun_out/frame_ab_f.json; cut -c1-330 gpurun_out/frame_ab_f.log
python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-extras > gpurun_out/bench_r2_h.json 2> gpurun_out/bench_r2_h.err; echo "bench rc=$?"; tail -c 300 gpurun_out/bench_r2_h.err
