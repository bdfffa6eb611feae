#!/bin/bash
set +e
timeout 1200 python -m pytest tests -m gpu -x -q -k "tc_conv_layer or swizzled or fused_lateral or cost_reg or feature_net" 2>&1 | tail -8 > gpurun_out/pytest_gpu_j1.log; cat gpurun_out/pytest_gpu_j1.log
timeout 600 python tools_fused_lat_timeline.py > gpurun_out/fused_lat_timeline_c.log 2>&1; echo "rc=$?"; cut -c1-300 gpurun_out/fused_lat_timeline_c.log | head -14; grep "==" gpurun_out/fused_lat_timeline_c.log
timeout 900 python tools_frame_ab.py > gpurun_out/frame_ab_h.log 2>&1; echo "frame_ab rc=$?"; cp gpurun_out/frame_ab.json gpurun_out/frame_ab_h.json; cut -c1-330 gpurun_out/frame_ab_h.log
python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-extras > gpurun_out/bench_r2_j.json 2> gpurun_out/bench_r2_j.err; echo "bench rc=$?"; tail -c 300 gpurun_out/bench_r2_j.err
