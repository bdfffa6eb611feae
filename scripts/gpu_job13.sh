#!/bin/bash
set +e
timeout 1200 python -m pytest tests -m gpu -x -q -k "tc_conv_layer or swizzled or fused_lateral or cost_reg or feature_net or render_rays or end_to_end or selftest" 2>&1 | tail -8 > gpurun_out/pytest_gpu_i1.log; cat gpurun_out/pytest_gpu_i1.log
timeout 900 python tools_frame_ab.py > gpurun_out/frame_ab_g.log 2>&1; echo "frame_ab rc=$?"; cp gpurun_out/frame_ab.json gpurun_out/frame_ab_g.json; cut -c1-330 gpurun_out/frame_ab_g.log
timeout 600 python tools_fused_lat_timeline.py > gpurun_out/fused_lat_timeline_b.log 2>&1; echo "rc=$?"; cut -c1-300 gpurun_out/fused_lat_timeline_b.log | head -12
timeout 900 python tools_conv2_sweep.py quick > gpurun_out/conv2_sweep_e.log 2>&1; echo "sweep rc=$?"; tail -2 gpurun_out/conv2_sweep_e.log
cp gpurun_out/conv2_sweep.json gpurun_out/conv2_sweep_e.json
python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-extras > gpurun_out/bench_r2_i.json 2> gpurun_out/bench_r2_i.err; echo "bench rc=$?"; tail -c 300 gpurun_out/bench_r2_i.err
