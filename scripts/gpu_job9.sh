#!/bin/bash
set +e
timeout 900 python -m pytest tests -m gpu -x -q -k "tc_conv_layer or swizzled or fused_lateral or cost_reg or feature_net" 2>&1 | tail -8 > gpurun_out/pytest_gpu_g1.log; cat gpurun_out/pytest_gpu_g1.log
timeout 900 python tools_conv2_sweep.py quick > gpurun_out/conv2_sweep_d.log 2>&1; echo "sweep rc=$?"; tail -2 gpurun_out/conv2_sweep_d.log
cp gpurun_out/conv2_sweep.json gpurun_out/conv2_sweep_d.json
timeout 600 python tools_conv2_timeline.py > gpurun_out/conv2_timeline_b.log 2>&1; echo "timeline rc=$?"; cp gpurun_out/conv2_timeline.json gpurun_out/conv2_timeline_b.json
timeout 900 python tools_frame_ab.py > gpurun_out/frame_ab_e.log 2>&1; echo "frame_ab rc=$?"; cp gpurun_out/frame_ab.json gpurun_out/frame_ab_e.json; cut -c1-400 gpurun_out/frame_ab_e.log
python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-extras > gpurun_out/bench_r2_g.json 2> gpurun_out/bench_r2_g.err; echo "bench rc=$?"; tail -c 300 gpurun_out/bench_r2_g.err
