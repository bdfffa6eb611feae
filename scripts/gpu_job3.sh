#!/bin/bash
set +e
timeout 900 python -m pytest tests -m gpu -x -q -k "tc_conv_layer or swizzled or fused_lateral" 2>&1 | tail -8 > gpurun_out/pytest_gpu_d1.log; cat gpurun_out/pytest_gpu_d1.log
timeout 900 python tools_conv2_sweep.py > gpurun_out/conv2_sweep_b.log 2>&1; echo "sweep rc=$?"; tail -2 gpurun_out/conv2_sweep_b.log
cp gpurun_out/conv2_sweep.json gpurun_out/conv2_sweep_b.json
timeout 600 python tools_ray_timeline.py > gpurun_out/ray_timeline.log 2>&1; echo "timeline rc=$?"; tail -3 gpurun_out/ray_timeline.log | cut -c1-300
timeout 900 python tools_frame_ab.py > gpurun_out/frame_ab_b.log 2>&1; echo "frame_ab rc=$?"; cp gpurun_out/frame_ab.json gpurun_out/frame_ab_b.json
timeout 1500 python -m pytest tests -m gpu -x -q -k "not tc_conv_layer and not swizzled and not fused_lateral and not c4_1024 and not c5_1920 and not headline_512" 2>&1 | tail -8 > gpurun_out/pytest_gpu_d2.log; cat gpurun_out/pytest_gpu_d2.log
