#!/bin/bash
set +e
timeout 600 python -m pytest tests -m gpu -x -q -k "fused_lateral or tc_conv_layer" 2>&1 | tail -8 > gpurun_out/pytest_gpu_f1.log; cat gpurun_out/pytest_gpu_f1.log
timeout 900 python tools_frame_ab.py > gpurun_out/frame_ab_d.log 2>&1; echo "frame_ab rc=$?"; cp gpurun_out/frame_ab.json gpurun_out/frame_ab_d.json; grep -o '"config": "[^"]*", "single_fps": [0-9.]*, "inflight4_fps": [0-9.]*' gpurun_out/frame_ab_d.log
LAUNCHES_PER_FORWARD=40 bash scripts/gpu_profile.sh
