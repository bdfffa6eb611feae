#!/bin/bash
set +e
timeout 900 python tools_frame_ab.py > gpurun_out/frame_ab.log 2>&1; echo "frame_ab rc=$?"; tail -3 gpurun_out/frame_ab.log | cut -c1-400
timeout 900 python tools_conv2_sweep.py > gpurun_out/conv2_sweep_full.log 2>&1; echo "sweep rc=$?"; tail -2 gpurun_out/conv2_sweep_full.log
cp gpurun_out/conv2_sweep.json gpurun_out/conv2_sweep_full.json
LAUNCHES_PER_FORWARD=40 bash scripts/gpu_profile.sh
