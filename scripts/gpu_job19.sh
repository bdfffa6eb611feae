#!/bin/bash
set +e
LAUNCHES_PER_FORWARD=39 bash scripts/gpu_profile.sh
timeout 300 python tools_ray_timeline.py > gpurun_out/ray_timeline_b.log 2>&1; echo "ray timeline rc=$?"; tail -30 gpurun_out/ray_timeline_b.log | cut -c1-250
