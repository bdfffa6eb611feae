#!/bin/bash
set +e
timeout 600 python tools_fused_lat_timeline.py > gpurun_out/fused_lat_timeline.log 2>&1; echo "rc=$?"; cut -c1-300 gpurun_out/fused_lat_timeline.log
