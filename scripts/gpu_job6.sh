#!/bin/bash
set +e
timeout 300 python tools_ldtm_bench.py > gpurun_out/ldtm_bench.log 2>&1; echo "ldtm rc=$?"; cat gpurun_out/ldtm_bench.log | cut -c1-200
LAUNCHES_PER_FORWARD=39 bash scripts/gpu_profile.sh
