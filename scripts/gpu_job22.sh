#!/bin/bash
set +e
timeout 1500 python -m pytest tests -m gpu -x -q -k "render_rays or end_to_end or fullsize or composite or golden or network or human" 2>&1 | tail -8 > gpurun_out/pytest_gpu_q.log; cat gpurun_out/pytest_gpu_q.log
python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-extras > gpurun_out/bench_r2_q.json 2> gpurun_out/bench_r2_q.err; echo "bench rc=$?"; tail -c 300 gpurun_out/bench_r2_q.err
timeout 300 python tools_ray_timeline.py > gpurun_out/ray_timeline_c.log 2>&1; echo "ray timeline rc=$?"; grep single_role gpurun_out/ray_timeline_c.log | cut -c1-330
