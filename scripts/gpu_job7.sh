#!/bin/bash
set +e
timeout 600 python tools_conv2_timeline.py > gpurun_out/conv2_timeline.log 2>&1; echo "timeline rc=$?"; tail -40 gpurun_out/conv2_timeline.log | cut -c1-220
timeout 2400 python -m pytest tests -m gpu -x -q 2>&1 | tail -8 > gpurun_out/pytest_gpu_f.log; cat gpurun_out/pytest_gpu_f.log
python bench.py --steps 20 --warmup 5 > gpurun_out/bench_r2_f.json 2> gpurun_out/bench_r2_f.err; echo "bench rc=$?"; tail -c 300 gpurun_out/bench_r2_f.err
python bench.py --workload c5 --steps 10 --warmup 3 > gpurun_out/bench_r2_f_c5.json 2> gpurun_out/bench_r2_f_c5.err; echo "bench c5 rc=$?"
python bench.py --workload c4 --steps 20 --warmup 5 > gpurun_out/bench_r2_f_c4.json 2> gpurun_out/bench_r2_f_c4.err; echo "bench c4 rc=$?"
