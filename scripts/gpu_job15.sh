#!/bin/bash
set +e
L=39; SKIP=$(( L * 5 ))
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --kernel-name-base demangled -k regex:"enerf::" -s $SKIP -c $(( L * 2 )) --csv --log-file gpurun_out/r2_launches_b.csv \
  python bench.py --steps 4 --warmup 3 --no-cpu-baseline --no-extras --graph 0 --inflight 1 > gpurun_out/r2_launches_b_bench.json 2> gpurun_out/r2_launches_b_bench.err
echo "launch list rc=$?"
timeout 2400 python -m pytest tests -m gpu -x -q 2>&1 | tail -8 > gpurun_out/pytest_gpu_k.log; cat gpurun_out/pytest_gpu_k.log
