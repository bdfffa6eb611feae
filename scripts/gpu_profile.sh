#!/bin/bash
# profiling payload (GPU box, repo root): launch list + one --set full capture of the top kernels + sanitizer runs
set +e
L=${LAUNCHES_PER_FORWARD:-42}
SKIP=$(( L * 5 ))     # launch-count probe + 3 warm-up steps + 1 step of slack (eager, one frame in flight)
# (only this library's kernels: the first forward also runs ~330 ATen launches of the one-off weight packing)
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --kernel-name-base demangled -k regex:"enerf::" -s $SKIP -c $(( L * 2 )) --csv --log-file gpurun_out/r2_launches.csv \
  python bench.py --steps 4 --warmup 3 --no-cpu-baseline --no-extras --graph 0 --inflight 1 > gpurun_out/r2_launches_bench.json 2> gpurun_out/r2_launches_bench.err
echo "launch list rc=$?"
timeout 1500 ncu --set full --clock-control none --import-source on --kernel-name-base demangled \
  -k regex:"enerf::" -s $SKIP -c $L -o gpurun_out/r2_prof \
  python bench.py --steps 4 --warmup 3 --no-cpu-baseline --no-extras --graph 0 --inflight 1 > gpurun_out/r2_prof_bench.json 2> gpurun_out/r2_prof_bench.err
echo "full capture rc=$?"
ENERF_B200_OVERLAP=0 timeout 500 compute-sanitizer --tool memcheck --print-limit 20 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2_sanitizer_memcheck.log 2>&1
echo "memcheck rc=$?"; tail -4 gpurun_out/r2_sanitizer_memcheck.log
ENERF_B200_OVERLAP=0 timeout 500 compute-sanitizer --tool racecheck --print-limit 20 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2_sanitizer_racecheck.log 2>&1
echo "racecheck rc=$?"; tail -4 gpurun_out/r2_sanitizer_racecheck.log
