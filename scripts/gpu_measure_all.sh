#!/bin/bash
# round-2 final measurements on one B200: full GPU test suite, bench at the three workloads, frames-in-flight sweep
set +e
timeout 2400 python -m pytest tests -m gpu -x -q 2>&1 | tail -8 > gpurun_out/pytest_gpu_final.log; cat gpurun_out/pytest_gpu_final.log
python bench.py --steps 20 --warmup 5 > gpurun_out/bench_final_c2.json 2> gpurun_out/bench_final_c2.err; echo "bench c2 rc=$?"; tail -c 200 gpurun_out/bench_final_c2.err
for n in 2 6 8; do
  python bench.py --steps 20 --warmup 5 --inflight $n --no-cpu-baseline --no-extras > gpurun_out/bench_final_c2_if$n.json 2> gpurun_out/bench_final_c2_if$n.err; echo "bench inflight $n rc=$?"
done
python bench.py --workload c4 --steps 20 --warmup 5 > gpurun_out/bench_final_c4.json 2> gpurun_out/bench_final_c4.err; echo "bench c4 rc=$?"
python bench.py --workload c5 --steps 10 --warmup 3 > gpurun_out/bench_final_c5.json 2> gpurun_out/bench_final_c5.err; echo "bench c5 rc=$?"
python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_final_ref.json 2> gpurun_out/bench_final_ref.err; echo "bench ref rc=$?"
