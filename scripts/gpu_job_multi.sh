#!/bin/bash
# multi-GPU payload: N = $1 ranks (default 2): frame-parallel throughput + the band layout's intra-frame latency in one line
N=${1:-2}
set +e
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 20 --warmup 5 \
  > gpurun_out/bench_r2_${N}gpu.json 2> gpurun_out/bench_r2_${N}gpu.err
echo "bench ${N} gpu rc=$?"; tail -c 400 gpurun_out/bench_r2_${N}gpu.err; head -c 600 gpurun_out/bench_r2_${N}gpu.json
