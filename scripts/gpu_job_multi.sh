#!/bin/bash
# multi-GPU payload: N = $1 ranks (default 2): frame-parallel throughput + the band layout's intra-frame latency in one line
N=${1:-2}
set +e
python - > gpurun_out/intra_world1.log 2>&1 <<'PY'
import torch, bench
from enerf_b200 import dist as edist
cfg, net, batch, wl = bench.build_problem("c2")
net = net.cuda()
gb = bench.to_dev({k: v for k, v in batch.items() if not k.startswith("rays_")}, torch.device("cuda"))
print(edist.measure_intra_frame(net, gb, 0, 1, torch.device("cuda"), steps=10, warmup=3))
PY
echo "world-1 intra check rc=$?"; tail -c 600 gpurun_out/intra_world1.log
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 20 --warmup 5 \
  > gpurun_out/bench_r2_${N}gpu.json 2> gpurun_out/bench_r2_${N}gpu.err
echo "bench ${N} gpu rc=$?"; tail -c 600 gpurun_out/bench_r2_${N}gpu.err; python -c "
import json; d=json.load(open('gpurun_out/bench_r2_${N}gpu.json')); print(d['value'], d['e2e']['value'], d['config'].get('intra_frame'))"
