#!/bin/bash
set -x
mkdir -p gpurun_out
N=${1:-2}
RUN="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
timeout 300 $RUN --master-port 29552 bench.py --gpus $N --steps 3 --warmup 2 > gpurun_out/bench_${N}gpu_final.log 2>&1; grep -E '^\{' gpurun_out/bench_${N}gpu_final.log | cut -c1-300
