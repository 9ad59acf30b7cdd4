#!/bin/bash
# round 2, session T (8 GPUs): the N=8 bench line with the known-zero-shards path (in-run parity check now also runs from the
# collective reset in both exchange modes).
set -x
mkdir -p gpurun_out
( time timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29521 bench.py --gpus 8 --steps 3 --warmup 2 --c5-layers 0 > gpurun_out/r2t_bench8.json 2> gpurun_out/r2t_bench8.err ) 2>&1 | tail -3; echo "bench rc=$?"
python -c "
import json; d=json.loads(open('gpurun_out/r2t_bench8.json').read().strip().splitlines()[-1]); print(d['value'], d['ms_per_step'], d['e2e'], d['parity_check'], d['config']['pass_compiler'])"; tail -3 gpurun_out/r2t_bench8.err
