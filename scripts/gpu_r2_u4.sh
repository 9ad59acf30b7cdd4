#!/bin/bash
# round 2, session U (4 GPUs): the N=4 bench line (QFT-34) with the final build and its in-run parity check (both modes, from a
# random state and from the collective reset).
set -x
mkdir -p gpurun_out
( time timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29531 bench.py --gpus 4 --steps 3 --warmup 3 > gpurun_out/r2u_bench4.json 2> gpurun_out/r2u_bench4.err ) 2>&1 | tail -3; echo "bench rc=$?"
python -c "
import json; d=json.loads(open('gpurun_out/r2u_bench4.json').read().strip().splitlines()[-1]); print(d['value'], d['ms_per_step'], d['e2e']['value'], d['parity_check'], d['config']['pass_compiler'], d['c5_random']['seconds_per_circuit'], d['roofline']['nvlink']['achieved'], d['roofline']['local_passes'])"; tail -3 gpurun_out/r2u_bench4.err
