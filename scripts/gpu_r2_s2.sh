#!/bin/bash
# round 2, session S (2 GPUs): known-zero shards (lazy clear, zero-filling pull loads): sharded parity in both exchange modes and in
# the single-process group, then the N=2 bench line (e2e is the number that should move).
set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_sharded.py -x -q -k "2 or single_process" > gpurun_out/r2s_pytest_sharded2.log 2>&1; echo "pytest rc=$?"; tail -8 gpurun_out/r2s_pytest_sharded2.log
( time timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 4 --warmup 3 > gpurun_out/r2s_bench2.json 2> gpurun_out/r2s_bench2.err ) 2>&1 | tail -3; echo "bench rc=$?"
python -c "
import json; d=json.loads(open('gpurun_out/r2s_bench2.json').read().strip().splitlines()[-1]); print(d['value'], d['ms_per_step'], d['e2e'], d['parity_check']['ok'], d['c5_random']['seconds_per_circuit'], d['config']['pass_compiler'])"; tail -3 gpurun_out/r2s_bench2.err
