#!/bin/bash
# round 2, session O (2 GPUs): remap hoisting + schedule cache on shards: sharded parity (both exchange modes), the N=2 bench line as
# the driver runs it, then a traced run for per-step times.
set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_sharded.py -x -q -k "2" > gpurun_out/r2o_pytest_sharded2.log 2>&1; echo "pytest rc=$?"; tail -6 gpurun_out/r2o_pytest_sharded2.log
( time timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 4 --warmup 3 > gpurun_out/r2o_bench2.json 2> gpurun_out/r2o_bench2.err ) 2>&1 | tail -3; echo "bench rc=$?"
tail -1 gpurun_out/r2o_bench2.json | cut -c1-600; tail -3 gpurun_out/r2o_bench2.err
QVM_DIST_TRACE=1 timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29513 bench.py --gpus 2 --steps 2 --warmup 2 --c5-layers 0 --no-parity-check > gpurun_out/r2o_bench2_trace.log 2>&1; grep "^\[dist\]" gpurun_out/r2o_bench2_trace.log | tail -16
