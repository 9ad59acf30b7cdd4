#!/bin/bash
# round 2, session Q (8 GPUs): the N=8 bench line (QFT-35, 64 GiB per GPU: absorbed SWAPs, hoisted exchange, prewarmed kernels) with
# its in-run parity check, then the world-8 sharded parity test in pull mode.
set -x
mkdir -p gpurun_out
( time timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29521 bench.py --gpus 8 --steps 4 --warmup 3 > gpurun_out/r2q_bench8.json 2> gpurun_out/r2q_bench8.err ) 2>&1 | tail -3; echo "bench rc=$?"
tail -1 gpurun_out/r2q_bench8.json | cut -c1-400; tail -3 gpurun_out/r2q_bench8.err
timeout 400 python -m pytest tests/test_gpu_sharded.py -x -q -k "8-False" > gpurun_out/r2q_pytest_sharded8.log 2>&1; echo "pytest rc=$?"; tail -5 gpurun_out/r2q_pytest_sharded8.log
