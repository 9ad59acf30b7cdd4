#!/bin/bash
# round 2, session H (8 GPUs): the N=8 bench line (35 qubits, 64 GiB per GPU) with its in-run parity check (both exchange
# modes), NVLink record and the random 1q/CZ circuit; then the world-8 sharded parity test in pull mode.
set -x
mkdir -p gpurun_out
nvidia-smi -L | wc -l; nproc; free -g | head -2
timeout 1200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29521 bench.py --gpus 8 --steps 3 --warmup 2 > gpurun_out/r2h_bench8.json 2> gpurun_out/r2h_bench8.err; echo "bench rc=$?"
tail -1 gpurun_out/r2h_bench8.json | cut -c1-6000; tail -5 gpurun_out/r2h_bench8.err
timeout 600 python -m pytest tests/test_gpu_sharded.py -x -q -k "8-False" > gpurun_out/r2h_pytest_sharded8.log 2>&1; echo "pytest rc=$?"; tail -5 gpurun_out/r2h_pytest_sharded8.log
