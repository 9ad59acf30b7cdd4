#!/bin/bash
# short 8-GPU session: sharded parity at world 8, weak-scaling bench (QFT-33 at 16 GiB/GPU) + QFT-35
set -x
mkdir -p gpurun_out
RUN="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1"
timeout 300 python -m pytest tests/test_gpu_sharded.py -x -q -m gpu -k "8 and False" > gpurun_out/pytest_8gpu_final.log 2>&1; tail -2 gpurun_out/pytest_8gpu_final.log
timeout 300 $RUN --master-port 29542 bench.py --gpus 8 --steps 4 --warmup 2 > gpurun_out/bench_8gpu_qft33_final.log 2>&1; grep -E '^\{' gpurun_out/bench_8gpu_qft33_final.log | cut -c1-300
QVM_DIST_TRACE=1 timeout 600 $RUN --master-port 29543 bench.py --gpus 8 --qubits 32 --steps 2 --warmup 1 > gpurun_out/bench_8gpu_qft35_final.log 2>&1; grep -E '^\{' gpurun_out/bench_8gpu_qft35_final.log | cut -c1-300; grep -E 'dist\]' gpurun_out/bench_8gpu_qft35_final.log | tail -9
