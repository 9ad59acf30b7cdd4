#!/bin/bash
# 8-GPU session: sharded parity at world 8 (fused pull + in place), weak-scaling bench (QFT-33 at 16 GiB/GPU),
# the 35-qubit runs (64 GiB/GPU + 64 GiB alternate buffer)
set -x
mkdir -p gpurun_out
RUN="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1"
timeout 600 python -m pytest tests/test_gpu_sharded.py -x -q -m gpu -k "8" > gpurun_out/pytest_8gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_8gpu.log
tail -5 gpurun_out/pytest_8gpu.log
timeout 300 $RUN --master-port 29542 bench.py --gpus 8 --steps 3 --warmup 2 > gpurun_out/bench_8gpu_qft33.log 2>&1; grep -E '^\{' gpurun_out/bench_8gpu_qft33.log | cut -c1-300
QVM_DIST_TRACE=1 timeout 300 $RUN --master-port 29541 bench.py --gpus 8 --steps 1 --warmup 1 > gpurun_out/bench_8gpu_qft33_trace.log 2>&1; grep -E 'dist\]' gpurun_out/bench_8gpu_qft33_trace.log | tail -14
QVM_DIST_TRACE=1 timeout 600 $RUN --master-port 29543 bench.py --gpus 8 --qubits 32 --steps 2 --warmup 1 > gpurun_out/bench_8gpu_qft35.log 2>&1; grep -E '^\{' gpurun_out/bench_8gpu_qft35.log | cut -c1-300; grep -E 'dist\]' gpurun_out/bench_8gpu_qft35.log | tail -14; tail -3 gpurun_out/bench_8gpu_qft35.log | cut -c1-300
QVM_DIST_TRACE=1 timeout 600 $RUN --master-port 29544 bench.py --gpus 8 --qubits 32 --workload random --layers 4 --steps 2 --warmup 1 > gpurun_out/bench_8gpu_random35.log 2>&1; grep -E '^\{' gpurun_out/bench_8gpu_random35.log | cut -c1-300; grep -E 'dist\]' gpurun_out/bench_8gpu_random35.log | tail -16; tail -3 gpurun_out/bench_8gpu_random35.log | cut -c1-300
