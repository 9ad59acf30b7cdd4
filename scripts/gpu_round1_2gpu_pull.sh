#!/bin/bash
# 2-GPU session: sharded parity over NVLink in both remap modes + sharded bench: fused pull / separate pull / in place
set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_sharded.py -x -q -m gpu > gpurun_out/pytest_2gpu_pull.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_2gpu_pull.log
tail -30 gpurun_out/pytest_2gpu_pull.log
RUN="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
QVM_DIST_TRACE=1 timeout 600 $RUN --master-port 29531 bench.py --gpus 2 --steps 3 --warmup 2 > gpurun_out/bench_2gpu_pull.log 2>&1; echo "bench rc=$?"
grep -E '^\{|dist\]' gpurun_out/bench_2gpu_pull.log | tail -14
timeout 600 $RUN --master-port 29533 bench.py --gpus 2 --steps 3 --warmup 2 > gpurun_out/bench_2gpu_pull_notrace.log 2>&1; echo "bench rc=$?"
grep -E '^\{' gpurun_out/bench_2gpu_pull_notrace.log | cut -c1-200
QVM_REMAP_INPLACE=1 timeout 600 $RUN --master-port 29532 bench.py --gpus 2 --steps 3 --warmup 2 > gpurun_out/bench_2gpu_inplace.log 2>&1; echo "bench rc=$?"
grep -E '^\{' gpurun_out/bench_2gpu_inplace.log | cut -c1-200
