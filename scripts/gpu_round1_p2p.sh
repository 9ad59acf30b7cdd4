#!/bin/bash
set -x
mkdir -p gpurun_out
timeout 300 ./scripts/p2p_probe 4096 > gpurun_out/p2p_probe.txt 2>&1; cat gpurun_out/p2p_probe.txt
timeout 600 python -m pytest tests -x -q -m gpu > gpurun_out/pytest_gpu_f.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu_f.log; tail -4 gpurun_out/pytest_gpu_f.log
