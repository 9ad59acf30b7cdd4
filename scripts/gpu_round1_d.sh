#!/bin/bash
set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests -x -q -m gpu > gpurun_out/pytest_gpu_d.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu_d.log
tail -5 gpurun_out/pytest_gpu_d.log
timeout 300 python scripts/e2e_breakdown.py 30 > gpurun_out/e2e_breakdown.txt 2>&1; cat gpurun_out/e2e_breakdown.txt
timeout 900 python scripts/bench_configs.py c1 c3 c4 qft > gpurun_out/configs_d.jsonl 2> gpurun_out/configs_d.err; cat gpurun_out/configs_d.jsonl; tail -5 gpurun_out/configs_d.err
