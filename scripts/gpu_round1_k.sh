#!/bin/bash
# session K: threaded micro-op kernel (profile of record), persistent sampler scratch
set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests -x -q -m gpu > gpurun_out/pytest_gpu_k.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu_k.log
tail -3 gpurun_out/pytest_gpu_k.log
timeout 600 python bench.py --steps 5 --warmup 3 > gpurun_out/bench_k.json 2> gpurun_out/bench_k.err; echo "bench rc=$?"
cat gpurun_out/bench_k.json; tail -5 gpurun_out/bench_k.err
QVMCUDA_TRACE=1 timeout 300 python scripts/e2e_breakdown.py 30 > gpurun_out/e2e_breakdown_k.txt 2>&1; tail -9 gpurun_out/e2e_breakdown_k.txt
timeout 400 ncu --metrics gpu__time_duration.sum,smsp__inst_executed.sum --clock-control none -c 60 --csv --log-file gpurun_out/launches_k.csv python scripts/prof_driver.py 30 all > gpurun_out/prof_k.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:qv_tile_kernel -c 2 -o gpurun_out/prof_tile_k_single python scripts/prof_driver.py 30 unfused > gpurun_out/prof_full_k1.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:qv_tile_kernel -c 6 -o gpurun_out/prof_tile_k_fused python scripts/prof_driver.py 30 fused > gpurun_out/prof_full_k2.log 2>&1
ls -la gpurun_out/*.ncu-rep
