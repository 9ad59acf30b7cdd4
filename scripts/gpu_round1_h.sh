#!/bin/bash
# session H: micro-op v3 + store permutations
set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests -x -q -m gpu > gpurun_out/pytest_gpu_h.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu_h.log
tail -5 gpurun_out/pytest_gpu_h.log
timeout 600 python bench.py --steps 5 --warmup 3 > gpurun_out/bench_h.json 2> gpurun_out/bench_h.err; echo "bench rc=$?"
cat gpurun_out/bench_h.json; tail -5 gpurun_out/bench_h.err
timeout 400 ncu --metrics gpu__time_duration.sum,smsp__inst_executed.sum --clock-control none -c 60 --csv --log-file gpurun_out/launches_h.csv python scripts/prof_driver.py 30 all > gpurun_out/prof_h.log 2>&1
QVMCUDA_REG_BITS=3 timeout 400 ncu --metrics gpu__time_duration.sum,smsp__inst_executed.sum --clock-control none -c 60 --csv --log-file gpurun_out/launches_h_m3.csv python scripts/prof_driver.py 30 fused > gpurun_out/prof_h_m3.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:qv_tile_kernel -c 2 -o gpurun_out/prof_tile_h python scripts/prof_driver.py 30 fused > gpurun_out/prof_full_h.log 2>&1
