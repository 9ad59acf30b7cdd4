#!/bin/bash
set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests -x -q -m gpu > gpurun_out/pytest_gpu_e.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu_e.log
tail -5 gpurun_out/pytest_gpu_e.log
QVMCUDA_TRACE=1 timeout 300 python scripts/e2e_breakdown.py 30 > gpurun_out/e2e_breakdown_e.txt 2>&1; cat gpurun_out/e2e_breakdown_e.txt | tail -25
timeout 600 python bench.py --steps 5 --warmup 3 > gpurun_out/bench_e.json 2> gpurun_out/bench_e.err; echo "bench rc=$?"
cat gpurun_out/bench_e.json; tail -5 gpurun_out/bench_e.err
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/launches_e.csv python scripts/prof_driver.py 30 all > gpurun_out/prof_e.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:qv_tile_kernel -c 9 -o gpurun_out/prof_tile_e python scripts/prof_driver.py 30 all > gpurun_out/prof_full_e.log 2>&1
tail -3 gpurun_out/prof_full_e.log
