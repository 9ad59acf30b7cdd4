#!/bin/bash
# GPU session: single-GPU tests + bench + profile of the cp.async kernel
set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests -x -q -m gpu > gpurun_out/pytest_gpu_c.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu_c.log
tail -5 gpurun_out/pytest_gpu_c.log
timeout 600 python bench.py --steps 5 --warmup 3 > gpurun_out/bench_c.json 2> gpurun_out/bench_c.err; echo "bench rc=$?"
cat gpurun_out/bench_c.json; tail -5 gpurun_out/bench_c.err
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/launches_c.csv python scripts/prof_driver.py 30 all > gpurun_out/prof_c.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:qv_tile_kernel -c 9 -o gpurun_out/prof_tile_c python scripts/prof_driver.py 30 all > gpurun_out/prof_full_c.log 2>&1
tail -3 gpurun_out/prof_full_c.log
