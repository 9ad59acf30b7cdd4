#!/bin/bash
# session I: flat-switch (threaded) dispatch + prefetched micro-op headers
set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests -x -q -m gpu > gpurun_out/pytest_gpu_i.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu_i.log
tail -3 gpurun_out/pytest_gpu_i.log
timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/bench_i.json 2> gpurun_out/bench_i.err; echo "bench rc=$?"
cat gpurun_out/bench_i.json; tail -5 gpurun_out/bench_i.err
QVMCUDA_REG_BITS=3 timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/bench_i_m3.json 2> gpurun_out/bench_i_m3.err
cat gpurun_out/bench_i_m3.json
timeout 400 ncu --metrics gpu__time_duration.sum,smsp__inst_executed.sum --clock-control none -c 60 --csv --log-file gpurun_out/launches_i.csv python scripts/prof_driver.py 30 all > gpurun_out/prof_i.log 2>&1
QVMCUDA_REG_BITS=3 timeout 400 ncu --metrics gpu__time_duration.sum,smsp__inst_executed.sum --clock-control none -c 60 --csv --log-file gpurun_out/launches_i_m3.csv python scripts/prof_driver.py 30 fused > gpurun_out/prof_i_m3.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:qv_tile_kernel -c 1 -o gpurun_out/prof_tile_i python scripts/prof_driver.py 30 fused > gpurun_out/prof_full_i.log 2>&1
