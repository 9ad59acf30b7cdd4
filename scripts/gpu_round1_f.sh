#!/bin/bash
# session F: micro-op tile kernel (flat pre-resolved ops, in-place FP64, 16 amplitudes per thread)
set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests -x -q -m gpu > gpurun_out/pytest_gpu_g.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu_g.log
tail -5 gpurun_out/pytest_gpu_g.log
timeout 600 python bench.py --steps 5 --warmup 3 > gpurun_out/bench_g.json 2> gpurun_out/bench_g.err; echo "bench rc=$?"
cat gpurun_out/bench_g.json; tail -5 gpurun_out/bench_g.err
QVMCUDA_REG_BITS=3 timeout 600 python bench.py --steps 5 --warmup 3 > gpurun_out/bench_g_m3.json 2> gpurun_out/bench_g_m3.err; echo "bench rc=$?"
cat gpurun_out/bench_g_m3.json
QVMCUDA_REG_BITS=4 timeout 600 python bench.py --steps 5 --warmup 3 > gpurun_out/bench_g_m4.json 2> gpurun_out/bench_g_m4.err; echo "bench rc=$?"
cat gpurun_out/bench_g_m4.json
timeout 400 ncu --metrics gpu__time_duration.sum,smsp__inst_executed.sum --clock-control none -c 60 --csv --log-file gpurun_out/launches_g.csv python scripts/prof_driver.py 30 all > gpurun_out/prof_g.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:qv_tile_kernel -c 11 -o gpurun_out/prof_tile_g python scripts/prof_driver.py 30 all > gpurun_out/prof_full_g.log 2>&1
tail -3 gpurun_out/prof_full_g.log
