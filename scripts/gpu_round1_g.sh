#!/bin/bash
# session G: micro-op kernel + store permutations: bench (auto / m=3 / m=4), launch list, full capture of the fused passes
set -x
mkdir -p gpurun_out
timeout 600 python bench.py --steps 5 --warmup 3 > gpurun_out/bench_g.json 2> gpurun_out/bench_g.err; echo "bench rc=$?"
cat gpurun_out/bench_g.json; tail -5 gpurun_out/bench_g.err
timeout 400 ncu --metrics gpu__time_duration.sum,smsp__inst_executed.sum --clock-control none -c 60 --csv --log-file gpurun_out/launches_g.csv python scripts/prof_driver.py 30 all > gpurun_out/prof_g.log 2>&1
QVMCUDA_REG_BITS=3 timeout 400 ncu --metrics gpu__time_duration.sum,smsp__inst_executed.sum --clock-control none -c 60 --csv --log-file gpurun_out/launches_g_m3.csv python scripts/prof_driver.py 30 fused > gpurun_out/prof_g_m3.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:qv_tile_kernel -c 4 -o gpurun_out/prof_tile_g python scripts/prof_driver.py 30 fused > gpurun_out/prof_full_g.log 2>&1
QVMCUDA_REG_BITS=3 timeout 600 ncu --set full --clock-control none --import-source on -k regex:qv_tile_kernel -c 1 -o gpurun_out/prof_tile_g_m3 python scripts/prof_driver.py 30 fused > gpurun_out/prof_full_g_m3.log 2>&1
ls -la gpurun_out/*.ncu-rep
