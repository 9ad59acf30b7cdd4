#!/bin/bash
# first GPU session: parity, smoke, bench, launch list, one full capture
set -x
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,memory.total,clocks.max.sm --format=csv > gpurun_out/gpu.txt 2>&1
timeout 900 python -m pytest tests -x -q -m gpu > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
tail -30 gpurun_out/pytest_gpu.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?" >> gpurun_out/smoke.log
cat gpurun_out/smoke.log
timeout 600 python bench.py --steps 3 --warmup 3 > gpurun_out/bench_b.json 2> gpurun_out/bench_b.err; echo "bench rc=$?"
cat gpurun_out/bench_b.json; tail -5 gpurun_out/bench_b.err
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/launches_b.csv python scripts/prof_driver.py 30 all > gpurun_out/prof_b.log 2>&1
tail -5 gpurun_out/prof_b.log
timeout 600 ncu --set full --clock-control none --import-source on -k regex:qv_tile_kernel -c 9 -o gpurun_out/prof_tile_b python scripts/prof_driver.py 30 all > gpurun_out/prof_full_b.log 2>&1
tail -3 gpurun_out/prof_full_b.log
ls -la gpurun_out
