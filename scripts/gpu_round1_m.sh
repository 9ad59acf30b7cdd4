#!/bin/bash
# session M: trimmed dispatch path, expectation row, new GPU tests
set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests -x -q -m gpu > gpurun_out/pytest_gpu_m.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu_m.log
tail -4 gpurun_out/pytest_gpu_m.log
timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/bench_m.json 2> gpurun_out/bench_m.err; echo "bench rc=$?"
cat gpurun_out/bench_m.json | cut -c1-400; tail -5 gpurun_out/bench_m.err
timeout 400 ncu --metrics gpu__time_duration.sum,smsp__inst_executed.sum --clock-control none -c 60 --csv --log-file gpurun_out/launches_m.csv python scripts/prof_driver.py 30 fused > gpurun_out/prof_m.log 2>&1
python scripts/parse_launches.py gpurun_out/launches_m.csv
