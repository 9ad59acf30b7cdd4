#!/bin/bash
# session O: butterfly + gated-diagonal fusion
set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests -x -q -m gpu > gpurun_out/pytest_gpu_o.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu_o.log
tail -4 gpurun_out/pytest_gpu_o.log
timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/bench_o.json 2> gpurun_out/bench_o.err; echo "bench rc=$?"
cat gpurun_out/bench_o.json | cut -c1-300; tail -5 gpurun_out/bench_o.err
timeout 400 ncu --metrics gpu__time_duration.sum,smsp__inst_executed.sum --clock-control none -c 60 --csv --log-file gpurun_out/launches_o.csv python scripts/prof_driver.py 30 all > gpurun_out/prof_o.log 2>&1
python scripts/parse_launches.py gpurun_out/launches_o.csv
