#!/bin/bash
# launch list of the bench command itself (profiler pass: its printed numbers are not bench values)
mkdir -p gpurun_out
timeout 100 ncu --metrics gpu__time_duration.sum --clock-control none -c 30 --csv --log-file gpurun_out/launches_bench.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline --unfused-gates 4 > gpurun_out/launches_bench.log 2>&1
python scripts/parse_launches.py gpurun_out/launches_bench.csv | head -40
