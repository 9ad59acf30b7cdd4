#!/bin/bash
# session L: profile of record for the threaded micro-op kernel (captures summarised on the box: reps are large)
set -x
mkdir -p gpurun_out
timeout 400 ncu --metrics gpu__time_duration.sum,smsp__inst_executed.sum --clock-control none -c 60 --csv --log-file gpurun_out/launches_k.csv python scripts/prof_driver.py 30 all > gpurun_out/prof_k.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:qv_tile_kernel -c 9 -o /tmp/prof_tile_k python scripts/prof_driver.py 30 all > gpurun_out/prof_full_k.log 2>&1
python scripts/summarize_profile.py gpurun_out/launches_k.csv /tmp/prof_tile_k.ncu-rep gpurun_out/summary_k.md "capture K" > /dev/null 2>&1
ncu -i /tmp/prof_tile_k.ncu-rep --page source --csv --print-source sass --kernel-id :::8 > gpurun_out/source_k_fused.csv 2>/dev/null
ncu -i /tmp/prof_tile_k.ncu-rep --page source --csv --print-source sass --kernel-id :::1 > gpurun_out/source_k_single.csv 2>/dev/null
ls -la /tmp/prof_tile_k.ncu-rep gpurun_out/
gzip -f gpurun_out/source_k_fused.csv gpurun_out/source_k_single.csv
