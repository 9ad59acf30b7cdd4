#!/bin/bash
# round 2, session X (1 GPU): ncu --set full capture of the five compiled passes of the final QFT-30 schedule (swap routing).
set -x
mkdir -p gpurun_out
timeout 120 ncu --metrics gpu__time_duration.sum,smsp__inst_executed.sum --clock-control none -c 20 --csv --log-file gpurun_out/r2x_launches.csv python scripts/prof_driver.py 30 fused > gpurun_out/r2x_prof.log 2>&1
timeout 240 ncu --set full --clock-control none --import-source on -k regex:qvj_kernel -c 5 -o /tmp/r2x_full python scripts/prof_driver.py 30 fused > gpurun_out/r2x_prof_full.log 2>&1; echo "ncu rc=$?"
python scripts/summarize_profile.py gpurun_out/r2x_launches.csv /tmp/r2x_full.ncu-rep gpurun_out/r2x_summary.md "round 2 capture X: the five compiled passes of the final QFT-30 schedule (swap routing)" > /dev/null 2>&1
ncu -i /tmp/r2x_full.ncu-rep --page raw --csv 2>/dev/null | gzip > gpurun_out/r2x_raw.csv.gz
ls -la gpurun_out | grep r2x; head -70 gpurun_out/r2x_summary.md
