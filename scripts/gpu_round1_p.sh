#!/bin/bash
# session P: profile of record after the butterfly / fusion work (summarised on the box), full bench with CPU baseline,
# other configurations
set -x
mkdir -p gpurun_out
timeout 600 python bench.py --steps 5 --warmup 3 > gpurun_out/bench_p.json 2> gpurun_out/bench_p.err; echo "bench rc=$?"
cat gpurun_out/bench_p.json; tail -3 gpurun_out/bench_p.err
timeout 300 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_p_ref.json 2> gpurun_out/bench_p_ref.err; cat gpurun_out/bench_p_ref.json
timeout 600 python scripts/bench_configs.py c1 c3 c4 qft > gpurun_out/configs_p.jsonl 2> gpurun_out/configs_p.err
QVMCUDA_TRACE=1 timeout 300 python scripts/e2e_breakdown.py 30 > gpurun_out/e2e_breakdown_p.txt 2>&1; tail -9 gpurun_out/e2e_breakdown_p.txt
timeout 400 ncu --metrics gpu__time_duration.sum,smsp__inst_executed.sum --clock-control none -c 60 --csv --log-file gpurun_out/launches_p.csv python scripts/prof_driver.py 30 all > gpurun_out/prof_p.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:qv_tile_kernel -c 9 -o /tmp/prof_tile_p python scripts/prof_driver.py 30 all > gpurun_out/prof_full_p.log 2>&1
python scripts/summarize_profile.py gpurun_out/launches_p.csv /tmp/prof_tile_p.ncu-rep gpurun_out/summary_p.md "capture P" > /dev/null 2>&1
ncu -i /tmp/prof_tile_p.ncu-rep --page source --csv --print-source sass --kernel-id :::8 2>/dev/null | gzip > gpurun_out/source_p_fused.csv.gz
ls -la gpurun_out | tail -12
