#!/bin/bash
# round 2, session D: FP64 probe with cycle counts; boundary tests; bench; profile of record for QFT-30 (launch list + full capture)
set -x
mkdir -p gpurun_out
./scripts/fp64_probe > gpurun_out/r2d_fp64_probe.txt 2>&1; cat gpurun_out/r2d_fp64_probe.txt
timeout 900 python -m pytest tests/test_gpu_boundary.py -x -q > gpurun_out/r2d_pytest_boundary.log 2>&1; echo "pytest rc=$?"; tail -15 gpurun_out/r2d_pytest_boundary.log
timeout 600 python bench.py --steps 10 --warmup 3 > gpurun_out/r2d_bench.json 2> gpurun_out/r2d_bench.err; cut -c1-300 gpurun_out/r2d_bench.json; tail -3 gpurun_out/r2d_bench.err
timeout 300 python scripts/bench_configs.py c1 c3 c4 > gpurun_out/r2d_configs.jsonl 2> gpurun_out/r2d_configs.err
python - <<'PY'
import json
for l in open('gpurun_out/r2d_configs.jsonl'):
    d=json.loads(l); print('  ', d['config'][:44], d.get('fused',{}).get('ms', d.get('ms')), d.get('fused',{}).get('passes', d.get('passes')), d.get('api_fused_ms'))
PY
timeout 400 ncu --metrics gpu__time_duration.sum,smsp__inst_executed.sum --clock-control none -c 60 --csv --log-file gpurun_out/r2d_launches.csv python scripts/prof_driver.py 30 all > gpurun_out/r2d_prof.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:qvj_kernel -c 4 -o /tmp/r2d_full python scripts/prof_driver.py 30 fused > gpurun_out/r2d_prof_full.log 2>&1
python scripts/summarize_profile.py gpurun_out/r2d_launches.csv /tmp/r2d_full.ncu-rep gpurun_out/r2d_summary.md "round 2 capture D: compiled passes (QFT-30)" > /dev/null 2>&1
ncu -i /tmp/r2d_full.ncu-rep --page raw --csv 2>/dev/null | gzip > gpurun_out/r2d_raw.csv.gz
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 40 --csv --log-file gpurun_out/r2d_launches_bench.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline --unfused-gates 4 > gpurun_out/r2d_bench_under_ncu.log 2>&1
ls -la gpurun_out | grep r2d
