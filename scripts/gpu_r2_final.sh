#!/bin/bash
# round 2, last call: the full GPU suite + smoke + the bench line with the final build, the other configurations, and the ncu
# launch list of the bench command itself (numbers printed under ncu are not bench values).
set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests -x -q -m gpu > gpurun_out/r2z_pytest.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/r2z_pytest.log
timeout 120 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2z_smoke.log 2>&1; tail -2 gpurun_out/r2z_smoke.log
timeout 300 python bench.py > gpurun_out/r2z_bench.json 2> gpurun_out/r2z_bench.err; echo "bench rc=$?"; python -c "
import json; d=json.loads(open('gpurun_out/r2z_bench.json').read().strip().splitlines()[-1]); print(d['value'], d['ms_per_step'], d['e2e']['value'], d['roofline']['frac'], d['unfused']['frac_of_peak'], d['cpu_baseline']['value'], d['ratios'], d['gpu_launches'], d['clocks'])"; tail -2 gpurun_out/r2z_bench.err
timeout 200 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/r2z_bench_ref.json 2>/dev/null; cut -c1-300 gpurun_out/r2z_bench_ref.json
timeout 300 python scripts/bench_configs.py > gpurun_out/r2z_configs.jsonl 2> gpurun_out/r2z_configs.err; cut -c1-330 gpurun_out/r2z_configs.jsonl
timeout 200 ncu --metrics gpu__time_duration.sum,smsp__inst_executed.sum --clock-control none -c 60 --csv --log-file gpurun_out/r2z_launches_bench.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline --unfused-gates 4 > gpurun_out/r2z_launches_bench.log 2>&1
python scripts/parse_launches.py gpurun_out/r2z_launches_bench.csv | head -70
