#!/bin/bash
# round 2, session N (1 GPU): schedule cache + lazy reset through the full suite, bench line.
set -x
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/r2n_pytest.log 2>&1; echo "pytest rc=$?"; tail -5 gpurun_out/r2n_pytest.log
timeout 600 python bench.py --steps 10 --warmup 3 > gpurun_out/r2n_bench.json 2> gpurun_out/r2n_bench.err; echo "bench rc=$?"; python -c "
import json; d=json.loads(open('gpurun_out/r2n_bench.json').read().strip().splitlines()[-1]); print(d['value'], d['ms_per_step'], d['e2e'], d['config']['pass_compiler'], d['roofline']['frac'], d['cpu_baseline']['value'])"; tail -3 gpurun_out/r2n_bench.err
timeout 120 python scripts/e2e_breakdown.py > gpurun_out/r2n_e2e_breakdown.txt 2>&1; tail -12 gpurun_out/r2n_e2e_breakdown.txt
