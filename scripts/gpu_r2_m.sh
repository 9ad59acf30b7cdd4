#!/bin/bash
# round 2, session M: lazy reset (SET-TO-ZERO-STATE as a flag; the first compiled pass synthesises its tiles): parity, bench (e2e).
set -x
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py -x -q -k "lazy" > gpurun_out/r2m_pytest_lazy.log 2>&1; echo "pytest rc=$?"; tail -15 gpurun_out/r2m_pytest_lazy.log
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/r2m_pytest.log 2>&1; echo "pytest rc=$?"; tail -5 gpurun_out/r2m_pytest.log
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --unfused-gates 8 > gpurun_out/r2m_bench.json 2> gpurun_out/r2m_bench.err; echo "bench rc=$?"; python -c "
import json; d=json.loads(open('gpurun_out/r2m_bench.json').read().strip().splitlines()[-1]); print(d['value'], d['ms_per_step'], d['e2e'], d['config']['pass_compiler'])"; tail -3 gpurun_out/r2m_bench.err
QVMCUDA_LAZY_RESET=0 timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --unfused-gates 8 > gpurun_out/r2m_bench_nolazy.json 2>/dev/null; python -c "
import json; d=json.loads(open('gpurun_out/r2m_bench_nolazy.json').read().strip().splitlines()[-1]); print(d['value'], d['ms_per_step'], d['e2e'])"
timeout 120 python scripts/e2e_breakdown.py > gpurun_out/r2m_e2e_breakdown.txt 2>&1; tail -20 gpurun_out/r2m_e2e_breakdown.txt
