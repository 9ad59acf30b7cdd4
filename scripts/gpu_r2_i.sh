#!/bin/bash
# round 2, session I: the persistent TMA-fed variant of the compiled passes (QVMCUDA_JIT_VARIANT=8) -- parity, then A/B on QFT-30.
set -x
mkdir -p gpurun_out
export QVMCUDA_JIT_CACHE=/tmp/qvj_tma
QVMCUDA_JIT_VARIANT=8 timeout 300 python -m pytest tests/test_gpu_parity.py -x -q -k "qft_parity or random_circuits or large_state or random_layer" > gpurun_out/r2i_pytest_tma.log 2>&1; echo "pytest rc=$?"; tail -15 gpurun_out/r2i_pytest_tma.log
QVMCUDA_JIT_VARIANT=8 timeout 300 python bench.py --steps 8 --warmup 3 --no-cpu-baseline --unfused-gates 4 > gpurun_out/r2i_bench_tma.json 2> gpurun_out/r2i_bench_tma.err; echo "bench rc=$?"; cut -c1-260 gpurun_out/r2i_bench_tma.json; tail -3 gpurun_out/r2i_bench_tma.err
unset QVMCUDA_JIT_CACHE
timeout 300 python bench.py --steps 8 --warmup 3 --no-cpu-baseline --unfused-gates 4 > gpurun_out/r2i_bench_classic.json 2> gpurun_out/r2i_bench_classic.err; cut -c1-260 gpurun_out/r2i_bench_classic.json
QVMCUDA_JIT_CACHE=/tmp/qvj_tma QVMCUDA_JIT_VARIANT=8 timeout 300 ncu --metrics gpu__time_duration.sum,smsp__inst_executed.sum,dram__throughput.avg.pct_of_peak_sustained_elapsed,sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active,smsp__issue_active.avg.pct_of_peak_sustained_active --clock-control none -c 20 --csv --log-file gpurun_out/r2i_launches_tma.csv python scripts/prof_driver.py 30 fused > gpurun_out/r2i_prof_tma.log 2>&1
python - <<'PY'
import csv
rows=[r for r in csv.reader(open('gpurun_out/r2i_launches_tma.csv')) if len(r)>5]
h=rows[0]; ik,iv,im,ii=h.index("Kernel Name"),h.index("Metric Value"),h.index("Metric Name"),h.index("ID")
by={}
for r in rows[1:]:
    by.setdefault(r[ii],{'k':r[ik][:24]})[r[im].split('.')[0][-24:]]=r[iv]
for i,d in by.items(): print(i,d)
PY
