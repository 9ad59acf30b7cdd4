#!/bin/bash
# round 2, session J: classic kernel with the tile loaded by ONE tensor copy (QVMCUDA_JIT_VARIANT=16): parity + A/B; then the
# 2-GPU line again (deterministic random-circuit schedule) is left to session K.
set -x
mkdir -p gpurun_out
export QVMCUDA_JIT_CACHE=/tmp/qvj_tl
QVMCUDA_JIT_VARIANT=16 timeout 300 python -m pytest tests/test_gpu_parity.py -x -q -k "qft_parity or random_circuits or large_state or random_layer" > gpurun_out/r2j_pytest_tmaload.log 2>&1; echo "pytest rc=$?"; tail -6 gpurun_out/r2j_pytest_tmaload.log
for i in 1 2; do
QVMCUDA_JIT_VARIANT=16 timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --unfused-gates 4 > gpurun_out/r2j_bench_tmaload_$i.json 2> gpurun_out/r2j_bench_tmaload.err; echo "bench rc=$?"; cut -c1-200 gpurun_out/r2j_bench_tmaload_$i.json; tail -3 gpurun_out/r2j_bench_tmaload.err
QVMCUDA_JIT_CACHE=/tmp/qvj_cl timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --unfused-gates 4 > gpurun_out/r2j_bench_classic_$i.json 2>/dev/null; cut -c1-200 gpurun_out/r2j_bench_classic_$i.json
done
QVMCUDA_JIT_VARIANT=16 timeout 300 ncu --metrics gpu__time_duration.sum,smsp__inst_executed.sum,dram__throughput.avg.pct_of_peak_sustained_elapsed,sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active,smsp__issue_active.avg.pct_of_peak_sustained_active --clock-control none -c 20 --csv --log-file gpurun_out/r2j_launches_tmaload.csv python scripts/prof_driver.py 30 fused > gpurun_out/r2j_prof_tmaload.log 2>&1
python - <<'PY'
import csv
rows=[r for r in csv.reader(open('gpurun_out/r2j_launches_tmaload.csv')) if len(r)>5]
h=rows[0]; ik,iv,im,ii=h.index("Kernel Name"),h.index("Metric Value"),h.index("Metric Name"),h.index("ID")
by={}
for r in rows[1:]:
    by.setdefault(r[ii],{'k':r[ik][:24]})[r[im].split('.')[0][-24:]]=r[iv]
for i,d in by.items(): print(i,d)
PY
