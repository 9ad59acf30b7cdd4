#!/bin/bash
# round 2, session K: swap routing (QFT-30 in 5 passes), full GPU suite, bench line, launch list of the routed schedule,
# A/B against QVMCUDA_ROUTE_SWAPS=0.
set -x
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/r2k_pytest.log 2>&1; echo "pytest rc=$?"; tail -5 gpurun_out/r2k_pytest.log
timeout 120 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2k_smoke.log 2>&1; tail -2 gpurun_out/r2k_smoke.log
timeout 600 python bench.py --steps 10 --warmup 3 > gpurun_out/r2k_bench.json 2> gpurun_out/r2k_bench.err; echo "bench rc=$?"; cut -c1-400 gpurun_out/r2k_bench.json; tail -3 gpurun_out/r2k_bench.err
QVMCUDA_ROUTE_SWAPS=0 timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --unfused-gates 4 > gpurun_out/r2k_bench_noroute.json 2>/dev/null; cut -c1-200 gpurun_out/r2k_bench_noroute.json
timeout 300 ncu --metrics gpu__time_duration.sum,smsp__inst_executed.sum,dram__throughput.avg.pct_of_peak_sustained_elapsed,sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active,smsp__issue_active.avg.pct_of_peak_sustained_active,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 20 --csv --log-file gpurun_out/r2k_launches.csv python scripts/prof_driver.py 30 fused > gpurun_out/r2k_prof.log 2>&1
python - <<'PY'
import csv
rows=[r for r in csv.reader(open('gpurun_out/r2k_launches.csv')) if len(r)>5]
h=rows[0]; ik,iv,im,ii=h.index("Kernel Name"),h.index("Metric Value"),h.index("Metric Name"),h.index("ID")
by={}
for r in rows[1:]:
    by.setdefault(r[ii],{'k':r[ik][:24]})[r[im].split('.')[0][-24:]]=r[iv]
for i,d in by.items(): print(i,d)
PY
