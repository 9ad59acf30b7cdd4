#!/bin/bash
# round 2, session F: wide swizzle for the bit-reversal write-back, dense-gate kernels A/B (DMMA vs scalar), boundary tests,
# bench, launch list + full capture of record.
set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_boundary.py -x -q > gpurun_out/r2f_pytest_boundary.log 2>&1; echo "pytest rc=$?"; tail -12 gpurun_out/r2f_pytest_boundary.log
timeout 600 python bench.py --steps 10 --warmup 3 > gpurun_out/r2f_bench.json 2> gpurun_out/r2f_bench.err; cut -c1-300 gpurun_out/r2f_bench.json; tail -3 gpurun_out/r2f_bench.err
timeout 300 python scripts/bench_dense.py > gpurun_out/r2f_dense_dmma.jsonl 2> gpurun_out/r2f_dense_dmma.err; cat gpurun_out/r2f_dense_dmma.jsonl; tail -2 gpurun_out/r2f_dense_dmma.err
QVMCUDA_BIG=scalar timeout 600 python scripts/bench_dense.py > gpurun_out/r2f_dense_scalar.jsonl 2> gpurun_out/r2f_dense_scalar.err; cat gpurun_out/r2f_dense_scalar.jsonl
timeout 400 ncu --metrics gpu__time_duration.sum,smsp__inst_executed.sum --clock-control none -c 60 --csv --log-file gpurun_out/r2f_launches.csv python scripts/prof_driver.py 30 all > gpurun_out/r2f_prof.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:qvj_kernel -c 4 -o /tmp/r2f_full python scripts/prof_driver.py 30 fused > gpurun_out/r2f_prof_full.log 2>&1
python scripts/summarize_profile.py gpurun_out/r2f_launches.csv /tmp/r2f_full.ncu-rep gpurun_out/r2f_summary.md "round 2 capture F: compiled passes (QFT-30)" > /dev/null 2>&1
head -60 gpurun_out/r2f_summary.md
QVM_DENSE_QUBITS=26 timeout 300 ncu --set full --clock-control none -k regex:qv_bigmma_kernel -c 3 -o /tmp/r2f_mma python scripts/bench_dense.py > gpurun_out/r2f_prof_mma.log 2>&1
ncu -i /tmp/r2f_mma.ncu-rep --page raw --csv 2>/dev/null | python -c "
import csv,sys
rr=list(csv.reader(sys.stdin)); h=rr[0]; idx={n:i for i,n in enumerate(h)}
keys=['gpu__time_duration.sum','dram__bytes_read.sum','dram__bytes_write.sum','sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active','smsp__issue_active.avg.pct_of_peak_sustained_active','sm__inst_executed_pipe_fp64.sum','launch__registers_per_thread','smsp__inst_executed.sum','sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active']
for r in rr[2:]:
    print({k:r[idx[k]] for k in keys if k in idx})
" > gpurun_out/r2f_mma_metrics.txt 2>&1; cat gpurun_out/r2f_mma_metrics.txt
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/r2f_pytest.log 2>&1; echo "pytest rc=$?"; tail -5 gpurun_out/r2f_pytest.log
