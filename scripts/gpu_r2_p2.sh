#!/bin/bash
# round 2, session P (2 GPUs): single-process shard group (qvmcuda_shard_attach_local): parity, then ncu launch lists of the
# exchange passes with the NVLink counters.
set -x
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_sharded.py -x -q -k "single_process" > gpurun_out/r2p_pytest_local.log 2>&1; echo "pytest rc=$?"; tail -8 gpurun_out/r2p_pytest_local.log
timeout 200 python scripts/prof_pull_single.py 30 2 3 > gpurun_out/r2p_single_timing.log 2>&1; cat gpurun_out/r2p_single_timing.log | tail -6
M="gpu__time_duration.sum,smsp__inst_executed.sum,dram__bytes_read.sum,dram__bytes_write.sum,nvlrx__bytes.sum,nvltx__bytes.sum"
timeout 300 ncu --metrics $M --clock-control none -k regex:"qvj_kernel|qv_tile_kernel" -c 30 --csv --log-file gpurun_out/r2p_pull_launches.csv python scripts/prof_pull_single.py 28 2 3 > gpurun_out/r2p_prof_pull.log 2>&1; echo "ncu rc=$?"; tail -5 gpurun_out/r2p_prof_pull.log
M2="gpu__time_duration.sum,nvlrx__bytes_data_user.sum,nvltx__bytes_data_user.sum,lts__t_requests_srcunit_ltcfabric.sum,lts__t_sectors_srcunit_ltcfabric.sum,lts__ltcfabric2lts_cycles_active.avg.pct_of_peak_sustained_elapsed,dram__throughput.avg.pct_of_peak_sustained_elapsed"
timeout 300 ncu --metrics $M2 --clock-control none -k regex:"qvj_kernel|qv_tile_kernel" -c 30 --csv --log-file gpurun_out/r2p_pull_launches2.csv python scripts/prof_pull_single.py 28 2 3 > gpurun_out/r2p_prof_pull2.log 2>&1; echo "ncu rc=$?"; tail -3 gpurun_out/r2p_prof_pull2.log
python - <<'PY'
import csv
for f in ('gpurun_out/r2p_pull_launches.csv','gpurun_out/r2p_pull_launches2.csv'):
    try:
        rows=[r for r in csv.reader(open(f)) if len(r)>5]
        h=rows[0]; ik,iv,im,ii=h.index("Kernel Name"),h.index("Metric Value"),h.index("Metric Name"),h.index("ID")
        by={}
        for r in rows[1:]:
            by.setdefault(r[ii],{'k':r[ik][:14]})[r[im][:34]]=r[iv]
        for i,d in by.items(): print(i,d)
    except Exception as e: print(f, e)
PY
