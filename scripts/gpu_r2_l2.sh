#!/bin/bash
# round 2, session L (2 GPUs): sharded parity (both exchange modes, absorbed and executed SWAPs), the N=2 bench line as the driver
# runs it, a traced run for per-step times, and an ncu launch list of a small sharded run with NVLink counters.
set -x
mkdir -p gpurun_out
nvidia-smi -L | wc -l
timeout 900 python -m pytest tests/test_gpu_sharded.py -x -q -k "2" > gpurun_out/r2l_pytest_sharded2.log 2>&1; echo "pytest rc=$?"; tail -6 gpurun_out/r2l_pytest_sharded2.log
( time timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 4 --warmup 3 > gpurun_out/r2l_bench2.json 2> gpurun_out/r2l_bench2.err ) 2>&1 | tail -3; echo "bench rc=$?"
tail -1 gpurun_out/r2l_bench2.json | cut -c1-1500; tail -3 gpurun_out/r2l_bench2.err
QVM_DIST_TRACE=1 timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29513 bench.py --gpus 2 --steps 2 --warmup 2 --c5-layers 0 --no-parity-check > gpurun_out/r2l_bench2_trace.log 2>&1; grep "^\[dist\]" gpurun_out/r2l_bench2_trace.log | tail -24
ncu --query-metrics 2>/dev/null | grep -i "nvl\|fabric" | head -60 > gpurun_out/r2l_nvlink_metric_names.txt; wc -l gpurun_out/r2l_nvlink_metric_names.txt
M="gpu__time_duration.sum,smsp__inst_executed.sum,dram__bytes_read.sum,dram__bytes_write.sum,nvlrx__bytes.sum,nvltx__bytes.sum,lts__t_sectors_srcunit_ltcfabric.sum"
timeout 400 ncu --target-processes all --metrics $M --clock-control none -k regex:"qvj_kernel|qv_tile_kernel" -c 24 --csv --log-file gpurun_out/r2l_sharded_launches_%p.csv python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29515 scripts/prof_sharded.py 28 2 > gpurun_out/r2l_prof_sharded.log 2>&1; echo "ncu rc=$?"; tail -5 gpurun_out/r2l_prof_sharded.log
ls -la gpurun_out/r2l_sharded_launches_* ; head -c 1500 $(ls gpurun_out/r2l_sharded_launches_* | head -1)
