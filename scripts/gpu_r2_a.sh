#!/bin/bash
# round 2, session A: first run of the pass compiler on hardware -- parity suite through compiled passes, bench A/B
# (compiled vs interpreted), launch list + full ncu capture of the compiled QFT-30 passes.
set -x
mkdir -p gpurun_out
nproc; nvidia-smi -L
# (1) parity: every fused pass with >= 3 micro-ops runs through its compiled kernel (NVRTC on the box)
QVMCUDA_TRACE= timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r2a_pytest.log 2>&1; echo "pytest rc=$?"; tail -5 gpurun_out/r2a_pytest.log
# (2) bench, compiled passes (disk cache shipped in-tree) vs a cold cache vs the interpreter
timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/r2a_bench_jit.json 2> gpurun_out/r2a_bench_jit.err; cat gpurun_out/r2a_bench_jit.json; tail -3 gpurun_out/r2a_bench_jit.err
QVMCUDA_JIT_CACHE=/tmp/qvj_cold QVMCUDA_TRACE=1 timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/r2a_bench_cold.json 2> gpurun_out/r2a_bench_cold.err; cat gpurun_out/r2a_bench_cold.json; grep qvjit gpurun_out/r2a_bench_cold.err | head
QVMCUDA_JIT=off timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/r2a_bench_interp.json 2> gpurun_out/r2a_bench_interp.err; cat gpurun_out/r2a_bench_interp.json
# (3) other configurations, compiled vs interpreted
timeout 600 python scripts/bench_configs.py c1 c3 c4 > gpurun_out/r2a_configs_jit.jsonl 2> gpurun_out/r2a_configs_jit.err; cat gpurun_out/r2a_configs_jit.jsonl
QVMCUDA_JIT=off timeout 600 python scripts/bench_configs.py c3 c4 > gpurun_out/r2a_configs_interp.jsonl 2> gpurun_out/r2a_configs_interp.err
# (4) profiles
timeout 400 ncu --metrics gpu__time_duration.sum,smsp__inst_executed.sum --clock-control none -c 60 --csv --log-file gpurun_out/r2a_launches.csv python scripts/prof_driver.py 30 all > gpurun_out/r2a_prof.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:qvj_kernel -c 4 -o /tmp/r2a_full python scripts/prof_driver.py 30 fused > gpurun_out/r2a_prof_full.log 2>&1
python scripts/summarize_profile.py gpurun_out/r2a_launches.csv /tmp/r2a_full.ncu-rep gpurun_out/r2a_summary.md "round 2 capture A: compiled passes" > /dev/null 2>&1
ncu -i /tmp/r2a_full.ncu-rep --page source --csv --print-source sass --kernel-id :::0 2>/dev/null | gzip > gpurun_out/r2a_source_pass0.csv.gz
ncu -i /tmp/r2a_full.ncu-rep --page details --csv 2>/dev/null | gzip > gpurun_out/r2a_details.csv.gz
ls -la gpurun_out | tail -15
