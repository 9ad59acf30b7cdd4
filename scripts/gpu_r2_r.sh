#!/bin/bash
# round 2, session R (1 GPU): A/B for 12-wire passes (4 rounds x 8 amplitudes vs 3 rounds x 16 amplitudes vs 2 CTAs x 128 registers);
# the other BASELINE configurations with the final build.
set -x
mkdir -p gpurun_out
export QVMCUDA_JIT_CACHE=/tmp/qvj_r
for rb in 3 4; do QVMCUDA_REG_BITS=$rb timeout 200 python scripts/bench_wide_pass.py 30 2>&1 | tail -1 | cut -c1-400; done | tee gpurun_out/r2r_wide_pass.jsonl
for v in 2 18 1 17; do QVMCUDA_JIT_VARIANT=$v timeout 200 python scripts/bench_wide_pass.py 30 2>&1 | tail -1 | cut -c1-200; done | tee -a gpurun_out/r2r_wide_pass.jsonl
unset QVMCUDA_JIT_CACHE
timeout 600 python scripts/bench_configs.py > gpurun_out/r2r_configs.jsonl 2> gpurun_out/r2r_configs.err; cut -c1-420 gpurun_out/r2r_configs.jsonl; tail -3 gpurun_out/r2r_configs.err
