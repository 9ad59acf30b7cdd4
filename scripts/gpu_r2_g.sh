#!/bin/bash
# round 2, session G: A/B of the diagonal planner's single-table rule x the wide swizzle on QFT-30 and random layers;
# dense-gate kernels after the address-arithmetic change; new boundary tests.
set -x
mkdir -p gpurun_out
for sc in 1 0; do for ws in 1 0; do
  export QVMCUDA_DIAG_SINGLE_CHUNK=$sc QVMCUDA_JIT_WIDE_SWZ=$ws QVMCUDA_JIT_CACHE=/tmp/qvj_${sc}${ws}
  timeout 300 python bench.py --steps 8 --warmup 3 --no-cpu-baseline --unfused-gates 4 > gpurun_out/r2g_bench_sc${sc}_ws${ws}.json 2>/dev/null
  timeout 300 python scripts/bench_configs.py c3 c4 > gpurun_out/r2g_configs_sc${sc}_ws${ws}.jsonl 2>/dev/null
  python - <<PY
import json
d=json.loads(open('gpurun_out/r2g_bench_sc${sc}_ws${ws}.json').read().strip().splitlines()[-1])
print('single_chunk=${sc} wide_swz=${ws}: QFT-30', round(d['ms_per_step'],2), 'ms', round(d['value']), 'gates/s, e2e', round(d['e2e']['value']))
for l in open('gpurun_out/r2g_configs_sc${sc}_ws${ws}.jsonl'):
    c=json.loads(l); print('     ', c['config'][:44], c.get('fused',{}).get('ms', c.get('ms')))
PY
done; done
unset QVMCUDA_DIAG_SINGLE_CHUNK QVMCUDA_JIT_WIDE_SWZ QVMCUDA_JIT_CACHE
timeout 300 python scripts/bench_dense.py > gpurun_out/r2g_dense_dmma.jsonl 2> gpurun_out/r2g_dense_dmma.err; cat gpurun_out/r2g_dense_dmma.jsonl; tail -2 gpurun_out/r2g_dense_dmma.err
timeout 900 python -m pytest tests/test_gpu_boundary.py -x -q > gpurun_out/r2g_pytest_boundary.log 2>&1; echo "pytest rc=$?"; tail -12 gpurun_out/r2g_pytest_boundary.log
