#!/bin/bash
# round 2, session C: FP64 vector vs tensor (DMMA) probe; Euler split A/B on the dense-heavy configurations; the new bench
# line; the GPU parity suite with test durations.
set -x
mkdir -p gpurun_out
./scripts/fp64_probe > gpurun_out/r2c_fp64_probe.txt 2>&1; cat gpurun_out/r2c_fp64_probe.txt
for e in 0 1 -1; do
  QVMCUDA_EULER=$e timeout 300 python scripts/bench_configs.py c3 c4 > gpurun_out/r2c_configs_euler$e.jsonl 2> gpurun_out/r2c_configs_euler$e.err
  echo "euler $e"; python - "gpurun_out/r2c_configs_euler$e.jsonl" <<'PY'
import json,sys
for l in open(sys.argv[1]):
    d=json.loads(l); print('  ', d['config'][:44], d.get('fused',{}).get('ms', d.get('ms')), d.get('fused',{}).get('passes', d.get('passes')))
PY
done
QVMCUDA_EULER=1 QVMCUDA_JIT_VARIANT=0 timeout 300 python scripts/bench_configs.py c3 > gpurun_out/r2c_configs_euler1_v0.jsonl 2>/dev/null; python - <<'PY'
import json
for l in open('gpurun_out/r2c_configs_euler1_v0.jsonl'):
    d=json.loads(l); print(' e1v0 ', d['config'][:44], d.get('fused',{}).get('ms', d.get('ms')))
PY
timeout 600 python bench.py --steps 10 --warmup 3 > gpurun_out/r2c_bench.json 2> gpurun_out/r2c_bench.err; cat gpurun_out/r2c_bench.json; tail -3 gpurun_out/r2c_bench.err
timeout 300 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r2c_bench_ref.json 2> gpurun_out/r2c_bench_ref.err; cat gpurun_out/r2c_bench_ref.json
timeout 1500 python -m pytest tests -m gpu -x -q --durations=15 > gpurun_out/r2c_pytest.log 2>&1; echo "pytest rc=$?"; tail -25 gpurun_out/r2c_pytest.log
