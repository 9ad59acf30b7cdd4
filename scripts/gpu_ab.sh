#!/bin/bash
# A/B: group loop not unrolled (default build) vs unrolled x2 (libqvmcuda_unroll.so)
set -x
mkdir -p gpurun_out
for v in default unroll; do
  if [ $v = unroll ]; then export QVMCUDA_LIB=$PWD/qvm_b200/libqvmcuda_unroll.so; fi
  timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/bench_ab_$v.json 2> gpurun_out/bench_ab_$v.err
  python - <<PY
import json
d=json.loads(open('gpurun_out/bench_ab_$v.json').read().strip().split('\n')[-1])
print('$v', round(d['value']), round(d['ms_per_step'],2), 'e2e', round(d['e2e']['value']), 'unfused', round(d['unfused']['ms_per_gate_pass'],3))
PY
done
