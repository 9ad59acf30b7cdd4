#!/bin/bash
# session J: 8- vs 16-amplitude groups on the other configurations, M=3 full capture
set -x
mkdir -p gpurun_out
for m in 3 4; do
  QVMCUDA_REG_BITS=$m timeout 600 python scripts/bench_configs.py c3 c4 qft > gpurun_out/configs_j_m$m.jsonl 2> gpurun_out/configs_j_m$m.err
  python - <<PY
import json
for l in open('gpurun_out/configs_j_m$m.jsonl'):
    d=json.loads(l)
    print('m=$m', d['config'][:50], 'fused ms', round(d.get('fused',{}).get('ms',d.get('ms',0)),3), 'passes', d.get('fused',{}).get('passes',d.get('passes')))
PY
done
QVMCUDA_REG_BITS=3 timeout 600 ncu --set full --clock-control none --import-source on -k regex:qv_tile_kernel -c 1 -o gpurun_out/prof_tile_j_m3 python scripts/prof_driver.py 30 fused > gpurun_out/prof_full_j_m3.log 2>&1
QVMCUDA_TRACE=1 timeout 300 python scripts/e2e_breakdown.py 30 > gpurun_out/e2e_breakdown_j.txt 2>&1; tail -12 gpurun_out/e2e_breakdown_j.txt
