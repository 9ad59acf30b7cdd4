#!/bin/bash
# round 2, session B: host-side matrix fusion + generator variants on the dense-heavy configurations
# (random 1q/CZ layers, density QAOA), 16-amplitude rounds under the pass compiler.
set -x
mkdir -p gpurun_out
for v in 0 1 2; do
  QVMCUDA_JIT_VARIANT=$v timeout 300 python scripts/bench_configs.py c3 c4 > gpurun_out/r2b_configs_v$v.jsonl 2> gpurun_out/r2b_configs_v$v.err
done
QVMCUDA_REG_BITS=4 timeout 300 python scripts/bench_configs.py c3 c4 > gpurun_out/r2b_configs_m4.jsonl 2> gpurun_out/r2b_configs_m4.err
QVMCUDA_JIT=off timeout 300 python scripts/bench_configs.py c3 c4 > gpurun_out/r2b_configs_interp.jsonl 2> gpurun_out/r2b_configs_interp.err
for f in gpurun_out/r2b_configs_*.jsonl; do echo $f; python - "$f" <<'PY'
import json,sys
for l in open(sys.argv[1]):
    d=json.loads(l); print('  ', d['config'][:44], d.get('fused',{}).get('ms', d.get('ms')), d.get('fused',{}).get('passes', d.get('passes')))
PY
done
QVMCUDA_REG_BITS=4 timeout 300 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/r2b_bench_m4.json 2> gpurun_out/r2b_bench_m4.err; cut -c1-200 gpurun_out/r2b_bench_m4.json
QVMCUDA_JIT_VARIANT=1 timeout 300 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/r2b_bench_v1.json 2> gpurun_out/r2b_bench_v1.err; cut -c1-200 gpurun_out/r2b_bench_v1.json
# profile: random layers 25 q and the density circuit through compiled passes
cat > /tmp/prof_c.py <<'PY'
import sys, os
sys.path.insert(0, os.getcwd()); sys.path.insert(0, os.path.join(os.getcwd(), 'tests'))
from qvm_b200 import circuits, gates as G, qvm
n = 25
vec = qvm.DeviceVector(1 << n); vec.set_zero_state()
tape = qvm.Tape(n, circuits.random_layer_circuit(n, 10, 0), fuse=True); print(tape.describe())
vec.run_tape(tape); vec.synchronize(); vec.close()
n = 14
circ = circuits.qaoa_maxcut_circuit(n, circuits.line_graph(n)); dep = G.depolarizing_kraus_map(0.01)
ops = []
for m, q in circ:
    ops.append((m, q)); ops.extend((dep, (qq,)) for qq in q)
st = qvm.DensityMatrixState(n); gl = qvm.density_gate_list(n, ops)
tape = qvm.Tape(2 * n, gl, fuse=True); print(tape.describe())
st.vec.set_zero_state(); st.vec.run_tape(tape); st.vec.synchronize()
PY
timeout 400 ncu --metrics gpu__time_duration.sum,smsp__inst_executed.sum,sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active,smsp__issue_active.avg.pct_of_peak_sustained_active,dram__throughput.avg.pct_of_peak_sustained_elapsed --clock-control none -c 60 --csv --log-file gpurun_out/r2b_launches_c.csv python /tmp/prof_c.py > gpurun_out/r2b_prof_c.log 2>&1
tail -40 gpurun_out/r2b_prof_c.log
python - <<'PY'
import csv
rows=[r for r in csv.reader(open('gpurun_out/r2b_launches_c.csv')) if len(r)>5]
h=rows[0]; ik,iv,im,ii=h.index("Kernel Name"),h.index("Metric Value"),h.index("Metric Name"),h.index("ID")
by={}
for r in rows[1:]:
    by.setdefault(r[ii],{'k':r[ik][:30]})[r[im]]=r[iv]
for i,d in by.items(): print(i,d)
PY
