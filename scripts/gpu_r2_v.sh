#!/bin/bash
# round 2, session V (1 GPU): sanity of the final binary (schedule cache / lazy reset / parity subsets, short bench), then an A/B of the
# "rolled group loop" generator variant on QFT-30 (it won 5 % on the 12-wire pass, gpurun_out/r2r_wide_pass.jsonl).
set -x
mkdir -p gpurun_out
timeout 400 python -m pytest tests/test_gpu_parity.py tests/test_gpu_boundary.py -x -q -k "schedule_cache or lazy or compiled_tape_reuse or qft_parity or shared_memory or call_stack" > gpurun_out/r2v_pytest.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/r2v_pytest.log
timeout 200 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --unfused-gates 4 > gpurun_out/r2v_bench.json 2>/dev/null; python -c "
import json; d=json.loads(open('gpurun_out/r2v_bench.json').read().strip().splitlines()[-1]); print('default', d['value'], d['ms_per_step'], d['e2e']['value'])"
export QVMCUDA_JIT_CACHE=/tmp/qvj_v17
QVMCUDA_JIT_VARIANT=17 timeout 200 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --unfused-gates 4 > gpurun_out/r2v_bench_v17.json 2>/dev/null; python -c "
import json; d=json.loads(open('gpurun_out/r2v_bench_v17.json').read().strip().splitlines()[-1]); print('rolled+tma', d['value'], d['ms_per_step'], d['e2e']['value'])"
