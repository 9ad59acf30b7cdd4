#!/bin/bash
# round 2, session E (2 GPUs): sharded parity tests in both exchange modes with the final build, then the N=2 bench line
# (33 qubits, 64 GiB per GPU) with its in-run parity check, NVLink roofline record and the random 1q/CZ circuit.
set -x
mkdir -p gpurun_out
nvidia-smi -L; nvidia-smi topo -m | head -12
timeout 900 python -m pytest tests/test_gpu_sharded.py -x -q -k "2" > gpurun_out/r2e_pytest_sharded2.log 2>&1; echo "pytest rc=$?"; tail -8 gpurun_out/r2e_pytest_sharded2.log
QVM_DIST_TRACE=1 timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 3 --warmup 2 --c5-layers 8 > gpurun_out/r2e_bench2.json 2> gpurun_out/r2e_bench2.err; echo "bench rc=$?"
grep -v "^\[dist\]" gpurun_out/r2e_bench2.json | tail -1 | cut -c1-3000; grep "^\[dist\]" gpurun_out/r2e_bench2.json | tail -40; tail -5 gpurun_out/r2e_bench2.err
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --impl reference --gpus 2 --steps 2 --warmup 1 > gpurun_out/r2e_bench2_ref.json 2> gpurun_out/r2e_bench2_ref.err; cat gpurun_out/r2e_bench2_ref.json | cut -c1-600
