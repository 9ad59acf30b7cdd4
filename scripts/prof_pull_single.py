"""ncu driver for an EXCHANGE pass: one process, two (or more) devices, one shard each (qvmcuda_shard_attach_local), a few QFT
circuits in a row.  A multi-process run cannot be profiled (ncu fails on IPC-mapped peer memory, gpurun_out/r2l_prof_sharded.log);
here the pull passes of every shard appear in one launch list, with the NVLink counters of the device that runs them.
usage: prof_pull_single.py LOCAL_QUBITS WORLD RUNS"""
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from qvm_b200 import circuits  # noqa: E402
from qvm_b200.dist import LocalShardGroup  # noqa: E402

local = int(sys.argv[1]) if len(sys.argv) > 1 else 28
world = int(sys.argv[2]) if len(sys.argv) > 2 else 2
runs = int(sys.argv[3]) if len(sys.argv) > 3 else 3
n = local + world.bit_length() - 1
grp = LocalShardGroup(n, list(range(world)))
gates = circuits.qft_circuit(range(n))
for r in range(runs):
    grp._sync()
    t0 = time.perf_counter()
    grp.apply_gates(gates)
    grp._sync()
    print(f"circuit {r}: {1e3 * (time.perf_counter() - t0):.2f} ms, steps so far {grp.steps}, exchange steps {grp.peer_steps}", flush=True)
print("norm2", grp.norm2())
grp.close()
