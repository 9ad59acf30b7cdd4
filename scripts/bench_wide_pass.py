"""A/B for passes that mix 12 wires (the first local pass of a sharded QFT): 8 amplitudes per thread need 4 rounds, 16 amplitudes per
thread 3.  One-pass circuit on a 30-qubit state: the QFT's Hadamards on 12 chosen wires + every controlled phase that touches them.
Run once per setting of QVMCUDA_REG_BITS (3 / 4); prints one JSON line."""
import json
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from qvm_b200 import circuits, qvm  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 30
chosen = set(range(4)) | set(range(n - 8, n))
full = circuits.qft_circuit(range(n))
gates = []
for m, q in full:
    if len(q) == 1 and q[0] in chosen:
        gates.append((m, q))
    elif len(q) == 2 and np.count_nonzero(m - np.diag(np.diag(m))) == 0 and (q[0] in chosen or q[1] in chosen):
        gates.append((m, q))
vec = qvm.DeviceVector(1 << n)
vec.set_zero_state()
tape = qvm.Tape(n, gates, fuse=True)
desc = tape.describe()
for _ in range(3):
    vec.run_tape(tape)
vec.synchronize()
t0 = time.perf_counter()
reps = 10
for _ in range(reps):
    vec.run_tape(tape)
vec.synchronize()
dt = (time.perf_counter() - t0) / reps
print(json.dumps({"reg_bits": os.environ.get("QVMCUDA_REG_BITS", "default"), "variant": os.environ.get("QVMCUDA_JIT_VARIANT", "auto"),
                  "gates": len(gates), "passes": tape.info()["passes"], "ms": 1e3 * dt, "norm2": vec.norm2(),
                  "schedule": [l.strip()[:110] for l in desc.splitlines()[1:-2]]}))
