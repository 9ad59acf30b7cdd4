"""Times every public-API call of bench.py's e2e step separately (30 qubits)."""
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from qvm_b200 import circuits, qvm  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 30
vec = qvm.DeviceVector(1 << n)
gates = circuits.qft_circuit(range(n))
u = np.random.default_rng(1).random(1000)


def t(name, fn, reps=3):
    fn()
    vec.synchronize()
    t0 = time.perf_counter()
    for _ in range(reps):
        fn()
        vec.synchronize()
    print(f"{name:28s} {1e3 * (time.perf_counter() - t0) / reps:9.3f} ms", flush=True)


t("set_zero_state", vec.set_zero_state)
t("apply_gates(fused)", lambda: vec.apply_gates(gates, fuse=True))
tape = qvm.Tape(n, gates, fuse=True)
t("run_tape(fused)", lambda: vec.run_tape(tape))
t("sample(1000)", lambda: vec.sample(u))
t("sample(100000)", lambda: vec.sample(np.random.default_rng(2).random(100000)))
t("prob_excited(0)", lambda: vec.prob_excited(0))
t("norm2", vec.norm2)
t("collapse(3)", lambda: vec.collapse(3, 1, 1.0))
t("scale", lambda: vec.scale(1.0))
t("reset + apply_gates(fused)", lambda: (vec.set_zero_state(), vec.apply_gates(gates, fuse=True)))
