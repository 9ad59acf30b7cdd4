"""A/B of the dense k-qubit gate kernels at 28 qubits (run under gpurun): QVMCUDA_BIG=scalar python scripts/bench_dense.py
vs python scripts/bench_dense.py.  One JSON line per k."""
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from helpers import rand_unitary  # noqa: E402
from qvm_b200 import qvm  # noqa: E402

n = int(os.environ.get("QVM_DENSE_QUBITS", "28"))
vec = qvm.DeviceVector(1 << n)
vec.set_zero_state()
rng = np.random.default_rng(1)
path = "scalar" if os.environ.get("QVMCUDA_BIG") == "scalar" else "dmma"
for k in (3, 4, 5, 6, 8):
    for label, qs in (("low", tuple(range(k))), ("spread", tuple(int(x) for x in np.linspace(1, n - 2, k).astype(int)))):
        m = rand_unitary(k, rng)
        vec.apply_matrix(m, qs)
        vec.synchronize()
        reps = 3
        t0 = time.perf_counter()
        for _ in range(reps):
            vec.apply_matrix(m, qs)
        vec.synchronize()
        dt = (time.perf_counter() - t0) / reps
        flops = 8.0 * (1 << k) * (1 << n)       # 4 * 2^k real FMA per amplitude
        print(json.dumps({"kernel": path, "k": k, "qubits": label, "n": n, "ms": 1e3 * dt, "hbm_gbs": 32 * (1 << n) / dt / 1e9,
                          "fp64_tflops": flops / dt / 1e12}), flush=True)
print(json.dumps({"norm2": vec.norm2()}))
vec.close()
