"""Times the remaining BASELINE.json configurations on one GPU (run under gpurun); one JSON line each.
C1 20H, C3 5x4x25 / entangle-25 / random layers + 10^5-shot sampling, C4 14-qubit noisy QAOA on vec(rho)."""
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch  # noqa: E402

from helpers import load_bench_circuit  # noqa: E402
from qvm_b200 import _lib, circuits, gates as G, qvm  # noqa: E402


def timed(vec, fn, reps=5, warm=2):
    for _ in range(warm):
        fn()
    vec.synchronize()
    t0 = time.perf_counter()
    for _ in range(reps):
        fn()
    vec.synchronize()
    return (time.perf_counter() - t0) / reps


def pure(name, n, circ, shots=0):
    vec = qvm.DeviceVector(1 << n)
    out = {"config": name, "n_qubits": n, "gates": len(circ), "state_bytes": 16 << n}
    for fuse in (True, False):
        tape = qvm.Tape(n, circ, fuse=fuse)
        info = tape.info()

        def step():
            vec.set_zero_state()
            vec.run_tape(tape)
        dt = timed(vec, step)
        key = "fused" if fuse else "unfused"
        out[key] = {"ms": 1e3 * dt, "gates_per_s": len(circ) / dt, "passes": info["passes"],
                    "GBps_algorithmic": info["passes"] * 32 * (1 << n) / dt / 1e9}
    # through the immediate API (schedule + upload + run every call): what a Lisp RUN does
    dt = timed(vec, lambda: (vec.set_zero_state(), vec.apply_gates(circ, fuse=True)))
    out["api_fused_ms"] = 1e3 * dt
    if shots:
        u = np.random.default_rng(2024).random(shots)
        dt = timed(vec, lambda: vec.sample(u, strict=False), reps=3, warm=1)
        out["sampling"] = {"shots": shots, "ms": 1e3 * dt, "shots_per_s": shots / dt}
    out["norm2"] = vec.norm2()
    vec.close()
    print(json.dumps(out), flush=True)


def density(n):
    edges = circuits.line_graph(n)
    circ = circuits.qaoa_maxcut_circuit(n, edges)
    dep = G.depolarizing_kraus_map(0.01)
    ops = []
    for m, q in circ:
        ops.append((m, q))
        for qq in q:
            ops.append((dep, (qq,)))
    st = qvm.DensityMatrixState(n)
    gl = qvm.density_gate_list(n, ops)
    tape = qvm.Tape(2 * n, gl, fuse=True)
    info = tape.info()

    def step():
        st.vec.set_zero_state()
        st.vec.run_tape(tape)
    dt = timed(st.vec, step, reps=3, warm=1)
    probs = st.measurement_probabilities()
    out = {"config": f"density qaoa line-graph({n}) + depolarizing p=0.01 after every gate", "n_qubits": n,
           "vec_rho_bytes": 16 << (2 * n), "unitaries": len(circ), "channels": len(ops) - len(circ),
           "index_bit_gates": len(gl), "passes": info["passes"], "ms": 1e3 * dt,
           "ops_per_s": len(ops) / dt, "GBps_algorithmic": info["passes"] * 32 * (1 << (2 * n)) / dt / 1e9,
           "trace": float(probs.sum())}
    # the reference's traffic for the same ops: 2 passes per unitary, 656 B/elt per 1q channel
    out["reference_bytes_per_elt"] = 64 * len(circ) + 656 * (len(ops) - len(circ))
    out["our_bytes_per_elt"] = 32 * info["passes"]
    st.vec.close()
    print(json.dumps(out), flush=True)


if __name__ == "__main__":
    which = sys.argv[1:] or ["c1", "c3", "c4"]
    if "c1" in which:
        c, _, n = load_bench_circuit("20H")
        pure("C1 bench/20H.quil", n, c)
    if "c3" in which:
        c, _, n = load_bench_circuit("5x4x25")
        pure("C3 bench/5x4x25.quil", n, c, shots=100000)
        c, _, n = load_bench_circuit("entangle-25")
        pure("C3 bench/entangle-25.quil (gates only)", n, c, shots=100000)
        pure("C3 random layers 20q x 10 (seed 0)", 20, circuits.random_layer_circuit(20, 10, 0), shots=100000)
        pure("C3 random layers 25q x 10 (seed 0)", 25, circuits.random_layer_circuit(25, 10, 0), shots=100000)
    if "c4" in which:
        density(int(os.environ.get("QVM_DENSITY_QUBITS", "14")))
    if "qft" in which:
        for n in (20, 24, 26, 28):
            pure(f"QFT-{n}", n, circuits.qft_circuit(range(n)))
