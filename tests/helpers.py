"""Shared test helpers (tests only)."""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SUPPORT = os.path.join(ROOT, "tests", "support")
_EMU = None


def rand_state(n_amps_log2: int, seed: int = 12345) -> np.ndarray:
    """numpy default_rng(seed), standard-normal re/im, normalised in f64."""
    r = np.random.default_rng(seed)
    v = r.standard_normal(1 << n_amps_log2) + 1j * r.standard_normal(1 << n_amps_log2)
    return np.ascontiguousarray(v / np.linalg.norm(v), dtype=np.complex128)


def rand_unitary(k: int, rng) -> np.ndarray:
    d = 1 << k
    a = rng.standard_normal((d, d)) + 1j * rng.standard_normal((d, d))
    q, r = np.linalg.qr(a)
    return np.ascontiguousarray(q * (np.diag(r) / np.abs(np.diag(r))), dtype=np.complex128)


def emulator():
    """TEST-ONLY CPU interpreter of the tile programs (tests/support/qv_emulator.cpp)."""
    global _EMU
    if _EMU is None:
        so = os.path.join(SUPPORT, "libqvemu.so")
        srcs = [os.path.join(SUPPORT, "qv_emulator.cpp"), os.path.join(ROOT, "qvm_b200", "csrc", "qv_sched.cpp"),
                os.path.join(ROOT, "qvm_b200", "csrc", "qv_jit_gen.cpp")]
        deps = srcs + [os.path.join(ROOT, "qvm_b200", "csrc", f) for f in ("qv_ops.h", "qv_program.h", "qv_sched.h", "qv_jit.h")]
        def stale():
            return not os.path.exists(so) or any(os.path.getmtime(d) > os.path.getmtime(so) for d in deps)
        if stale():
            # pytest-xdist workers race here after a source change: one builds (into a temporary name, renamed when complete),
            # the others wait for the lock and find the library fresh
            import fcntl
            with open(so + ".lock", "w") as lock:
                fcntl.flock(lock, fcntl.LOCK_EX)
                if stale():
                    tmp = f"{so}.{os.getpid()}.tmp"
                    subprocess.check_call(["/usr/bin/g++", "-O2", "-std=c++17", "-fPIC", "-shared", "-Wno-unknown-pragmas",
                                           "-o", tmp] + srcs + ["-ldl"])
                    os.replace(tmp, so)
        _EMU = C.CDLL(so)
        _EMU.qvtest_run.restype = C.c_int
    return _EMU


def emulator_jit_host(on: bool) -> int:
    """Switch the emulator between interpreting the micro-ops and running the pass compiler's generated C++
    (compiled for the host with g++ -DQVJ_HOST).  Returns the number of passes run compiled so far."""
    emu = emulator()
    emu.qvtest_set_jit_host.restype = C.c_int
    return emu.qvtest_set_jit_host(int(on), os.path.join(ROOT, "qvm_b200", "csrc").encode())


def flatten_circuit(circ):
    """[(matrix, quil-order qubits)] -> (ks, qubits LSB-first flat, matrices flat as doubles)."""
    ks = np.array([len(q) for _, q in circ], dtype=np.int32)
    qf = np.array([x for _, q in circ for x in reversed(q)], dtype=np.int32)
    mf = np.concatenate([np.ascontiguousarray(m, dtype=np.complex128).ravel() for m, _ in circ]).view(np.float64)
    return ks, qf, np.ascontiguousarray(mf)


def run_emulator(psi, n, circ, fuse=True, tile_bits=12, absorb_swaps=False, reg_bits=0, route_swaps=-1, l2p_in=None):
    """reg_bits: 3 / 4 force the 8- / 16-amplitudes-per-thread round format, 0 = the scheduler's own choice.
    route_swaps: 1 / 0 force / forbid swap routing (absorbed SWAPs, permutation executed by the passes' write-backs),
    -1 = the scheduler's cost model.  l2p_in: layout the state is in when the circuit starts (default: canonical)."""
    ks, qf, mf = flatten_circuit(circ)
    emulator().qvtest_set_reg_bits(int(reg_bits))
    emulator().qvtest_set_route_swaps(int(route_swaps))
    desc = C.create_string_buffer(1 << 16)
    l2p = np.arange(n, dtype=np.int32) if l2p_in is None else np.ascontiguousarray(l2p_in, dtype=np.int32).copy()
    p = lambda a: a.ctypes.data_as(C.c_void_p)
    rc = emulator().qvtest_run(p(psi), n, len(circ), p(ks), p(qf), p(mf), int(fuse), tile_bits, int(absorb_swaps),
                               p(l2p), desc, len(desc))
    if rc < 0:
        raise RuntimeError(desc.value.decode())
    return rc, desc.value.decode(), l2p


def run_oracle(psi, circ):
    from oracle import oracle as O
    for m, q in circ:
        O.apply_matrix(psi, m, q)
    return psi


def assert_close(a, b, rel=1e-12, abs_=1e-14):
    """north_star tolerance: 1e-12 relative / 1e-14 absolute on amplitudes."""
    a = np.asarray(a)
    b = np.asarray(b)
    err = np.abs(a - b)
    tol = abs_ + rel * np.abs(b)
    bad = err > tol
    assert not bad.any(), f"max err {err.max():.3e} at {int(np.argmax(err))}; {int(bad.sum())} entries out of tolerance"


def random_circuit(n, n_gates, rng, max_dense=3):
    from qvm_b200 import gates as G
    circ = []
    names1 = ["H", "X", "Y", "Z", "S", "T"]
    for _ in range(n_gates):
        kind = rng.integers(0, 12)
        if kind == 0:
            circ.append((G.gate_matrix(names1[rng.integers(0, len(names1))]), (int(rng.integers(0, n)),)))
        elif kind == 1:
            nm = ["RX", "RY", "RZ", "PHASE"][rng.integers(0, 4)]
            circ.append((G.gate_matrix(nm, [rng.uniform(0, 6.28)]), (int(rng.integers(0, n)),)))
        elif kind == 2:
            circ.append((rand_unitary(1, rng), (int(rng.integers(0, n)),)))
        elif n >= 2 and kind in (3, 4):
            a, b = rng.choice(n, 2, replace=False)
            nm = ["CNOT", "CZ", "SWAP", "ISWAP"][rng.integers(0, 4)]
            circ.append((G.gate_matrix(nm), (int(a), int(b))))
        elif n >= 2 and kind == 5:
            a, b = rng.choice(n, 2, replace=False)
            nm = ["CPHASE", "CPHASE01", "PISWAP", "RZZ", "RXX"][rng.integers(0, 5)]
            circ.append((G.gate_matrix(nm, [rng.uniform(0, 6.28)]), (int(a), int(b))))
        elif n >= 2 and kind == 6:
            a, b = rng.choice(n, 2, replace=False)
            circ.append((rand_unitary(2, rng), (int(a), int(b))))
        elif n >= 3 and kind == 7:
            a, b, c = rng.choice(n, 3, replace=False)
            nm = ["CCNOT", "CSWAP"][rng.integers(0, 2)]
            circ.append((G.gate_matrix(nm), (int(a), int(b), int(c))))
        elif n >= 3 and kind == 8 and max_dense >= 3:
            q = rng.choice(n, 3, replace=False)
            circ.append((rand_unitary(3, rng), tuple(int(x) for x in q)))
        elif n >= 3 and kind == 9:
            # controlled random 2q unitary and a forked 1q rotation
            q = rng.choice(n, 3, replace=False)
            if rng.integers(0, 2):
                circ.append((G.controlled(rand_unitary(2, rng)), tuple(int(x) for x in q)))
            else:
                circ.append((G.forked(G.gate_matrix("RX", [rng.uniform(0, 6.28)]), G.gate_matrix("RX", [rng.uniform(0, 6.28)])),
                             (int(q[0]), int(q[1]))))
        elif n >= 4 and kind == 10 and max_dense >= 4:
            q = rng.choice(n, 4, replace=False)
            circ.append((rand_unitary(4, rng), tuple(int(x) for x in q)))
        else:
            d = np.exp(1j * rng.uniform(0, 6.28, size=8))
            k = min(3, n)
            q = rng.choice(n, k, replace=False)
            circ.append((np.diag(d[: 1 << k]), tuple(int(x) for x in q)))
    return circ


def run_emulator_sharded(psi, n, world, circ, fuse=True, tile_bits=12, absorb_swaps=False, remap_pull=False, reg_bits=0):
    """Emulates `world` ranks (shards back to back in psi) running the sharded schedule; remap_pull selects
    out-of-place pull remaps into alternate shard buffers instead of in-place peer passes."""
    ks, qf, mf = flatten_circuit(circ)
    desc = C.create_string_buffer(1 << 16)
    l2p = np.arange(n, dtype=np.int32)
    p = lambda a: a.ctypes.data_as(C.c_void_p)
    emu = emulator()
    emu.qvtest_run_sharded.restype = C.c_int
    emu.qvtest_set_remap_pull(int(remap_pull))
    emu.qvtest_set_reg_bits(int(reg_bits))
    rc = emu.qvtest_run_sharded(p(psi), n, world, len(circ), p(ks), p(qf), p(mf), int(fuse), tile_bits,
                                int(absorb_swaps), p(l2p), desc, len(desc))
    if rc < 0:
        raise RuntimeError(desc.value.decode())
    return rc // 1000, rc % 1000, desc.value.decode(), l2p


def unpermute(psi_phys, l2p):
    """psi_phys is indexed by physical bits; logical qubit q lives at physical bit l2p[q]."""
    n = len(l2p)
    idx = np.arange(psi_phys.size)
    phys = np.zeros_like(idx)
    for q in range(n):
        phys |= ((idx >> q) & 1) << int(l2p[q])
    return psi_phys[phys]


def load_bench_circuit(name):
    """Benchmark inputs of the reference (bench/*.quil), from the committed fixture
    tests/golden/bench_circuits.json -> ([(matrix, qubits)], measures, n_qubits)."""
    import json
    from qvm_b200 import gates as G
    data = json.load(open(os.path.join(ROOT, "tests", "golden", "bench_circuits.json")))[name]
    circ, measures = [], []
    for ins in data["instructions"]:
        if ins[0] == "G":
            circ.append((G.gate_matrix(ins[1], ins[2]), tuple(ins[3])))
        else:
            measures.append((ins[1], tuple(ins[2]) if ins[2] else None))
    return circ, measures, data["n_qubits"]
