"""Host logic under random configurations: random circuits (1-4 qubit dense gates, controlled / forked gates,
multi-qubit diagonals, Hadamard ladders) x tile size x 8-/16-amplitude rounds x 1/2/4/8 emulated ranks x in-place /
pull remaps x fusion x absorbed SWAPs, interpreted on the CPU through the kernel's own op code, against the oracle."""
import numpy as np
import pytest

from helpers import (assert_close, rand_state, random_circuit, run_emulator, run_emulator_sharded, run_oracle,
                     unpermute)
from qvm_b200 import gates as G


def _case(seed):
    rng = np.random.default_rng(seed)
    world = int(rng.choice([1, 1, 2, 4, 8]))
    g = world.bit_length() - 1
    n = int(rng.integers(max(2, g + 4), 14))
    nl = n - g
    tb = int(rng.integers(min(4, nl), min(12, nl) + 1))
    cfg = dict(world=world, n=n, tile_bits=tb, reg_bits=int(rng.choice([0, 3, 4])), fuse=bool(rng.integers(0, 2)),
               absorb_swaps=bool(rng.integers(0, 2)), remap_pull=bool(rng.integers(0, 2)))
    circ = random_circuit(n, int(rng.integers(1, 70)), rng, max_dense=4)
    for _ in range(int(rng.integers(0, 10))):
        circ.insert(int(rng.integers(0, len(circ) + 1)), (G.gate_matrix("H"), (int(rng.integers(0, n)),)))
    if n >= 5 and rng.integers(0, 2):
        qs = tuple(int(x) for x in rng.choice(n, 5, replace=False))
        circ.insert(int(rng.integers(0, len(circ) + 1)), (np.diag(np.exp(1j * rng.uniform(0, 6.28, size=32))), qs))
    return cfg, circ


@pytest.mark.parametrize("seed", range(48))
def test_random_configuration(seed):
    cfg, circ = _case(7000 + seed)
    n = cfg["n"]
    a = rand_state(n, seed)
    ref = run_oracle(a.copy(), circ)
    if cfg["world"] == 1:
        _, desc, l2p = run_emulator(a, n, circ, fuse=cfg["fuse"], tile_bits=cfg["tile_bits"],
                                    absorb_swaps=cfg["absorb_swaps"], reg_bits=cfg["reg_bits"])
    else:
        _, _, desc, l2p = run_emulator_sharded(a, n, cfg["world"], circ, fuse=cfg["fuse"], tile_bits=cfg["tile_bits"],
                                               absorb_swaps=cfg["absorb_swaps"], remap_pull=cfg["remap_pull"],
                                               reg_bits=cfg["reg_bits"])
    assert_close(unpermute(a, l2p), ref)


def _embed(m, qubits, n):
    """Dense 2^n x 2^n matrix of gate m on `qubits` (Quil order: first qubit = most significant matrix bit)."""
    k = len(qubits)
    full = np.zeros((1 << n, 1 << n), dtype=np.complex128)
    for col in range(1 << n):
        sub_c = 0
        for j, q in enumerate(qubits):
            sub_c |= ((col >> q) & 1) << (k - 1 - j)
        for sub_r in range(1 << k):
            row = col
            for j, q in enumerate(qubits):
                b = (sub_r >> (k - 1 - j)) & 1
                row = (row & ~(1 << q)) | (b << q)
            full[row, col] += m[sub_r, sub_c]
    return full


@pytest.mark.parametrize("seed", range(6))
def test_density_channels_on_vectorised_rho(seed):
    """DENSITY-QVM semantics (src/apply-gate.lisp:42-99): unitaries as conj(U) (x) U and Kraus lists as ONE superoperator
    on the 2n index bits of vec(rho) (row-major: index = row * 2^n + col), through the scheduler and the emulator,
    against rho' = sum_j K_j rho K_j^dagger computed with dense matrices."""
    from helpers import rand_unitary
    from qvm_b200 import qvm as Q
    rng = np.random.default_rng(300 + seed)
    n = int(rng.integers(3, 6))
    psi = rand_state(n, seed)
    rho = np.outer(psi, psi.conj())
    ops = []
    for _ in range(int(rng.integers(4, 14))):
        kind = int(rng.integers(0, 4))
        q = int(rng.integers(0, n))
        if kind == 0:
            ops.append((rand_unitary(1, rng), (q,)))
        elif kind == 1:
            ops.append((G.depolarizing_kraus_map(float(rng.uniform(0.01, 0.3))), (q,)))
        elif kind == 2:
            a, b = (int(x) for x in rng.choice(n, 2, replace=False))
            ops.append((G.gate_matrix("CNOT"), (a, b)))
        else:
            g = float(rng.uniform(0.05, 0.5))     # amplitude damping: non-unitary, non-diagonal Kraus pair
            ops.append(([np.array([[1, 0], [0, np.sqrt(1 - g)]]), np.array([[0, np.sqrt(g)], [0, 0]])], (q,)))
    ref = rho.copy()
    for gate, qubits in ops:
        ks = gate if isinstance(gate, (list, tuple)) else [gate]
        ref = sum(_embed(np.asarray(k, dtype=np.complex128), qubits, n) @ ref @ _embed(np.asarray(k, dtype=np.complex128), qubits, n).conj().T
                  for k in ks)
    vec = np.ascontiguousarray(rho.reshape(-1))
    gl = Q.density_gate_list(n, ops)
    run_emulator(vec, 2 * n, gl, fuse=True, tile_bits=int(rng.integers(4, 2 * n + 1)), reg_bits=int(rng.choice([3, 4])))
    assert_close(vec.reshape(1 << n, 1 << n), ref, rel=1e-12, abs_=1e-14)
    assert abs(np.trace(vec.reshape(1 << n, 1 << n)) - 1) < 1e-12
