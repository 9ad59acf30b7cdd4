"""Circuit generators for the benchmark configurations (host side).

A circuit here is a list of (matrix, qubits) with QUBITS in Quil argument order
(first = most significant matrix-index bit), the form `apply-gate-to-state`
receives (src/apply-gate.lisp:106-160).
"""
from __future__ import annotations

import math
from typing import List, Sequence, Tuple

import numpy as np

from . import gates as G

Circuit = List[Tuple[np.ndarray, Tuple[int, ...]]]


def qft_circuit(qubits: Sequence[int]) -> Circuit:
    """QFT-CIRCUIT examples/qft.lisp:23-62 (+ BIT-REVERSAL-CIRCUIT :9-21).

    qft(q . qs) = qft(qs) ++ [CPHASE(pi/2^(n-i)) q qi ...](pushed, hence reversed) ++ [H q];
    then floor(n/2) SWAPs qubits[i] <-> qubits[n-1-i]."""
    qubits = list(qubits)

    def qft(qs: List[int]) -> Circuit:
        q, rest = qs[0], qs[1:]
        if not rest:
            return [(G.gate_matrix("H"), (q,))]
        n = 1 + len(rest)
        cr: Circuit = []
        for i, qi in zip(range(n - 1, 0, -1), rest):
            angle = math.pi / (2 ** (n - i))
            cr.insert(0, (G.gate_matrix("CPHASE", [angle]), (q, qi)))
        return qft(rest) + cr + [(G.gate_matrix("H"), (q,))]

    out = qft(qubits)
    n = len(qubits)
    if n >= 2:
        for i in range(n // 2):
            out.append((G.gate_matrix("SWAP"), (qubits[i], qubits[n - 1 - i])))
    return out


def hadamard_circuit(n: int) -> Circuit:
    """bench/20H.quil, bench/25H.quil and the app's `hadamard` benchmark
    (app/src/benchmark-programs.lisp:7-138): H on every qubit."""
    return [(G.gate_matrix("H"), (q,)) for q in range(n)]


def bell_circuit(n: int) -> Circuit:
    """tests/gate-tests.lisp:75-89 / the app's `bell` benchmark: H 0; CNOT 0 i."""
    out: Circuit = [(G.gate_matrix("H"), (0,))]
    for i in range(1, n):
        out.append((G.gate_matrix("CNOT"), (0, i)))
    return out


def random_layer_circuit(n: int, layers: int, seed: int) -> Circuit:
    """per layer RZ.RY.RZ with seeded angles on every qubit,
    then CZ on a seeded random perfect matching."""
    rng = np.random.default_rng(seed)
    out: Circuit = []
    for _ in range(layers):
        ang = rng.uniform(0.0, 2.0 * math.pi, size=(n, 3))
        for q in range(n):
            out.append((G.gate_matrix("RZ", [ang[q, 0]]), (q,)))
            out.append((G.gate_matrix("RY", [ang[q, 1]]), (q,)))
            out.append((G.gate_matrix("RZ", [ang[q, 2]]), (q,)))
        perm = rng.permutation(n)
        for i in range(0, n - 1, 2):
            out.append((G.gate_matrix("CZ"), (int(perm[i]), int(perm[i + 1]))))
    return out


def qaoa_maxcut_circuit(n: int, edges: Sequence[Tuple[int, int]], gamma: float = 0.7, beta: float = 0.3) -> Circuit:
    """p=1 MAXCUT QAOA (examples/qaoa.lisp:47-66) expanded by the fixed rule of SURVEY.md
    section 8(d) C4: H on all, per edge CNOT.RZ(2 gamma).CNOT, RX(2 beta) on all."""
    out: Circuit = [(G.gate_matrix("H"), (q,)) for q in range(n)]
    for a, b in edges:
        out.append((G.gate_matrix("CNOT"), (a, b)))
        out.append((G.gate_matrix("RZ", [2.0 * gamma]), (b,)))
        out.append((G.gate_matrix("CNOT"), (a, b)))
    for q in range(n):
        out.append((G.gate_matrix("RX", [2.0 * beta]), (q,)))
    return out


def line_graph(n: int): return [(i, i + 1) for i in range(n - 1)]
def complete_graph(n: int): return [(i, j) for i in range(n) for j in range(i + 1, n)]
