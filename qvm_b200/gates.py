"""Standard Quil gate matrices and gate modifiers (host side).

The reference obtains these from cl-quil (`quil:gate-matrix`, call sites
src/apply-gate.lisp:109-160) -- a third-party dependency that is not vendored in
the reference tree.  The definitions restated here are the ones the reference
ships in-tree as quil/stdgates.quil:1-206.  Matrix convention: row/column index
bit j (LSB = 0) belongs to Quil argument k-1-j, i.e. the FIRST argument is the
most significant bit (src/utilities.lisp:43-51).
"""
from __future__ import annotations

import cmath
import math
from typing import Callable, Dict, Sequence, Tuple

import numpy as np

_S2 = 1.0 / math.sqrt(2.0)


def _cis(x: float) -> complex:
    return complex(math.cos(x), math.sin(x))


def _m(rows) -> np.ndarray:
    return np.array(rows, dtype=np.complex128)


def _I(): return _m([[1, 0], [0, 1]])
def _X(): return _m([[0, 1], [1, 0]])
def _Y(): return _m([[0, -1j], [1j, 0]])
def _Z(): return _m([[1, 0], [0, -1]])
def _H(): return _m([[_S2, _S2], [_S2, -_S2]])
def _S(): return _m([[1, 0], [0, 1j]])
def _T(): return _m([[1, 0], [0, _cis(math.pi / 4)]])


def _RX(t):
    c, s = math.cos(t / 2), math.sin(t / 2)
    return _m([[c, -1j * s], [-1j * s, c]])


def _RY(t):
    c, s = math.cos(t / 2), math.sin(t / 2)
    return _m([[c, -s], [s, c]])


def _RZ(t): return _m([[_cis(-t / 2), 0], [0, _cis(t / 2)]])
def _PHASE(a): return _m([[1, 0], [0, _cis(a)]])


def _diag(*d): return np.diag(np.array(d, dtype=np.complex128))


def _CNOT(): return _m([[1, 0, 0, 0], [0, 1, 0, 0], [0, 0, 0, 1], [0, 0, 1, 0]])
def _CZ(): return _diag(1, 1, 1, -1)
def _SWAP(): return _m([[1, 0, 0, 0], [0, 0, 1, 0], [0, 1, 0, 0], [0, 0, 0, 1]])
def _ISWAP(): return _m([[1, 0, 0, 0], [0, 0, 1j, 0], [0, 1j, 0, 0], [0, 0, 0, 1]])
def _SQISWAP(): return _m([[1, 0, 0, 0], [0, _S2, 1j * _S2, 0], [0, 1j * _S2, _S2, 0], [0, 0, 0, 1]])
def _CPHASE00(a): return _diag(_cis(a), 1, 1, 1)
def _CPHASE01(a): return _diag(1, _cis(a), 1, 1)
def _CPHASE10(a): return _diag(1, 1, _cis(a), 1)
def _CPHASE(a): return _diag(1, 1, 1, _cis(a))
def _PSWAP(t): return _m([[1, 0, 0, 0], [0, 0, _cis(t), 0], [0, _cis(t), 0, 0], [0, 0, 0, 1]])


def _PISWAP(t):
    c, s = math.cos(t / 2), math.sin(t / 2)
    return _m([[1, 0, 0, 0], [0, c, 1j * s, 0], [0, 1j * s, c, 0], [0, 0, 0, 1]])


def _RZZ(p): return _diag(_cis(-p / 2), _cis(p / 2), _cis(p / 2), _cis(-p / 2))


def _RXX(p):
    c, s = math.cos(p / 2), math.sin(p / 2)
    return _m([[c, 0, 0, -1j * s], [0, c, -1j * s, 0], [0, -1j * s, c, 0], [-1j * s, 0, 0, c]])


def _RYY(p):
    c, s = math.cos(p / 2), math.sin(p / 2)
    return _m([[c, 0, 0, 1j * s], [0, c, -1j * s, 0], [0, -1j * s, c, 0], [1j * s, 0, 0, c]])


def _FSIM(t, p):
    c, s = math.cos(t / 2), math.sin(t / 2)
    return _m([[1, 0, 0, 0], [0, c, 1j * s, 0], [0, 1j * s, c, 0], [0, 0, 0, _cis(p)]])


def _perm(p: Sequence[int]) -> np.ndarray:
    n = len(p)
    m = np.zeros((n, n), dtype=np.complex128)
    for col, row in enumerate(p):
        m[row, col] = 1.0
    return m


def _CCNOT(): return _perm([0, 1, 2, 3, 4, 5, 7, 6])
def _CSWAP(): return _perm([0, 1, 2, 3, 4, 6, 5, 7])


# name -> (number of qubits, number of parameters, matrix function)
STANDARD_GATES: Dict[str, Tuple[int, int, Callable[..., np.ndarray]]] = {
    "I": (1, 0, _I), "X": (1, 0, _X), "Y": (1, 0, _Y), "Z": (1, 0, _Z), "H": (1, 0, _H),
    "S": (1, 0, _S), "T": (1, 0, _T),
    "RX": (1, 1, _RX), "RY": (1, 1, _RY), "RZ": (1, 1, _RZ), "PHASE": (1, 1, _PHASE),
    "CNOT": (2, 0, _CNOT), "CZ": (2, 0, _CZ), "SWAP": (2, 0, _SWAP), "ISWAP": (2, 0, _ISWAP),
    "SQISWAP": (2, 0, _SQISWAP),
    "CPHASE00": (2, 1, _CPHASE00), "CPHASE01": (2, 1, _CPHASE01), "CPHASE10": (2, 1, _CPHASE10),
    "CPHASE": (2, 1, _CPHASE), "PSWAP": (2, 1, _PSWAP), "PISWAP": (2, 1, _PISWAP), "XY": (2, 1, _PISWAP),
    "RZZ": (2, 1, _RZZ), "RXX": (2, 1, _RXX), "RYY": (2, 1, _RYY), "FSIM": (2, 2, _FSIM),
    "CCNOT": (3, 0, _CCNOT), "CSWAP": (3, 0, _CSWAP),
}


def gate_matrix(name: str, params: Sequence[float] = ()) -> np.ndarray:
    nq, npar, fn = STANDARD_GATES[name]
    if len(params) != npar:
        raise ValueError(f"gate {name} takes {npar} parameter(s), got {len(params)}")
    return fn(*[float(p) for p in params])


# ---- gate modifiers (tests/modifier-tests.lisp:7-130 pins their meaning) -----------------
def dagger(m: np.ndarray) -> np.ndarray:
    return np.ascontiguousarray(m.conj().T)


def controlled(m: np.ndarray) -> np.ndarray:
    """CONTROLLED G: the new (first) qubit is the MSB; identity when it is 0."""
    d = m.shape[0]
    out = np.eye(2 * d, dtype=np.complex128)
    out[d:, d:] = m
    return out


def forked(m0: np.ndarray, m1: np.ndarray) -> np.ndarray:
    """FORKED G(p0;p1): first qubit (MSB) selects G(p0) when 0, G(p1) when 1."""
    d = m0.shape[0]
    out = np.zeros((2 * d, 2 * d), dtype=np.complex128)
    out[:d, :d] = m0
    out[d:, d:] = m1
    return out


# ---- Kraus builders (src/basic-noise-qvm.lisp:251-278) ----------------------------------
def depolarizing_kraus_map(prob: float):
    """DEPOLARIZING-KRAUS-MAP src/basic-noise-qvm.lisp:251-258."""
    if not (0 < prob < 1):
        raise ValueError("DEPOLARIZATION-PROBABILITY must be between 0 and 1")
    pk0 = math.sqrt(1 - 0.75 * prob)
    pkn = math.sqrt(prob / 4)
    return [pk0 * _I(), pkn * _X(), pkn * _Y(), pkn * _Z()]


def damping_kraus_map(t1: float, elapsed: float):
    """DAMPING-KRAUS-MAP src/basic-noise-qvm.lisp:230-241 (column-major input)."""
    prob = 1 - math.exp(-elapsed / t1)
    k0 = _m([[0, math.sqrt(prob)], [0, 0]])
    k1 = _m([[1, 0], [0, math.sqrt(1 - prob)]])
    return [k0, k1]


def dephasing_kraus_map(t_phi: float, elapsed: float):
    """DEPHASING-KRAUS-MAP src/basic-noise-qvm.lisp:243-249."""
    prob = 1 - math.exp(-elapsed / t_phi)
    p0 = prob / 2
    p1 = 1 - p0
    return [math.sqrt(p0) * _I(), math.sqrt(p1) * _Z()]


def kraus_kron(k1s, k2s):
    """KRAUS-KRON src/basic-noise-qvm.lisp:260-269 (first list = MSB factor)."""
    ident = np.eye(2, dtype=np.complex128)
    if not k1s:
        return [np.kron(ident, k) for k in k2s]
    if not k2s:
        return [np.kron(k, ident) for k in k1s]
    return [np.kron(a, b) for a in k1s for b in k2s]


def check_kraus_ops(kraus, tol: float = 1e-5) -> None:
    """sum K^dagger K = I to 1e-5 (src/channel-qvm.lisp:109-133)."""
    d = kraus[0].shape[0]
    acc = np.zeros((d, d), dtype=np.complex128)
    for k in kraus:
        acc += k.conj().T @ k
    if not np.allclose(acc, np.eye(d), atol=tol):
        raise ValueError("Kraus operators do not satisfy the completeness relation")


__all__ = ["STANDARD_GATES", "gate_matrix", "dagger", "controlled", "forked", "depolarizing_kraus_map",
           "damping_kraus_map", "dephasing_kraus_map", "kraus_kron", "check_kraus_ops", "cmath"]
