"""oracle/oracle.py -- TEST INFRASTRUCTURE ONLY (see oracle/qvm_oracle.c header).

ctypes wrapper around libqvmoracle.so, the CPU restatement of the reference's
hot path.  Importable only from tests/, __graft_entry__.smoke() and bench.py's
cpu_baseline / --impl reference legs.  The product (qvm_b200/) never imports it.

Conventions: states are numpy complex128 vectors (interleaved re/im doubles,
src/floats.lisp:13-26).  `qubits` arguments here are in Quil ARGUMENT order
(first = MSB of the matrix index), exactly what `apply-gate-to-state` receives
(src/apply-gate.lisp:106-160); the wrapper reverses them into NAT-TUPLE order
(src/utilities.lisp:43-51) before calling C.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "libqvmoracle.so")


def build(force: bool = False) -> str:
    src = os.path.join(_HERE, "qvm_oracle.c")
    if force or not os.path.exists(_SO) or os.path.getmtime(_SO) < os.path.getmtime(src):
        subprocess.check_call(["make", "-C", _HERE, "-s"])
    return _SO


_lib = None


def lib():
    global _lib
    if _lib is None:
        build()
        L = C.CDLL(_SO)
        u64, i32, dbl, vp = C.c_uint64, C.c_int, C.c_double, C.c_void_p
        L.orc_inject_bit.restype = u64; L.orc_inject_bit.argtypes = [u64, i32]
        L.orc_eject_bit.restype = u64; L.orc_eject_bit.argtypes = [u64, i32]
        L.orc_index_to_address.restype = u64; L.orc_index_to_address.argtypes = [u64, i32, i32]
        L.orc_apply_matrix.restype = i32; L.orc_apply_matrix.argtypes = [vp, u64, i32, vp, vp]
        L.orc_apply_matrix_mt.restype = i32; L.orc_apply_matrix_mt.argtypes = [vp, u64, i32, vp, vp, i32]
        L.orc_max_threads.restype = i32
        L.orc_apply_permutation.restype = i32; L.orc_apply_permutation.argtypes = [vp, u64, i32, vp, vp]
        for f in ("orc_prob_excited", "orc_prob_ground"):
            getattr(L, f).restype = dbl; getattr(L, f).argtypes = [vp, u64, i32]
        for f in ("orc_norm2", "orc_norm"):
            getattr(L, f).restype = dbl; getattr(L, f).argtypes = [vp, u64]
        L.orc_probabilities.restype = None; L.orc_probabilities.argtypes = [vp, u64, vp]
        L.orc_inner_product.restype = None; L.orc_inner_product.argtypes = [vp, vp, u64, vp]
        L.orc_scale.restype = None; L.orc_scale.argtypes = [vp, u64, dbl]
        L.orc_normalize.restype = None; L.orc_normalize.argtypes = [vp, u64]
        L.orc_force_measurement.restype = None; L.orc_force_measurement.argtypes = [vp, u64, i32, i32, dbl]
        L.orc_measure.restype = i32; L.orc_measure.argtypes = [vp, u64, i32, dbl]
        L.orc_measure_compiled.restype = i32; L.orc_measure_compiled.argtypes = [vp, u64, i32, dbl]
        L.orc_cdf.restype = None; L.orc_cdf.argtypes = [vp, u64, vp]
        L.orc_sample_multiple.restype = None; L.orc_sample_multiple.argtypes = [vp, u64, vp, u64, vp]
        L.orc_sample_bisect.restype = u64; L.orc_sample_bisect.argtypes = [vp, u64, dbl]
        L.orc_sample_as_distribution.restype = None; L.orc_sample_as_distribution.argtypes = [vp, u64, vp, u64, vp]
        L.orc_measure_all.restype = u64; L.orc_measure_all.argtypes = [vp, u64, dbl]
        L.orc_zero_state.restype = None; L.orc_zero_state.argtypes = [vp, u64]
        L.orc_density_apply_unitary.restype = i32; L.orc_density_apply_unitary.argtypes = [vp, i32, i32, vp, vp]
        L.orc_density_apply_kraus.restype = i32; L.orc_density_apply_kraus.argtypes = [vp, i32, i32, vp, i32, vp]
        L.orc_density_prob_excited.restype = dbl; L.orc_density_prob_excited.argtypes = [vp, i32, i32]
        L.orc_density_force_measurement.restype = None; L.orc_density_force_measurement.argtypes = [vp, i32, i32, i32, dbl]
        L.orc_density_measure_discard.restype = None; L.orc_density_measure_discard.argtypes = [vp, i32, i32]
        L.orc_density_diag_probs.restype = None; L.orc_density_diag_probs.argtypes = [vp, i32, vp]
        L.orc_evolve_stochastic.restype = i32; L.orc_evolve_stochastic.argtypes = [vp, u64, i32, vp, i32, vp, dbl]
        L.orc_sample_tree.restype = None; L.orc_sample_tree.argtypes = [vp, u64, vp, u64, i32, vp]
        L.orc_sample_tree_total.restype = dbl; L.orc_sample_tree_total.argtypes = [vp, u64]
        L.orc_sample_tree_base.restype = None; L.orc_sample_tree_base.argtypes = [vp, u64, vp, u64, i32, dbl, vp]
        L.orc_sample_tree_sharded.restype = None; L.orc_sample_tree_sharded.argtypes = [vp, u64, i32, vp, u64, i32, vp]
        _lib = L
    return _lib


def _p(a: np.ndarray):
    return a.ctypes.data_as(C.c_void_p)


def _state(psi: np.ndarray) -> np.ndarray:
    assert psi.dtype == np.complex128 and psi.flags.c_contiguous and psi.ndim == 1
    return psi


def _nt(qubits) -> np.ndarray:
    """Quil argument order -> NAT-TUPLE order (reversed), int32."""
    return np.ascontiguousarray(list(reversed([int(q) for q in qubits])), dtype=np.int32)


def _mat(m, k) -> np.ndarray:
    m = np.ascontiguousarray(np.asarray(m, dtype=np.complex128))
    assert m.shape == (1 << k, 1 << k), (m.shape, k)
    return m


# ---------------------------------------------------------------- bit math
def inject_bit(x, n): return int(lib().orc_inject_bit(x, n))
def eject_bit(x, n): return int(lib().orc_eject_bit(x, n))
def index_to_address(index, qubit, state): return int(lib().orc_index_to_address(index, qubit, state))


# ---------------------------------------------------------------- pure state
def zero_state(n: int) -> np.ndarray:
    psi = np.zeros(1 << n, dtype=np.complex128)
    psi[0] = 1.0
    return psi


def apply_matrix(psi, matrix, qubits, threads: int = 0):
    """qvm:apply-matrix-operator with QUBITS in Quil argument order."""
    _state(psi)
    q = _nt(qubits)
    m = _mat(matrix, len(q))
    if threads and threads > 1:
        rc = lib().orc_apply_matrix_mt(_p(psi), psi.size, len(q), _p(q), _p(m), threads)
    else:
        rc = lib().orc_apply_matrix(_p(psi), psi.size, len(q), _p(q), _p(m))
    assert rc == 0
    return psi


def apply_permutation(psi, perm, qubits):
    q = _nt(qubits)
    pm = np.ascontiguousarray(perm, dtype=np.int32)
    assert pm.size == 1 << len(q)
    assert lib().orc_apply_permutation(_p(_state(psi)), psi.size, len(q), _p(q), _p(pm)) == 0
    return psi


def prob_excited(psi, q): return float(lib().orc_prob_excited(_p(_state(psi)), psi.size, q))
def prob_ground(psi, q): return float(lib().orc_prob_ground(_p(_state(psi)), psi.size, q))
def norm2(psi): return float(lib().orc_norm2(_p(_state(psi)), psi.size))
def norm(psi): return float(lib().orc_norm(_p(_state(psi)), psi.size))


def probabilities(psi) -> np.ndarray:
    """|psi_i|^2 for every basis state (PROBABILITY, src/wavefunction.lisp:44-50)."""
    out = np.zeros(psi.size, dtype=np.float64)
    lib().orc_probabilities(_p(_state(psi)), psi.size, _p(out))
    return out


def inner_product(a, b) -> complex:
    """<a|b> as PURE-STATE-EXPECTATION sums it (app/src/api/expectation.lisp:79-84)."""
    out = np.zeros(2, dtype=np.float64)
    lib().orc_inner_product(_p(_state(a)), _p(_state(b)), a.size, _p(out))
    return complex(out[0], out[1])
def scale(psi, a): lib().orc_scale(_p(_state(psi)), psi.size, a); return psi
def normalize(psi): lib().orc_normalize(_p(_state(psi)), psi.size); return psi


def force_measurement(psi, q, value, p1):
    lib().orc_force_measurement(_p(_state(psi)), psi.size, q, value, p1)
    return psi


def measure(psi, q, r): return int(lib().orc_measure(_p(_state(psi)), psi.size, q, r))
def measure_compiled(psi, q, r): return int(lib().orc_measure_compiled(_p(_state(psi)), psi.size, q, r))


def cdf(psi):
    out = np.empty(psi.size, dtype=np.float64)
    lib().orc_cdf(_p(_state(psi)), psi.size, _p(out))
    return out


def sample_multiple(psi, u):
    u = np.ascontiguousarray(u, dtype=np.float64)
    out = np.empty(u.size, dtype=np.uint64)
    lib().orc_sample_multiple(_p(_state(psi)), psi.size, _p(u), u.size, _p(out))
    return out


def sample_bisect(psi, p): return int(lib().orc_sample_bisect(_p(_state(psi)), psi.size, p))


def sample_as_distribution(psi, ps):
    ps = np.ascontiguousarray(ps, dtype=np.float64)
    out = np.empty(ps.size, dtype=np.uint64)
    lib().orc_sample_as_distribution(_p(_state(psi)), psi.size, _p(ps), ps.size, _p(out))
    return out


def sample_tree(psi, u, strict: bool):
    u = np.ascontiguousarray(u, dtype=np.float64)
    out = np.empty(u.size, dtype=np.uint64)
    lib().orc_sample_tree(_p(_state(psi)), psi.size, _p(u), u.size, 1 if strict else 0, _p(out))
    return out


def sample_tree_total(psi) -> float:
    return float(lib().orc_sample_tree_total(_p(_state(psi)), psi.size))


def sample_tree_base(psi, u, strict: bool, base: float):
    u = np.ascontiguousarray(u, dtype=np.float64)
    out = np.empty(u.size, dtype=np.uint64)
    lib().orc_sample_tree_base(_p(_state(psi)), psi.size, _p(u), u.size, 1 if strict else 0, float(base), _p(out))
    return out


def sample_tree_sharded(psi, world: int, u, strict: bool):
    """The sharded sampler's order: psi = `world` contiguous shards (physical index order, rank-major)."""
    u = np.ascontiguousarray(u, dtype=np.float64)
    out = np.empty(u.size, dtype=np.uint64)
    lib().orc_sample_tree_sharded(_p(_state(psi)), psi.size, int(world), _p(u), u.size, 1 if strict else 0, _p(out))
    return out


def measure_all(psi, p): return int(lib().orc_measure_all(_p(_state(psi)), psi.size, p))


# ---------------------------------------------------------------- density
def zero_density(n: int) -> np.ndarray:
    rho = np.zeros(1 << (2 * n), dtype=np.complex128)
    rho[0] = 1.0
    return rho


def density_apply_unitary(rho, n, U, qubits):
    q = _nt(qubits)
    assert lib().orc_density_apply_unitary(_p(_state(rho)), n, len(q), _p(q), _p(_mat(U, len(q)))) == 0
    return rho


def density_apply_kraus(rho, n, kraus, qubits):
    q = _nt(qubits)
    ks = np.ascontiguousarray(np.stack([_mat(k, len(q)) for k in kraus]))
    assert lib().orc_density_apply_kraus(_p(_state(rho)), n, len(q), _p(q), len(kraus), _p(ks)) == 0
    return rho


def density_prob_excited(rho, n, q): return float(lib().orc_density_prob_excited(_p(_state(rho)), n, q))


def density_force_measurement(rho, n, q, value, p1):
    lib().orc_density_force_measurement(_p(_state(rho)), n, q, value, p1)
    return rho


def density_measure_discard(rho, n, q):
    lib().orc_density_measure_discard(_p(_state(rho)), n, q)
    return rho


def density_diag_probs(rho, n):
    out = np.empty(1 << n, dtype=np.float64)
    lib().orc_density_diag_probs(_p(_state(rho)), n, _p(out))
    return out


def evolve_stochastic(psi, kraus, qubits, r):
    """%EVOLVE-PURE-STATE-STOCHASTICALLY (src/apply-gate.lisp:16-39) with draw r; returns the operator index."""
    q = _nt(qubits)
    ks = np.ascontiguousarray(np.stack([_mat(k, len(q)) for k in kraus]))
    return int(lib().orc_evolve_stochastic(_p(_state(psi)), psi.size, len(q), _p(q), len(kraus), _p(ks), r))


def max_threads() -> int:
    return int(lib().orc_max_threads())


def density_measure_all(rho, n, uniforms):
    """NAIVE-MEASURE-ALL on a density matrix (src/measurement.lisp:153-162): MEASURE q for q = n-1 .. 0, each with the
    decision rule of src/measurement.lisp:93-105 (bit = 0 if p1 = 0 -- WITHOUT drawing; else 1 if (random 1d0) <= p1; else 0)
    and the collapse of :43-68, fed the given uniforms in that order.  Returns the bits, element i = qubit i."""
    bits = [0] * n
    it = iter(uniforms)
    for q in range(n - 1, -1, -1):
        p1 = density_prob_excited(rho, n, q)
        bit = 0 if p1 == 0.0 else (1 if next(it) <= p1 else 0)      # COND evaluates (random 1d0) only when p1 /= 0
        density_force_measurement(rho, n, q, bit, p1)
        bits[q] = bit
    return bits
