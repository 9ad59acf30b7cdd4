"""ctypes binding of libqvmcuda (the C ABI in include/qvmcuda.h).

This is the Python twin of the CFFI binding in lisp/qvm-cuda.lisp: same entry
points, same argument conventions.  There is NO fallback: if the shared library
is missing or no CUDA device is present, loading / creating a state raises.
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("QVMCUDA_LIB") or os.path.join(_HERE, "libqvmcuda.so")   # QVMCUDA_LIB: A/B builds when profiling

FUSE = 1
ABSORB_SWAPS = 2


class QvmCudaError(RuntimeError):
    pass


_lib = None

_SIGS = {
    "qvmcuda_device_count": [C.POINTER(C.c_int)],
    "qvmcuda_launch_count": [C.POINTER(C.c_uint64)],
    "qvmcuda_state_create": [C.c_uint64, C.c_int, C.POINTER(C.c_void_p)],
    "qvmcuda_state_destroy": [C.c_void_p],
    "qvmcuda_state_length": [C.c_void_p, C.POINTER(C.c_uint64)],
    "qvmcuda_state_set_stream": [C.c_void_p, C.c_uint64],
    "qvmcuda_synchronize": [C.c_void_p],
    "qvmcuda_download": [C.c_void_p, C.c_void_p, C.c_uint64, C.c_uint64],
    "qvmcuda_upload": [C.c_void_p, C.c_void_p, C.c_uint64, C.c_uint64],
    "qvmcuda_set_zero_state": [C.c_void_p],
    "qvmcuda_set_basis_state": [C.c_void_p, C.c_uint64],
    "qvmcuda_copy": [C.c_void_p, C.c_void_p],
    "qvmcuda_apply_matrix": [C.c_void_p, C.c_int, C.c_void_p, C.c_void_p],
    "qvmcuda_apply_gates": [C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_uint32],
    "qvmcuda_tape_compile": [C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_uint32, C.POINTER(C.c_void_p)],
    "qvmcuda_tape_run": [C.c_void_p, C.c_void_p],
    "qvmcuda_tape_info": [C.c_void_p, C.c_void_p],
    "qvmcuda_tape_describe": [C.c_void_p, C.c_char_p, C.c_uint64],
    "qvmcuda_tape_destroy": [C.c_void_p],
    "qvmcuda_jit_stats": [C.c_void_p],
    "qvmcuda_tape_jit_source": [C.c_void_p, C.c_int, C.c_char_p, C.c_uint64, C.POINTER(C.c_uint64)],
    "qvmcuda_tape_jit_precompile": [C.c_void_p, C.POINTER(C.c_int), C.POINTER(C.c_int), C.c_char_p, C.c_uint64],
    "qvmcuda_prob_excited": [C.c_void_p, C.c_int, C.POINTER(C.c_double)],
    "qvmcuda_prob_ground": [C.c_void_p, C.c_int, C.POINTER(C.c_double)],
    "qvmcuda_norm2": [C.c_void_p, C.POINTER(C.c_double)],
    "qvmcuda_inner_product": [C.c_void_p, C.c_void_p, C.c_void_p],
    "qvmcuda_probabilities": [C.c_void_p, C.c_void_p, C.c_uint64, C.c_uint64],
    "qvmcuda_scale": [C.c_void_p, C.c_double],
    "qvmcuda_normalize": [C.c_void_p],
    "qvmcuda_collapse": [C.c_void_p, C.c_int, C.c_int, C.c_double],
    "qvmcuda_sample": [C.c_void_p, C.c_void_p, C.c_uint64, C.c_void_p, C.c_int],
    "qvmcuda_sample_total": [C.c_void_p, C.POINTER(C.c_double)],
    "qvmcuda_sample_shard": [C.c_void_p, C.c_void_p, C.c_uint64, C.c_void_p, C.c_int, C.c_double],
    "qvmcuda_density_apply_kraus": [C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_int, C.c_void_p, C.c_uint32],
    "qvmcuda_density_apply_ops": [C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_uint32],
    "qvmcuda_density_prob_excited": [C.c_void_p, C.c_int, C.c_int, C.POINTER(C.c_double)],
    "qvmcuda_density_collapse": [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_double],
    "qvmcuda_density_measure_discard": [C.c_void_p, C.c_int, C.c_int],
    "qvmcuda_density_diag_probs": [C.c_void_p, C.c_int, C.c_void_p],
    "qvmcuda_density_expectation": [C.c_void_p, C.c_int, C.c_void_p, C.c_void_p],
    "qvmcuda_set_identity_matrix": [C.c_void_p, C.c_int],
    "qvmcuda_shard_export": [C.c_void_p, C.c_void_p],
    "qvmcuda_shard_attach": [C.c_void_p, C.c_int, C.c_int, C.c_void_p],
    "qvmcuda_shard_export_alt": [C.c_void_p, C.c_void_p],
    "qvmcuda_shard_attach_alt": [C.c_void_p, C.c_void_p],
    "qvmcuda_shard_clear": [C.c_void_p],
    "qvmcuda_shard_set_zero_ranks": [C.c_void_p, C.c_uint32],
    "qvmcuda_shard_attach_local": [C.c_void_p, C.c_int, C.c_int],
    "qvmcuda_shard_compile": [C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_uint32, C.POINTER(C.c_void_p)],
    "qvmcuda_shard_plan": [C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_uint32,
                           C.POINTER(C.c_void_p)],
    "qvmcuda_tape_num_steps": [C.c_void_p, C.POINTER(C.c_int)],
    "qvmcuda_tape_step_flags": [C.c_void_p, C.c_int, C.POINTER(C.c_uint32)],
    "qvmcuda_tape_step_info": [C.c_void_p, C.c_int, C.c_void_p],
    "qvmcuda_tape_run_step": [C.c_void_p, C.c_void_p, C.c_int],
    "qvmcuda_tape_commit": [C.c_void_p, C.c_void_p],
    "qvmcuda_state_layout": [C.c_void_p, C.c_void_p, C.c_int],
}

EXPORTED_SYMBOLS = sorted(list(_SIGS) + ["qvmcuda_last_error"])


def lib():
    """Load libqvmcuda.so (built in-tree by __graft_entry__.build()).  Raises if it is missing."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise QvmCudaError(f"{LIB_PATH} not built: run `python -c 'import __graft_entry__ as g; g.build()'`")
        L = C.CDLL(LIB_PATH)
        L.qvmcuda_last_error.restype = C.c_char_p
        L.qvmcuda_last_error.argtypes = []
        for name, args in _SIGS.items():
            fn = getattr(L, name)
            fn.restype = C.c_int
            fn.argtypes = args
        _lib = L
    return _lib


def check(rc: int) -> None:
    if rc != 0:
        raise QvmCudaError(lib().qvmcuda_last_error().decode())


def ptr(a: np.ndarray):
    return a.ctypes.data_as(C.c_void_p)


def flatten_gates(gates):
    """[(matrix, qubits in Quil argument order)] -> (ks, qubits in NAT-TUPLE order, matrices as doubles)."""
    ks = np.ascontiguousarray([len(q) for _, q in gates], dtype=np.int32)
    qf = np.ascontiguousarray([x for _, q in gates for x in reversed(q)], dtype=np.int32)
    if len(gates):
        mf = np.concatenate([np.ascontiguousarray(m, dtype=np.complex128).ravel() for m, _ in gates]).view(np.float64)
    else:
        mf = np.zeros(0, dtype=np.float64)
    for (m, q), k in zip(gates, ks):
        if np.asarray(m).shape != (1 << int(k), 1 << int(k)):
            raise ValueError("gate matrix does not match its qubit count")
    return ks, qf, np.ascontiguousarray(mf)


def jit_stats() -> dict:
    """Counters of the pass compiler (qvmcuda_jit_stats)."""
    a = np.zeros(8, dtype=np.int64)
    check(lib().qvmcuda_jit_stats(ptr(a)))
    return {"compiled": int(a[0]), "cache_hits": int(a[1]), "disk_hits": int(a[2]), "failed": int(a[3]),
            "launches": int(a[4]), "compile_ms": int(a[5]), "policy": ["off", "sync", "async"][int(a[6])],
            "min_uops": int(a[7])}


def launch_count() -> int:
    n = C.c_uint64(0)
    check(lib().qvmcuda_launch_count(C.byref(n)))
    return int(n.value)


def device_count() -> int:
    n = C.c_int(0)
    check(lib().qvmcuda_device_count(C.byref(n)))
    return int(n.value)
