"""ctypes binding of baseline/libcpuport.so -- the CPU BASELINE timed by bench.py (not the parity checker)."""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_lib = None


def lib():
    global _lib
    if _lib is None:
        so = os.path.join(_HERE, "libcpuport.so")
        if not os.path.exists(so):
            subprocess.check_call(["make", "-C", _HERE, "-s"])
        _lib = C.CDLL(so)
        _lib.cpb_all_cores.restype = C.c_int
    return _lib


def all_cores() -> int:
    """Logical CPUs of the machine; deliberately ignores OMP_NUM_THREADS (torchrun sets it to 1)."""
    return int(lib().cpb_all_cores())


def zero_state(n: int, threads: int) -> np.ndarray:
    psi = np.empty(1 << n, dtype=np.complex128)
    lib().cpb_zero_state(psi.ctypes.data_as(C.c_void_p), n, threads)
    return psi


def apply_gate(psi: np.ndarray, matrix, qubits, threads: int) -> None:
    """qubits in Quil argument order (first = most significant matrix index bit), like the oracle's apply_matrix."""
    n = int(psi.size).bit_length() - 1
    m = np.ascontiguousarray(matrix, dtype=np.complex128)
    q = np.ascontiguousarray(list(reversed(qubits)), dtype=np.int32)
    lib().cpb_apply_gate(psi.ctypes.data_as(C.c_void_p), n, len(qubits), q.ctypes.data_as(C.c_void_p),
                         m.ctypes.data_as(C.c_void_p), int(threads))
