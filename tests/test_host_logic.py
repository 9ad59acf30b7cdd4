"""CPU-only checks of the host side: the C-ABI library loads and exports every declared symbol (no
compute without a GPU), the Quil reader, the circuit generators, gate modifiers."""
import math
import os
import re

import numpy as np
import pytest

from qvm_b200 import circuits as CC
from qvm_b200 import gates as G
from qvm_b200.quil import evaluate, parse_quil

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_cabi_exports_every_declared_symbol():
    from qvm_b200 import _lib
    lib = _lib.lib()
    header = open(os.path.join(ROOT, "include", "qvmcuda.h")).read()
    declared = sorted(set(re.findall(r"\b(qvmcuda_[a-z0-9_]+)\s*\(", header)))
    assert declared == _lib.EXPORTED_SYMBOLS
    for name in declared:
        assert getattr(lib, name) is not None


def test_no_cpu_fallback_without_device():
    import ctypes as C
    from qvm_b200 import _lib
    try:
        import torch
        if torch.cuda.is_available():
            pytest.skip("GPU present")
    except ImportError:
        pass
    h = C.c_void_p()
    rc = _lib.lib().qvmcuda_state_create(1 << 4, 0, C.byref(h))
    assert rc != 0 and _lib.lib().qvmcuda_last_error()


def test_tape_compiles_without_a_device():
    from qvm_b200 import qvm
    t = qvm.Tape(30, CC.qft_circuit(range(30)), fuse=True)
    info = t.info()
    assert info["gates"] == 480 and info["passes"] <= 10
    t2 = qvm.Tape(30, CC.qft_circuit(range(30)), fuse=False)
    assert t2.info()["passes"] == 480


def test_expression_evaluator():
    assert abs(evaluate("pi/2") - math.pi / 2) < 1e-15
    assert abs(evaluate("1/sqrt(2)") - 2 ** -0.5) < 1e-15
    assert abs(evaluate("-i*sin(%t/2)", {"t": 1.0}) - (-1j * math.sin(0.5))) < 1e-15
    assert abs(evaluate("cis(-%t/2)", {"t": 0.3}) - complex(math.cos(0.15), -math.sin(0.15))) < 1e-15
    assert evaluate("0.53968091093885451+0.0i") == complex(0.5396809109388545, 0.0)
    assert evaluate("0.0-0.22035744822848241i") == complex(0.0, -0.22035744822848241)
    assert evaluate("2^3") == 8


def test_quil_reader():
    p = parse_quil("""
DECLARE ro BIT[2]
DEFGATE G(%a):
    cos(%a), sin(%a)
    -sin(%a), cos(%a)
DEFGATE P AS PERMUTATION:
    0, 1, 3, 2
H 0
CPHASE(3.141592653589793) 13 14
G(0.0) 0
P 1 0
CONTROLLED DAGGER RX(pi/2) 2 0
FORKED RX(0.1, 0.2) 1 0
MEASURE 0 ro[0]
MEASURE 1
RESET
""")
    kinds = [type(i).__name__ for i in p.instructions]
    assert kinds == ["Declare", "GateApp", "GateApp", "GateApp", "GateApp", "GateApp", "GateApp", "Measure", "Measure", "Reset"]
    assert p.qubits_needed() == 15
    gs = [i for i in p.instructions if type(i).__name__ == "GateApp"]
    assert np.allclose(p.gate_matrix(gs[2]), np.eye(2))
    assert np.allclose(p.gate_matrix(gs[3]), G.gate_matrix("CNOT"))
    m = p.gate_matrix(gs[4])
    assert m.shape == (4, 4) and np.allclose(m[:2, :2], np.eye(2)) and np.allclose(m[2:, 2:], G.dagger(G.gate_matrix("RX", [math.pi / 2])))
    f = p.gate_matrix(gs[5])
    assert np.allclose(f[:2, :2], G.gate_matrix("RX", [0.1])) and np.allclose(f[2:, 2:], G.gate_matrix("RX", [0.2]))
    assert p.instructions[7].target == ("ro", 0) and p.instructions[8].target is None


def test_standard_gates_are_unitary_and_match_stdgates():
    rng = np.random.default_rng(0)
    for name, (nq, npar, fn) in G.STANDARD_GATES.items():
        m = G.gate_matrix(name, list(rng.uniform(0, 6, npar)))
        assert m.shape == (1 << nq, 1 << nq)
        assert np.allclose(m.conj().T @ m, np.eye(1 << nq), atol=1e-14), name
    assert np.allclose(G.gate_matrix("CPHASE", [math.pi]), G.gate_matrix("CZ"))
    assert np.allclose(G.gate_matrix("XY", [0.7]), G.gate_matrix("PISWAP", [0.7]))   # tests/gate-tests.lisp:115-127


def test_qft_circuit_shape():
    c = CC.qft_circuit(range(30))
    assert len(c) == 480
    assert c[0][1] == (29,) and c[1][1] == (28, 29) and c[2][1] == (28,)     # SURVEY appendix C order
    assert abs(np.angle(c[3][0][3, 3]) - math.pi / 4) < 1e-15 and c[3][1] == (27, 29)
    assert c[-15][1] == (0, 29) and c[-1][1] == (14, 15)
    assert CC.qft_circuit([5]) == [] or len(CC.qft_circuit([5])) == 1


def test_multishot_relabeling_and_bit_extraction():
    """app/src/api/multishot-measure.lisp:13-38,62-77: relabeled qubits, :unused-qubit reads 0, bits in request order;
    the sampled basis states come from the oracle's MEASURE-ALL sampler here (the GPU sampler is pinned to it bit for bit)."""
    from oracle import oracle as O
    from qvm_b200 import qvm as Q
    qs, n = Q.relabel_multishot_qubits([0, 2, 5], 2, [2, 0])
    assert qs == [1, 0, Q.UNUSED_QUBIT] and n == 2
    qs, n = Q.relabel_multishot_qubits([3], 1, [9, 8, 7, 3])
    assert qs == [3] and n == 4
    assert Q.relabel_multishot_qubits([1, 0], 2, None) == ([1, 0], 2)
    with pytest.raises(ValueError):
        Q.relabel_multishot_qubits([-1], 2, None)
    # GHZ on 3 qubits: every trial reads 000 or 111 whatever the request order; unused qubits read 0
    psi = np.zeros(8, dtype=np.complex128)
    psi[0] = psi[7] = math.sqrt(0.5)
    u = np.random.default_rng(1).random(200)
    states = O.sample_tree(psi, u, True)
    bits = Q.multishot_bits(states, [2, Q.UNUSED_QUBIT, 0, 1])
    assert len(bits) == 200 and all(b in ([0, 0, 0, 0], [1, 0, 1, 1]) for b in bits)
    assert 60 < sum(b[0] for b in bits) < 140
    assert Q.multishot_bits([5, 2], [0, 1, 2]) == [[1, 0, 1], [0, 1, 0]]
    assert Q.perform_multishot_measure("H 0", 1, [], 10) == [] and Q.perform_multishot_measure("H 0", 1, [0], 0) == []


def test_header_is_plain_c(tmp_path):
    """The boundary is a C ABI: include/qvmcuda.h must compile as C99 (no C++ / torch types in the signatures)."""
    import subprocess
    src = tmp_path / "hdr.c"
    src.write_text('#include "qvmcuda.h"\nint main(void) { return 0; }\n')
    subprocess.check_call(["gcc", "-std=c99", "-Wall", "-Wextra", "-pedantic", "-fsyntax-only",
                           "-I", os.path.join(ROOT, "include"), str(src)])
