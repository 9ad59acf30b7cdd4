"""The reference's gate-modifier known answers (tests/modifier-tests.lisp:7-130: CONTROLLED, DAGGER, FORKED and their
nestings) and the compiled == interpreted program of tests/gate-tests.lisp:28-44, run twice on the CPU: through the
oracle (one APPLY-MATRIX-OPERATOR per instruction) and through the scheduler + the emulator of the tile programs
(what the GPU executes).  cflonum= in the reference is a 1e-5-ish comparison (tests/utilities.lisp); we hold 1e-12."""
import numpy as np
import pytest

from helpers import assert_close, run_emulator
from oracle import oracle as O
from qvm_b200.quil import parse_quil


def _circuit(quil):
    p = parse_quil(quil)
    return [(p.gate_matrix(i), tuple(i.qubits)) for i in p.instructions if type(i).__name__ == "GateApp"]


def _both(n, quil):
    circ = _circuit(quil)
    a = O.zero_state(n)
    for m, q in circ:
        O.apply_matrix(a, m, q)
    outs = [a]
    for fuse in (True, False):
        for reg_bits in (3, 4):
            b = O.zero_state(n)
            run_emulator(b, n, circ, fuse=fuse, reg_bits=reg_bits)
            assert_close(b, a)
            outs.append(b)
    return a


def test_controlled_x_is_cnot():      # modifier-tests.lisp:7-21
    assert_close(_both(2, "H 0\nCONTROLLED X 0 1"), _both(2, "H 0\nCNOT 0 1"))


def test_controlled_controlled_x_is_ccnot():      # :23-41
    assert_close(_both(3, "H 0\nH 1\nH 2\nCONTROLLED CONTROLLED X 0 1 2"), _both(3, "H 0\nH 1\nH 2\nCCNOT 0 1 2"))


def test_dagger_inversion():      # :44-63
    a = _both(3, """H 0
CONTROLLED RY(2*pi/3) 0 1
RX(pi/3) 0
T 2
CSWAP 0 1 2
DAGGER CSWAP 0 1 2
DAGGER T 2
DAGGER RX(pi/3) 0
DAGGER CONTROLLED RY(2*pi/3) 0 1
DAGGER H 0""")
    assert abs(a[0] - 1) < 1e-12 and np.abs(a[1:]).max() < 1e-12


def test_forked_rx():      # :65-80
    assert np.allclose(np.abs(_both(2, "FORKED RX(0, pi) 0 1")) ** 2, [1, 0, 0, 0], atol=1e-12)
    assert np.allclose(np.abs(_both(2, "X 0\nFORKED RX(0, pi) 0 1")) ** 2, [0, 0, 0, 1], atol=1e-12)


def test_forked_rx_with_controlled_and_dagger():      # :82-96
    assert abs(_both(3, "CONTROLLED DAGGER FORKED RX(0, pi) 0 1 2")[7]) ** 2 < 1e-12
    assert abs(abs(_both(3, "X 0\nX 1\nCONTROLLED DAGGER FORKED RX(0, pi) 0 1 2")[7]) ** 2 - 1) < 1e-12


@pytest.mark.parametrize("angles,first,second", [("0, 0, 0, pi", 0, 7), ("pi, 0, 0, pi", 4, 7)])
def test_multiply_forked_rx(angles, first, second):      # :98-130 (the uniformly controlled rotation)
    p1 = np.abs(_both(3, f"FORKED FORKED RX({angles}) 0 1 2")) ** 2
    p2 = np.abs(_both(3, f"X 0\nX 1\nFORKED FORKED RX({angles}) 0 1 2")) ** 2
    e1, e2 = np.zeros(8), np.zeros(8)
    e1[first] = 1
    e2[second] = 1
    assert np.allclose(p1, e1, atol=1e-12) and np.allclose(p2, e2, atol=1e-12)


def test_compiled_program_of_gate_tests():      # tests/gate-tests.lisp:28-44: H0 H1 H2 H3; CNOT 2 0; CSWAP 1 3 2
    a = _both(4, "H 0\nH 1\nH 2\nH 3\nCNOT 2 0\nCSWAP 1 3 2")
    assert np.allclose(np.abs(a) ** 2, 1 / 16, atol=1e-14)      # permutations of the uniform superposition
