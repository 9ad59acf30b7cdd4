"""The reference's own measurement / gate test programs (tests/measurement-tests.lisp, tests/gate-tests.lisp), run through the
Python twin of the QVM API on the GPU in both execution modes the reference's suite uses (:interpret and :compile,
tests/suite.lisp:9-16).  Same programs, same assertions; the statistical tests use fewer repetitions than the reference's 10^5
(the tolerance of 0.05 is the reference's)."""
import math

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def Q():
    from qvm_b200 import qvm
    return qvm


def _modes(Q, fn, modes=(False, True)):
    for compiled in modes:
        Q.compile_before_running = compiled
        try:
            fn()
        finally:
            Q.compile_before_running = False


def _needed(prog):
    from qvm_b200.quil import parse_quil
    p = parse_quil(prog)
    return 1 + max(q for ins in p.instructions for q in ([getattr(ins, "qubit", None)] if getattr(ins, "qubit", None) is not None
                                                        else list(getattr(ins, "qubits", ()))))


@pytest.mark.parametrize("prog,want,modes", [
    # test-simple-measurements :22-35
    ("DECLARE ro BIT[3]\nX 0\nX 2\nMEASURE 0 ro[0]\nMEASURE 1 ro[1]\nMEASURE 2 ro[2]", (1, 0, 1), (False, True)),
    # test-simple-cheeky-measurements :37-51 (a fourth qubit keeps the chain from becoming a MEASURE-ALL)
    ("DECLARE ro BIT[3]\nX 0\nX 2\nX 3\nMEASURE 0 ro[0]\nMEASURE 1 ro[1]\nMEASURE 2 ro[2]", (1, 0, 1), (True,)),
    # test-simple-measurements-with-swap :53-68
    ("DECLARE ro BIT[3]\nX 0\nX 2\nSWAP 0 1\nSWAP 0 2\nMEASURE 0 ro[0]\nMEASURE 1 ro[1]\nMEASURE 2 ro[2]", (1, 1, 0), (False, True)),
    # test-simple-cheeky-measurements-with-swap :70-86
    ("DECLARE ro BIT[3]\nX 0\nX 2\nX 3\nSWAP 0 1\nSWAP 0 2\nMEASURE 0 ro[0]\nMEASURE 1 ro[1]\nMEASURE 2 ro[2]", (1, 1, 0), (True,)),
])
def test_simple_measurements(Q, prog, want, modes):
    def run():
        m = Q.run_program(_needed(prog), prog)
        assert tuple(int(b) for b in m.registers["ro"][:3]) == want
    _modes(Q, run, modes)


@pytest.mark.parametrize("cheeky", [False, True])
def test_unit_wavefunction_after_measurements(Q, cheeky):
    """:88-110: H q; MEASURE q for every qubit, 5 .. 15 qubits: the wavefunction keeps unit length."""
    def run():
        for n in range(5, 16):
            prog = (f"X {n}\n" if cheeky else "") + "".join(f"H {q}\nMEASURE {q}\n" for q in range(n))
            m = Q.run_program(n + (1 if cheeky else 0), prog, seed=n)
            assert abs(math.sqrt(m.state.vec.norm2()) - 1) < 1e-12
    _modes(Q, run, (True,) if cheeky else (False, True))


@pytest.mark.parametrize("gate,p_one", [("H 0", 0.5), ("RX(pi/4) 0", math.sin(math.pi / 8) ** 2),
                                        ("RX(3*pi/4) 0", 1 - math.sin(math.pi / 8) ** 2), ("RX(pi/8) 0", math.sin(math.pi / 16) ** 2)])
def test_rotation_measurement_statistics(Q, gate, p_one):
    """test-hadamard / quarter / three-quarter / eighth-rotation-measurements :134-182 via TEST-RANGE :112-130."""
    prog = f"DECLARE ro BIT\nI 1\n{gate}\nMEASURE 0 ro[0]\nRESET"
    reps = 3000

    def run():
        m = Q.make_qvm(2, seed=1234).load_program(prog)
        ones = 0
        for _ in range(reps):
            m.registers["ro"][:] = 0
            m.reset_quantum_state()
            m.run()
            ones += int(m.registers["ro"][0])
        got = ones / reps
        assert max(0.0, p_one - 0.05) <= got <= min(1.0, p_one + 0.05)
        assert max(0.0, 1 - p_one - 0.05) <= 1 - got <= min(1.0, 1 - p_one + 0.05)
    _modes(Q, run)


def test_measure_all_on_basis_states(Q):
    """test-measure-all :213-237 (and its basic-noise-qvm twin :239-264 without noise): X on the set bits of i, MEASURE-ALL returns
    those bits and leaves amplitude i = 1."""
    def run():
        for i in range(8):
            prog = "\n".join(("X" if i >> q & 1 else "I") + f" {q}" for q in range(3))
            m = Q.run_program(3, prog)
            bits = m.measure_all()
            assert list(bits) == [i >> q & 1 for q in range(3)]
            amps = m.amplitudes
            assert amps[i] == 1.0 and np.count_nonzero(amps) == 1
    _modes(Q, run)


def test_out_of_bounds_measurement(Q):
    """test-out-of-bounds-measurement :266-270."""
    for bad in (-1, 1):
        with pytest.raises(Exception):
            Q.make_qvm(1).measure(bad)
    assert Q.make_qvm(1).measure(0) == 0


def test_gate_tests(Q):
    """tests/gate-tests.lisp: test-hadamard :13-26, test-full-rotation :46-53, test-inversion :55-62, test-bell :75-89,
    test-swap :91-103, test-parametric-gate :105-113."""
    def run():
        for n in range(1, 11):
            amps = Q.run_program(n, "\n".join(f"H {q}" for q in range(n))).amplitudes
            np.testing.assert_allclose(np.abs(amps) ** 2, 2.0 ** -n, rtol=1e-12)
        for g in ("RX", "RY", "RZ"):
            amps = Q.run_program(1, "\n".join([f"{g}(pi/2) 0"] * 4)).amplitudes
            assert abs(abs(amps[0]) ** 2 - 1) < 1e-12
        assert abs(abs(Q.run_program(1, "X 0").amplitudes[1]) ** 2 - 1) < 1e-15
        for n in range(2, 11):          # GHZ: first and last amplitude have probability 1/2
            prog = "H 0\n" + "\n".join(f"CNOT {q} {q + 1}" for q in range(n - 1))
            p = np.abs(Q.run_program(n, prog).amplitudes) ** 2
            assert abs(p[0] - 0.5) < 1e-12 and abs(p[-1] - 0.5) < 1e-12
        amps = Q.run_program(2, "X 0\nSWAP 0 1").amplitudes
        assert amps[2] == 1.0 and np.count_nonzero(amps) == 1
        amps = Q.run_program(1, "DEFGATE G(%a):\n    cos(%a), sin(%a)\n    -sin(%a), cos(%a)\n\nG(0.0) 0").amplitudes
        assert amps[0] == 1.0 and amps[1] == 0.0
    _modes(Q, run)


# ---------------------------------------------------------------- tests/density-qvm-tests.lisp, tests/noisy-qvm-tests.lisp
def _pauli_noise_map(px, py, pz):
    """MAKE-PAULI-NOISE-MAP src/noisy-qvm.lisp:43-66 (the identity term comes first, as PUSH leaves it)."""
    from qvm_b200 import gates as G
    ops = [math.sqrt(p) * G.gate_matrix(s) for s, p in zip("XYZ", (px, py, pz))]
    psum = px + py + pz
    if psum < 1:
        ops.insert(0, math.sqrt(1.0 - psum) * np.eye(2, dtype=np.complex128))
    return ops


def _pauli_perturbed_1q_gate(name, px, py, pz):
    """MAKE-PAULI-PERTURBED-1Q-GATE src/noisy-qvm.lisp:69-78: the ideal gate followed by the noisy identity."""
    from qvm_b200 import gates as G
    u = G.gate_matrix(name)
    return [v @ u for v in _pauli_noise_map(px, py, pz)]


def _trace(m):
    return float(np.real(np.trace(m.state.matrix_view())))


def _purity(m):
    rho = m.state.matrix_view()
    return float(np.real(np.trace(rho @ rho)))


def test_density_qvm_parametric_gate_and_force_measurement(Q):
    """test-density-qvm-parametric-gate :25-34, -force-measurement-1q :36-44, -force-measurement-4q :46-60."""
    m = Q.make_density_qvm(1)
    m.load_program("DEFGATE G(%a):\n    cos(%a), sin(%a)\n    -sin(%a), cos(%a)\n\nG(0.0) 0").run()
    assert abs(np.real(m.state.matrix_view()[0, 0]) - 1) < 1e-4
    m = Q.make_density_qvm(1)
    m.load_program("H 0").run()
    m.state.vec.density_collapse(1, 0, 1, 1 / 0.5)                 # (force-measurement 1 0 state 0.5)
    assert abs(_trace(m) - 1) < 1e-12 and abs(np.real(m.state.matrix_view()[1, 1]) - 1) < 1e-12
    m = Q.make_density_qvm(4)
    m.load_program("H 0\nCNOT 0 1\nCNOT 1 3\nH 3").run()
    assert abs(_trace(m) - 1) < 1e-12
    p = m.state.vec.density_prob_excited(4, 3)
    m.state.vec.density_collapse(4, 3, 1, 1 / p)
    assert abs(_trace(m) - 1) < 1e-12


@pytest.mark.parametrize("p", [0.1, 0.4, 0.5, 0.6, 0.9])
def test_density_qvm_unitary_evolution_preserves_purity(Q, p):
    """test-density-qvm-1q-purity :90-98 and -2q-purity :100-108: mixtures loaded through (setf amplitudes)."""
    expected = (1 - p) ** 2 + p ** 2
    m = Q.make_density_qvm(1)
    m.amplitudes = np.diag([1 - p, p]).astype(np.complex128).ravel()
    assert abs(_purity(m) - expected) < 1e-12
    m.load_program("H 0").run()
    assert abs(_purity(m) - expected) < 1e-12
    m = Q.make_density_qvm(2)
    m.amplitudes = np.kron(np.diag([1 - p, p]), np.full((2, 2), 0.5)).astype(np.complex128).ravel()
    assert abs(_purity(m) - expected) < 1e-12
    m.load_program("CNOT 0 1").run()
    assert abs(_purity(m) - expected) < 1e-12


def test_density_qvm_noisy_readout_and_measure_all(Q):
    """test-noisy-readout-2q-qvm (tests/noisy-qvm-tests.lisp:114-139) on the density QVM (:110-112) and
    test-density-qvm-noisy-measure-all :116-140: POVM (0.8 0.1 / 0.2 0.9) on qubit 1."""
    m = Q.make_density_qvm(2, seed=7)
    m.load_program("DECLARE ro BIT[2]\nMEASURE 0 ro[0]\nMEASURE 1 ro[1]")
    wanted = {0, 1}
    tries = 500
    while wanted and tries > 0:
        tries -= 1
        m.reset_quantum_state()
        m.set_readout_povm(1, (0.8, 0.1, 0.2, 0.9))
        m.registers["ro"][:] = 0
        m.run()
        assert m.registers["ro"][0] == 0
        wanted.discard(int(m.registers["ro"][1]))
    assert tries > 0
    m = Q.make_density_qvm(2, seed=11)
    m.load_program("X 0\nX 1")
    wanted = {(1, 1), (1, 0)}
    tries = 500
    while wanted and tries > 0:
        tries -= 1
        m.reset_quantum_state()
        m.set_readout_povm(1, (0.8, 0.1, 0.2, 0.9))
        m.run()
        wanted.discard(tuple(m.measure_all()))
    assert tries > 0


def test_noisy_x_gate_with_certain_bit_flip(Q):
    """test-density-qvm-noisy-x-gate :158-169 and test-noisy-x-gate (tests/noisy-qvm-tests.lisp:98-112): X followed by a bit flip
    with probability 1 measures 0 -- on the density QVM (one superoperator) and on the pure-state QVM (stochastic Kraus)."""
    kraus = _pauli_perturbed_1q_gate("X", 1.0, 0.0, 0.0)
    m = Q.make_density_qvm(2)
    m.set_noisy_gate("X", (0,), kraus)
    m.load_program("DECLARE ro BIT\nX 0\nMEASURE 0 ro").run()
    assert m.registers["ro"][0] == 0
    for compiled in (False, True):
        Q.compile_before_running = compiled
        try:
            m = Q.make_qvm(2, seed=3)
            m.set_superoperator("X", (0,), kraus)
            m.load_program("DECLARE ro BIT\nX 0\nMEASURE 0 ro").run()
            assert m.registers["ro"][0] == 0
        finally:
            Q.compile_before_running = False
