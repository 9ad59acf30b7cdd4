"""GPU tests of the boundary in the order the reference's call stacks use it:
compiled-mode transitions (APPLY-GATE-TO-STATE with QUBITS = NIL, compiled MEASURE), the density machine's gate tape,
NAIVE-MEASURE-ALL on rho, and the remaining BASELINE.json configurations at sizes the oracle can follow."""
import numpy as np
import pytest

import helpers as H
from qvm_b200 import circuits as CC
from qvm_b200 import gates as G

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def Q():
    from qvm_b200 import qvm
    return qvm


@pytest.fixture(scope="module")
def O():
    from oracle import oracle
    return oracle


def test_compiled_mode_call_stack(Q, O):
    """src/transition.lisp:182-194 on a device state: TRANSITION hands a compiled gate application to
    APPLY-GATE-TO-STATE with QUBITS = NIL -- the qubits come from the instruction (lisp/operators.lisp) -- and the
    compiled MEASURE decides with p0 and `r < p0` (src/compile-gate.lisp:231-254).  Nothing runs on a host copy."""
    n = 10
    rng = np.random.default_rng(4)
    circ = CC.qft_circuit(range(n)) + H.random_circuit(n, 30, rng)
    psi = H.rand_state(n, 9)
    state = Q.PureState(n)
    state.vec.upload(psi)
    program = [Q.CompiledGateApplication(m, q) for m, q in circ]      # COMPILE-LOADED-PROGRAM's output, stand-in
    for instr in program:
        Q.apply_gate_to_state(instr, state, None)                      # (apply-gate-to-state instr (state qvm) nil)
    ref = H.run_oracle(psi.copy(), circ)
    H.assert_close(state.vec.download(), ref)
    with pytest.raises(ValueError):
        Q.apply_gate_to_state(circ[0][0], state, None)                 # a bare matrix has no qubits of its own
    # compiled MEASURE: same rule, same uniform as the oracle
    q = 3
    p0 = O.prob_ground(ref, q) if hasattr(O, "prob_ground") else 1.0 - O.prob_excited(ref, q)
    assert abs(state.vec.prob_ground(q) - p0) < 1e-13
    r = 0.41
    bit = 0 if r < p0 else 1
    inv = 1.0 / np.sqrt(p0) if bit == 0 else 1.0 / np.sqrt(1.0 - p0)
    state.vec.collapse(q, bit, inv)
    O.force_measurement(ref, q, bit, O.prob_excited(ref, q))
    H.assert_close(state.vec.download(), ref)
    state.vec.close()


def test_density_measure_all_matches_oracle(Q, O):
    """NAIVE-MEASURE-ALL on rho (src/measurement.lisp:153-162) with identical uniforms."""
    n = 4
    prog = "H 0\nCNOT 0 1\nRX(0.7) 2\nCNOT 2 3\nRY(1.1) 1"
    for seed in range(6):
        qvm = Q.DensityQVM(n, seed=seed)
        qvm.set_noisy_gate("RX", (2,), G.depolarizing_kraus_map(0.2))
        qvm.load_program(prog).run()
        rho = O.zero_density(n)
        for name, params, qubits in (("H", [], (0,)), ("CNOT", [], (0, 1))):
            O.density_apply_unitary(rho, n, G.gate_matrix(name, params), qubits)
        O.density_apply_kraus(rho, n, G.depolarizing_kraus_map(0.2), (2,))
        O.density_apply_unitary(rho, n, G.gate_matrix("CNOT"), (2, 3))
        O.density_apply_unitary(rho, n, G.gate_matrix("RY", [1.1]), (1,))
        H.assert_close(qvm.amplitudes, rho)
        uniforms = np.random.default_rng(1000 + seed).random(n)
        # (a probability that is exactly 0 in the oracle must be exactly 0 on the device too, or the two would consume a
        # different number of draws: the collapse writes exact zeros, so it is)
        it = iter(uniforms)
        qvm.random = lambda: float(next(it))
        bits = qvm.measure_all()
        want = O.density_measure_all(rho, n, uniforms)
        assert bits == want
        H.assert_close(qvm.amplitudes, rho)
        assert abs(np.trace(qvm.state.matrix_view()).real - 1.0) < 1e-12


@pytest.mark.parametrize("n", [8, 9])
def test_density_qaoa_with_two_qubit_channels(Q, O, n):
    """BASELINE configs[3] at a size the oracle follows: noisy QAOA on vec(rho) with the depolarizing channel after every
    gate -- 1q channels from DEPOLARIZING-KRAUS-MAP, 2q channels as their KRAUS-KRON tensor (16 Kraus operators on four
    index bits, src/basic-noise-qvm.lisp:251-269) -- through the density machine's gate tape in ONE library call."""
    circ = CC.qaoa_maxcut_circuit(n, CC.line_graph(n))
    dep = G.depolarizing_kraus_map(0.01)
    dep2 = G.kraus_kron(dep, dep)
    ops = []
    for m, q in circ:
        ops.append(([m], q))
        ops.append((dep if len(q) == 1 else dep2, q))
    st = Q.DensityMatrixState(n)
    st.vec.density_apply_ops(n, ops)
    rho = O.zero_density(n)
    for kraus, q in ops:
        O.density_apply_kraus(rho, n, kraus, q)
    H.assert_close(st.state_elements(), rho)
    probs = st.measurement_probabilities()
    assert abs(probs.sum() - 1.0) < 1e-12
    np.testing.assert_allclose(probs, O.density_diag_probs(rho, n), rtol=1e-12, atol=1e-14)
    st.vec.close()


def test_density_machine_batches_its_transitions(Q):
    """The 14-qubit configuration's circuit shape at 7 qubits: DensityQVM.run must reach the fused schedule through the
    protocol (one library call per stretch of gate transitions), not one pass per operator."""
    from qvm_b200 import _lib
    n = 7
    lines = [f"H {q}" for q in range(n)]
    for a, b in CC.line_graph(n):
        lines += [f"CNOT {a} {b}", f"RZ(1.4) {b}", f"CNOT {a} {b}"]
    lines += [f"RX(0.6) {q}" for q in range(n)]
    qvm = Q.DensityQVM(n, seed=1)
    for q in range(n):
        for name in ("H", "RX", "RZ"):
            pass
    qvm.load_program("\n".join(lines))
    before = _lib.launch_count()
    qvm.run()
    launches = _lib.launch_count() - before
    n_ops = len(lines)
    assert launches <= 8 < n_ops, f"{launches} kernel launches for {n_ops} operators"
    assert abs(qvm.state.measurement_probabilities().sum() - 1.0) < 1e-12


def test_bench_20H_on_gpu(Q):
    """BASELINE configs[0]: bench/20H.quil verbatim on PURE-STATE-QVM -- every amplitude 2^-10."""
    circ, _, n = H.load_bench_circuit("20H")
    vec = Q.DeviceVector(1 << n)
    vec.set_zero_state()
    vec.apply_gates(circ, fuse=True)
    amps = vec.download()
    np.testing.assert_allclose(amps, 2.0 ** -10, rtol=0, atol=1e-14)
    vec.set_zero_state()
    vec.apply_gates(circ, fuse=False)
    np.testing.assert_allclose(vec.download(), 2.0 ** -10, rtol=0, atol=1e-14)
    vec.close()


def test_qft_30_properties(Q):
    """BASELINE configs[1] at full size through size-independent properties: QFT|x> has uniform magnitudes and the phase
    ramp exp(2 pi i x k / 2^n) on the low indices; the norm survives; fused and unfused prefixes agree."""
    n = 30
    circ = CC.qft_circuit(range(n))
    vec = Q.DeviceVector(1 << n)
    x = 5
    vec.set_basis_state(x)
    vec.apply_gates(circ, fuse=True)
    assert abs(vec.norm2() - 1.0) < 1e-10
    head = vec.download(0, 4096)
    k = np.arange(4096)
    want = np.exp(2j * np.pi * x * k / float(1 << n)) * 2.0 ** (-n / 2)
    np.testing.assert_allclose(head, want, rtol=0, atol=1e-12 * 2.0 ** (-n / 2) * 100)
    for q in (0, 15, 29):
        assert abs(vec.prob_excited(q) - 0.5) < 1e-10
    vec.close()


def _dense_gate_cases(n, rng):
    cases = []
    for k in (3, 4, 5, 6, 8):
        qs = tuple(int(x) for x in rng.choice(n, k, replace=False))
        cases.append((H.rand_unitary(k, rng), qs))
    # a controlled 3-qubit block: the control is a non-mixing qubit of a 4-qubit gate
    u3 = H.rand_unitary(3, rng)
    cu = np.eye(16, dtype=np.complex128)
    cu[8:, 8:] = u3
    cases.append((cu, tuple(int(x) for x in rng.choice(n, 4, replace=False))))
    return cases


def test_dense_k_qubit_gates_tensor_path(Q, O):
    """APPLY-OPERATOR for k >= 3 (src/wavefunction.lisp:234-306) through the DMMA kernel (the default for 3 <= k <= 8)."""
    n = 13
    rng = np.random.default_rng(77)
    psi = H.rand_state(n, 2)
    vec = Q.DeviceVector(1 << n)
    for m, qs in _dense_gate_cases(n, rng):
        vec.upload(psi)
        vec.apply_matrix(m, qs)
        ref = O.apply_matrix(psi.copy(), m, qs)
        H.assert_close(vec.download(), ref)
    vec.close()


def test_dense_k_qubit_gates_scalar_path(tmp_path):
    """The same cases through the scalar kernel (QVMCUDA_BIG=scalar, read once per process): the A/B partner of the
    tensor-path kernel and the kernel that serves k > 8."""
    import os
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    script = tmp_path / "scalar_big.py"
    script.write_text(
        "import sys, numpy as np\n"
        f"sys.path.insert(0, {root!r}); sys.path.insert(0, {os.path.join(root, 'tests')!r})\n"
        "import helpers as H\n"
        "from oracle import oracle as O\n"
        "from qvm_b200 import qvm\n"
        "from test_gpu_boundary import _dense_gate_cases\n"
        "n = 13\n"
        "rng = np.random.default_rng(77)\n"
        "psi = H.rand_state(n, 2)\n"
        "vec = qvm.DeviceVector(1 << n)\n"
        "for m, qs in _dense_gate_cases(n, rng):\n"
        "    vec.upload(psi); vec.apply_matrix(m, qs)\n"
        "    H.assert_close(vec.download(), O.apply_matrix(psi.copy(), m, qs))\n"
        "print('scalar ok')\n")
    env = dict(os.environ, QVMCUDA_BIG="scalar")
    out = subprocess.run([sys.executable, str(script)], env=env, capture_output=True, text=True, timeout=600)
    assert out.returncode == 0 and "scalar ok" in out.stdout, out.stdout + out.stderr


def test_two_qubit_channels_on_rho_use_dense_path(Q, O):
    """KRAUS-KRON 2q depolarizing channels are 16x16 superoperators on four index bits of vec(rho): the k = 4 dense path."""
    n = 5
    dep2 = G.kraus_kron(G.depolarizing_kraus_map(0.05), G.depolarizing_kraus_map(0.1))
    st = Q.DensityMatrixState(n)
    ops = [([G.gate_matrix("H")], (0,)), ([G.gate_matrix("CNOT")], (0, 3)), (dep2, (0, 3)), (dep2, (4, 1))]
    st.vec.density_apply_ops(n, ops)
    rho = O.zero_density(n)
    for kraus, q in ops:
        O.density_apply_kraus(rho, n, kraus, q)
    H.assert_close(st.state_elements(), rho)
    st.vec.close()


def test_mixed_state_expectation_on_device(Q, O):
    """MIXED-STATE-EXPECTATION (app/src/api/expectation.lisp:91-107): tr(Q rho) reduced on the device."""
    n = 5
    rng = np.random.default_rng(8)
    qvm = Q.DensityQVM(n, seed=3)
    qvm.set_noisy_gate("X", (1,), G.depolarizing_kraus_map(0.3))
    qvm.load_program("H 0\nCNOT 0 1\nX 1\nRY(0.4) 2\nCNOT 2 4\nRX(1.3) 3").run()
    rho = qvm.state.matrix_view()
    a = rng.standard_normal((1 << n, 1 << n)) + 1j * rng.standard_normal((1 << n, 1 << n))
    herm = a + a.conj().T
    got = Q.mixed_state_expectation(qvm, herm)
    want = np.trace(herm @ rho)
    assert abs(got - want) <= 1e-12 * abs(want) + 1e-13
    assert abs(got.imag) < 1e-12
    z0 = np.kron(np.eye(1 << (n - 1)), np.diag([1.0, -1.0]))        # Z on qubit 0
    assert abs(Q.mixed_state_expectation(qvm, z0) - (1 - 2 * O.density_prob_excited(np.ascontiguousarray(rho.ravel()), n, 0))) < 1e-12


def test_unitary_qvm_matches_gate_products(Q, O):
    """UNITARY-QVM (src/unitary-qvm.lisp:30-137): the pure-state path on 2n bits computes the matrix of a program."""
    n = 4
    prog = "H 0\nCNOT 0 2\nRZ(0.3) 1\nCPHASE(0.9) 1 3\nISWAP 2 3\nRX(1.1) 0"
    u = Q.parsed_program_unitary_matrix(prog, n)
    # column c of the matrix = the program applied to basis state c (oracle, one column at a time)
    from qvm_b200.quil import parse_quil
    program = parse_quil(prog)
    circ = [(program.gate_matrix(x), x.qubits) for x in program.instructions]
    for c in range(1 << n):
        e = np.zeros(1 << n, dtype=np.complex128)
        e[c] = 1.0
        H.assert_close(u[:, c], H.run_oracle(e, circ))
    np.testing.assert_allclose(u.conj().T @ u, np.eye(1 << n), atol=1e-13)


def test_shared_memory_persistent_wavefunction(Q, O, tmp_path):
    """The app's --shared mode on a device state: the POSIX shared-memory object holds the host-visible copy; refresh() downloads
    the device state straight into the mapping, a client attaches through the info socket, push() uploads what it wrote."""
    import uuid
    from qvm_b200 import shm
    n = 12
    m = Q.make_qvm(n)
    name = f"QVMGPU{uuid.uuid4().hex[:10]}"
    sw = shm.share_wavefunction(m, name, socket_dir=str(tmp_path))
    try:
        view = shm.attach(name, str(tmp_path))
        assert view[0] == 1.0 and not view[1:].any()
        m.load_program("\n".join(f"H {q}" for q in range(n)) + "\nCNOT 0 1\n").run()
        assert view[0] == 1.0                      # not refreshed yet: the amplitudes live in HBM
        sw.refresh()
        H.assert_close(view, np.full(1 << n, 2.0 ** (-n / 2), dtype=np.complex128))
        psi = H.rand_state(n, 4)
        view[:] = psi
        sw.push()
        H.assert_close(m.amplitudes, psi)
        del view
    finally:
        sw.close()


def test_copy_wavefunction_every_length_combination(Q):
    """tests/wavefunction-tests.lisp:35-75 (test-copy-wavefunction): COPY-WAVEFUNCTION copies min(|src|, |dst|) amplitudes for every
    combination of vector lengths and leaves the rest of the destination alone (a device state has at least one qubit, so the
    reference's 1-element case starts at 2 here)."""
    srcs = {k: np.arange(1, k + 1).astype(np.complex128) for k in (2, 4, 8, 16)}
    for ks, a in srcs.items():
        for kd in (2, 4, 8, 16):
            s, d = Q.DeviceVector(ks), Q.DeviceVector(kd)
            s.upload(a)
            d.upload(np.full(kd, -7.0, dtype=np.complex128))
            d.copy_from(s)
            got = d.download()
            m = min(ks, kd)
            assert np.array_equal(got[:m], a[:m]) and (got[m:] == -7.0).all()
            s.close()
            d.close()


def test_unitary_qvm_reference_known_answers(Q):
    """tests/unitary-tests.lisp:3-24: CNOT 0 1; CNOT 1 0; CNOT 0 1 is the SWAP matrix; a unitary QVM refuses MEASURE."""
    u = Q.parsed_program_unitary_matrix("CNOT 0 1\nCNOT 1 0\nCNOT 0 1", 2)
    swap = np.array([[1, 0, 0, 0], [0, 0, 1, 0], [0, 1, 0, 0], [0, 0, 0, 1]], dtype=np.complex128)
    assert np.array_equal(u, swap)
    m = Q.UnitaryQVM(2)
    m.load_program("CNOT 0 1\nMEASURE 0")
    with pytest.raises(Exception):
        m.run()
