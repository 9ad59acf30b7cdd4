"""GPU parity tests proper (-m gpu): the CUDA path, called through the C ABI, against the oracle on the
same seeded inputs.  Tolerance = north_star: 1e-12 relative / 1e-14 absolute on amplitudes and
probabilities; sampled indices bit-exact for identical uniforms."""
import math

import numpy as np
import pytest

from helpers import assert_close, rand_state, rand_unitary, random_circuit, run_oracle
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


def gpu_run(Q, psi, circ, fuse=True, absorb=False, each=False):
    n = int(math.log2(psi.size))
    vec = Q.DeviceVector(1 << n)
    vec.upload(psi)
    if each:
        for m, q in circ:
            vec.apply_matrix(m, q)
    else:
        vec.apply_gates(circ, fuse=fuse, absorb_swaps=absorb)
    out = vec.download()
    vec.close()
    return out


def test_library_loads_on_gpu(Q):
    from qvm_b200 import _lib
    assert _lib.device_count() >= 1


@pytest.mark.parametrize("n", [1, 2, 3, 5, 9, 13, 16])
def test_1q_gate_every_position(Q, O, n):
    rng = np.random.default_rng(n)
    for q in range(n):
        m = rand_unitary(1, rng)
        a = rand_state(n, q)
        ref = O.apply_matrix(a.copy(), m, (q,))
        assert_close(gpu_run(Q, a, [(m, (q,))], each=True), ref)


@pytest.mark.parametrize("n", [2, 4, 8, 13, 16])
def test_2q_gate_positions(Q, O, n):
    rng = np.random.default_rng(100 + n)
    pairs = [(a, b) for a in range(n) for b in range(n) if a != b]
    if len(pairs) > 40:
        pairs = [pairs[i] for i in rng.choice(len(pairs), 40, replace=False)]
    for a, b in pairs:
        m = rand_unitary(2, rng)
        s = rand_state(n, a * 31 + b)
        ref = O.apply_matrix(s.copy(), m, (a, b))
        assert_close(gpu_run(Q, s, [(m, (a, b))], each=True), ref)


def test_named_gates_and_permutations(Q, O):
    n = 12
    rng = np.random.default_rng(7)
    for name in ["CNOT", "CZ", "SWAP", "ISWAP", "CCNOT", "CSWAP"]:
        k = G.STANDARD_GATES[name][0]
        for _ in range(6):
            q = tuple(int(x) for x in rng.choice(n, k, replace=False))
            s = rand_state(n, 3)
            ref = O.apply_matrix(s.copy(), G.gate_matrix(name), q)
            assert_close(gpu_run(Q, s, [(G.gate_matrix(name), q)], each=True), ref)
    # permutation gates compiled as transpositions leave psi'[i] = psi[perm[i]] (compile-gate.lisp:259-309)
    s = rand_state(n, 4)
    ref = O.apply_permutation(s.copy(), [0, 1, 2, 3, 4, 5, 7, 6], (2, 9, 5))
    assert_close(gpu_run(Q, s, [(G.gate_matrix("CCNOT"), (2, 9, 5))], each=True), ref)


@pytest.mark.parametrize("k", [3, 4, 5, 8])
def test_dense_k_qubit_gates(Q, O, k):
    n = 11
    rng = np.random.default_rng(k)
    for _ in range(3):
        q = tuple(int(x) for x in rng.choice(n, k, replace=False))
        m = rand_unitary(k, rng)
        s = rand_state(n, k)
        ref = O.apply_matrix(s.copy(), m, q)
        assert_close(gpu_run(Q, s, [(m, q)], each=True), ref)


@pytest.mark.parametrize("n", [2, 10, 16, 20])
@pytest.mark.parametrize("fuse", [True, False])
def test_qft_parity(Q, O, n, fuse):
    circ = CC.qft_circuit(range(n))
    a = rand_state(n)
    ref = run_oracle(a.copy(), circ)
    assert_close(gpu_run(Q, a, circ, fuse=fuse), ref)


def test_qft_absorbed_swaps_download_canonical(Q, O):
    n = 15
    circ = CC.qft_circuit(range(n))
    a = rand_state(n)
    ref = run_oracle(a.copy(), circ)
    assert_close(gpu_run(Q, a, circ, fuse=True, absorb=True), ref)


@pytest.mark.parametrize("seed", range(10))
def test_random_circuits(Q, O, seed):
    rng = np.random.default_rng(500 + seed)
    n = int(rng.integers(3, 17))
    circ = random_circuit(n, 60, rng, max_dense=4)
    a = rand_state(n, seed)
    ref = run_oracle(a.copy(), circ)
    assert_close(gpu_run(Q, a, circ, fuse=True), ref)
    assert_close(gpu_run(Q, a, circ, fuse=False), ref)


def test_random_layer_circuit_20q(Q, O):
    n = 20
    circ = CC.random_layer_circuit(n, 3, seed=0)
    a = rand_state(n)
    ref = run_oracle(a.copy(), circ)
    assert_close(gpu_run(Q, a, circ, fuse=True), ref)


def test_compiled_tape_reuse(Q, O):
    n = 16
    circ = CC.qft_circuit(range(n))
    tape = Q.Tape(n, circ, fuse=True)
    info = tape.info()
    assert info["gates"] == len(circ) and info["passes"] <= 4
    for seed in (1, 2):
        a = rand_state(n, seed)
        vec = Q.DeviceVector(1 << n)
        vec.upload(a)
        vec.run_tape(tape)
        assert_close(vec.download(), run_oracle(a.copy(), circ))
        vec.close()


def test_reductions_and_collapse(Q, O):
    n = 15
    a = rand_state(n, 11)
    vec = Q.DeviceVector(1 << n)
    vec.upload(a)
    assert abs(vec.norm2() - O.norm2(a)) <= 1e-12
    for q in range(n):
        assert abs(vec.prob_excited(q) - O.prob_excited(a, q)) <= 1e-12
        assert abs(vec.prob_ground(q) - O.prob_ground(a, q)) <= 1e-12
    for q, keep in [(0, 1), (7, 0), (14, 1)]:
        p1 = O.prob_excited(a, q)
        ref = O.force_measurement(a.copy(), q, keep, p1)
        vec.upload(a)
        inv = 1 / math.sqrt(p1) if keep == 1 else 1 / math.sqrt(1 - p1)
        vec.collapse(q, keep, inv)
        assert_close(vec.download(), ref)
    vec.upload(3.0 * a)
    vec.normalize()
    assert_close(vec.download(), O.normalize(3.0 * a.copy()))
    vec.close()


@pytest.mark.parametrize("n", [1, 3, 10, 11, 17, 21])
def test_sampler_bit_exact(Q, O, n):
    a = rand_state(n, 5)
    u = np.random.default_rng(2024).random(100000 if n >= 10 else 2000)
    u[:3] = [0.0, 1.0, 0.5]
    vec = Q.DeviceVector(1 << n)
    vec.upload(a)
    for strict in (False, True):
        got = vec.sample(u, strict=strict)
        want = O.sample_tree(a, u, strict)
        assert (got == want).all(), f"{int((got != want).sum())} sampler mismatches"
    # against the reference's sequential-scan sampler: only draws within 1e-12 of a CDF step may differ
    got = vec.sample(u, strict=False)
    seq = O.sample_multiple(a, u)
    cdf = O.cdf(a)
    for i in np.nonzero(got != seq)[0]:
        lo, hi = sorted((int(got[i]), int(seq[i])))
        assert np.abs(cdf[lo:hi + 1] - u[i]).min() < 1e-12
    vec.close()


def test_sampler_known_answers(Q, O):
    # tests/measurement-tests.lisp:184-211 (C(b) > p rule)
    from test_oracle_golden import SAMPLER_KATS
    for probs, cases in SAMPLER_KATS:
        v = np.sqrt(np.array(probs, dtype=np.float64)).astype(np.complex128)
        vec = Q.DeviceVector(v.size)
        vec.upload(v)
        for p, want in cases:
            assert int(vec.sample([p], strict=True)[0]) == want
        vec.close()


def test_measure_all_basis_states(Q):
    # tests/measurement-tests.lisp:213-237
    for i in range(8):
        qvm = Q.make_qvm(3, seed=i)
        prog = "\n".join(("X" if (i >> q) & 1 else "I") + f" {q}" for q in range(3))
        qvm.load_program(prog).run()
        bits = qvm.measure_all()
        assert bits == [(i >> q) & 1 for q in range(3)]
        assert abs(qvm.amplitudes[i] - 1) < 1e-14


@pytest.mark.parametrize("n", [1, 2, 4, 6])
def test_density_unitary_and_kraus(Q, O, n):
    rng = np.random.default_rng(40 + n)
    rho = O.zero_density(n)
    vec = Q.DeviceVector(1 << (2 * n))
    vec.set_zero_state()
    for step in range(12):
        k = 1 if n == 1 or rng.integers(0, 2) else 2
        q = tuple(int(x) for x in rng.choice(n, k, replace=False))
        if step % 3 == 2:
            kr = G.depolarizing_kraus_map(0.1) if k == 1 else G.kraus_kron(G.depolarizing_kraus_map(0.1), G.damping_kraus_map(5.0, 1.0))
            O.density_apply_kraus(rho, n, kr, q)
            vec.density_apply_kraus(n, kr, q)
        else:
            U = rand_unitary(k, rng)
            O.density_apply_unitary(rho, n, U, q)
            vec.density_apply_kraus(n, [U], q)
    assert_close(vec.download(), rho)
    assert abs(vec.density_diag_probs(n).sum() - 1) < 1e-12
    np.testing.assert_allclose(vec.density_diag_probs(n), O.density_diag_probs(rho, n), rtol=1e-12, atol=1e-14)
    q = n - 1
    p1 = O.density_prob_excited(rho, n, q)
    assert abs(vec.density_prob_excited(n, q) - p1) < 1e-12
    O.density_force_measurement(rho, n, q, 1, p1)
    vec.density_collapse(n, q, 1, 1 / p1)
    assert_close(vec.download(), rho)
    O.density_measure_discard(rho, n, 0)
    vec.density_measure_discard(n, 0)
    assert_close(vec.download(), rho)
    vec.close()


def test_density_golden(Q):
    # tests/state-representation-tests.lisp:41-53: H 0; CNOT 0 1 -> entries 0,3,12,15 = 1/2
    st = Q.DensityMatrixState(2)
    Q.apply_gate_to_state(G.gate_matrix("H"), st, (0,))
    Q.apply_gate_to_state(G.gate_matrix("CNOT"), st, (0, 1))
    v = st.state_elements()
    for i in (0, 3, 12, 15):
        assert abs(v[i] - 0.5) < 1e-14
    # tests/density-qvm-tests.lisp:142-156: H 0; MEASURE 0 (discard) -> diag(1/2, 1/2)
    qvm = Q.make_density_qvm(1)
    qvm.load_program("H 0\nMEASURE 0").run()
    m = qvm.state.matrix_view()
    assert abs(m[0, 0] - 0.5) < 1e-14 and abs(m[1, 1] - 0.5) < 1e-14 and abs(m[0, 1]) < 1e-14
    # tests/density-qvm-tests.lisp:36-44: force-measurement 1 on H|0> keeps trace 1, rho[1,1] = 1
    qvm = Q.make_density_qvm(1)
    qvm.load_program("H 0").run()
    qvm.state.vec.density_collapse(1, 0, 1, 1 / 0.5)
    m = qvm.state.matrix_view()
    assert abs(np.trace(m) - 1) < 1e-14 and abs(m[1, 1] - 1) < 1e-14


def test_qvm_api_gate_tests(Q):
    # tests/gate-tests.lisp:162-183: 2-qubit QFT truth table through run-program
    expected = [[0.5, 0.5, 0.5, 0.5], [0.5, 0.5j, -0.5, -0.5j], [0.5, -0.5, 0.5, -0.5], [0.5, -0.5j, -0.5, 0.5j]]
    for compiled in (False, True):
        Q.compile_before_running = compiled
        try:
            for t in range(4):
                qvm = Q.make_qvm(2)
                qvm.load_program("\n".join((["X 0"] if t & 1 else []) + (["X 1"] if t & 2 else []) + ["I 0"]))
                qvm.run()
                qvm.state.vec.apply_gates(CC.qft_circuit([0, 1]))
                np.testing.assert_allclose(qvm.amplitudes, expected[t], atol=1e-14)
            # :64-73 CNOT from CZ, :46-53 four RX(pi/2)
            qvm = Q.run_program(2, "X 0\nH 1\nCZ 0 1\nH 1")
            assert abs(abs(qvm.amplitudes[3]) ** 2 - 1) < 1e-13
            qvm = Q.run_program(1, "\n".join(["RX(pi/2) 0"] * 4))
            assert abs(abs(qvm.amplitudes[0]) ** 2 - 1) < 1e-13
            # :28-44 qubit ordering: compiled == interpreted
            amps = Q.run_program(4, "H 0\nH 1\nH 2\nH 3\nCNOT 2 0\nCSWAP 1 3 2").amplitudes
            np.testing.assert_allclose(np.abs(amps) ** 2, 1 / 16, atol=1e-14)
        finally:
            Q.compile_before_running = False


def test_measure_statistics_and_rules(Q, O):
    # entangle-25's shape at 12 qubits: after MEASURE 0 the second MEASURE is deterministic
    prog = "H 0\n" + "\n".join(f"CNOT {i} {i + 1}" for i in range(11)) + "\nDECLARE ro BIT[2]\nMEASURE 0 ro[0]\nMEASURE 10 ro[1]"
    ones = 0
    for seed in range(40):
        for compiled in (False, True):
            Q.compile_before_running = compiled
            try:
                qvm = Q.make_qvm(12, seed=seed).load_program(prog).run()
            finally:
                Q.compile_before_running = False
            assert qvm.registers["ro"][0] == qvm.registers["ro"][1]
            ones += int(qvm.registers["ro"][0])
            assert abs(qvm.state.vec.norm2() - 1) < 1e-12
    assert 15 < ones < 65


def test_large_state_properties(Q):
    """Full-size checks through size-independent properties (no oracle at this size): QFT|0> is uniform,
    the norm is preserved, and the inverse circuit restores the input."""
    n = 28
    circ = CC.qft_circuit(range(n))
    vec = Q.DeviceVector(1 << n)
    vec.set_zero_state()
    vec.apply_gates(circ, fuse=True)
    assert abs(vec.norm2() - 1) < 1e-10
    head = vec.download(0, 1024)
    np.testing.assert_allclose(head, 2.0 ** (-n / 2), rtol=1e-12)
    tail = vec.download((1 << n) - 1024, 1024)
    np.testing.assert_allclose(tail, 2.0 ** (-n / 2), rtol=1e-12)
    for q in (0, 13, 27):
        assert abs(vec.prob_excited(q) - 0.5) < 1e-10
    inverse = [(G.dagger(m), q) for m, q in reversed(circ)]
    vec.apply_gates(inverse, fuse=True)
    z = vec.download(0, 4)
    assert abs(z[0] - 1) < 1e-10 and np.abs(z[1:]).max() < 1e-10
    assert abs(vec.norm2() - 1) < 1e-10
    vec.close()


def test_density_noisy_qaoa_fused_vs_oracle(Q, O):
    """Config 4 at small size: QAOA line graph + depolarizing after every gate, one fused library call."""
    n = 5
    circ = CC.qaoa_maxcut_circuit(n, CC.line_graph(n))
    dep = G.depolarizing_kraus_map(0.01)
    ops = []
    for m, q in circ:
        ops.append((m, q))
        for qq in q:
            ops.append((dep, (qq,)))
    rho = O.zero_density(n)
    for gate, q in ops:
        if isinstance(gate, list):
            O.density_apply_kraus(rho, n, gate, q)
        else:
            O.density_apply_unitary(rho, n, gate, q)
    st = Q.DensityMatrixState(n)
    st.apply_ops(ops, fuse=True)
    assert_close(st.state_elements(), rho)
    assert abs(st.measurement_probabilities().sum() - 1) < 1e-12
    # 2q channel given as kraus-kron product (16 operators): one 4-index-bit superoperator
    kk = G.kraus_kron(dep, dep)
    O.density_apply_kraus(rho, n, kk, (3, 1))
    st.apply_ops([(kk, (3, 1))])
    assert_close(st.state_elements(), rho)


def test_stochastic_kraus_on_pure_state(Q, O):
    """%evolve-pure-state-stochastically, same uniform draw on both sides."""
    n = 10
    kraus = G.depolarizing_kraus_map(0.5)
    kk = G.kraus_kron(G.damping_kraus_map(3.0, 1.0), G.dephasing_kraus_map(2.0, 1.0))
    for r, ks, qs in [(0.05, kraus, (3,)), (0.7, kraus, (9,)), (0.93, kraus, (0,)), (0.999, kraus, (5,)),
                      (0.4, kk, (7, 2)), (0.97, kk, (1, 8))]:
        psi = rand_state(n, int(r * 1000))
        st = Q.PureState(n)
        st.set_state_elements(psi)
        j_gpu = Q.evolve_pure_state_stochastically(ks, st, qs, r)
        ref = psi.copy()
        j_ref = O.evolve_stochastic(ref, ks, qs, r)
        assert j_gpu == j_ref
        assert_close(st.state_elements(), ref)
    # tests/state-representation-tests.lisp:55-66: depolarised "I 0" sometimes flips the measured bit
    ones = 0
    for seed in range(60):
        qvm = Q.make_qvm(2, seed=seed)
        qvm.set_superoperator("I", (0,), G.depolarizing_kraus_map(0.5))
        qvm.load_program("DECLARE R0 BIT\nI 0\nMEASURE 0 R0").run()
        ones += int(qvm.registers["R0"][0])
    assert 0 < ones < 60


def test_sixteen_amplitude_kernel_variant(tmp_path):
    """The 128-thread / 16-amplitudes-per-thread instantiation (QVMCUDA_REG_BITS=4, read once per process, hence
    the subprocess) must give the oracle's amplitudes too; the default runs use the 8-amplitude kernel."""
    import os
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    script = tmp_path / "m4.py"
    script.write_text(
        "import sys, numpy as np\n"
        f"sys.path.insert(0, {root!r}); sys.path.insert(0, {os.path.join(root, 'tests')!r})\n"
        "import helpers\n"
        "from qvm_b200 import qvm, circuits as CC\n"
        "n = 16\n"
        "rng = np.random.default_rng(4)\n"
        "circ = CC.qft_circuit(range(n)) + helpers.random_circuit(n, 80, rng, max_dense=3) + CC.random_layer_circuit(n, 3, 2)\n"
        "psi = helpers.rand_state(n, 9)\n"
        "ref = helpers.run_oracle(psi.copy(), circ)\n"
        "tape = qvm.Tape(n, circ, fuse=True)\n"
        "assert ' m=4 ' in tape.describe(), tape.describe()\n"
        "vec = qvm.DeviceVector(1 << n)\n"
        "vec.upload(psi)\n"
        "vec.run_tape(tape)\n"
        "helpers.assert_close(vec.download(), ref)\n"
        "print('m4 ok')\n")
    env = dict(os.environ, QVMCUDA_REG_BITS="4")
    out = subprocess.run([sys.executable, str(script)], env=env, capture_output=True, text=True, timeout=600)
    assert out.returncode == 0 and "m4 ok" in out.stdout, out.stdout + out.stderr


def test_permutation_tail_is_folded_and_exact(Q, O):
    """entangle-style circuits: the CNOT chain is folded into the write-back addressing; permutation-only circuits
    move integer labels exactly (dqvm's index-tracer idea, dqvm/tests/program-tests.lisp:14-19)."""
    n = 18
    circ = [(G.gate_matrix("H"), (0,))] + [(G.gate_matrix("CNOT"), (q, q + 1)) for q in range(n - 1)]
    tape = Q.Tape(n, circ, fuse=True)
    assert "store_perm" in tape.describe()
    psi = rand_state(n, 3)
    assert_close(gpu_run(Q, psi, circ), run_oracle(psi.copy(), circ))
    rng = np.random.default_rng(8)
    perm = []
    for _ in range(60):
        a, b, c = (int(x) for x in rng.choice(n, 3, replace=False))
        perm.append([(G.gate_matrix("SWAP"), (a, b)), (G.gate_matrix("CNOT"), (a, b)), (G.gate_matrix("X"), (a,)),
                     (G.gate_matrix("CCNOT"), (a, b, c))][int(rng.integers(0, 4))])
    labels = np.arange(1 << n).astype(np.complex128)
    assert np.array_equal(gpu_run(Q, labels, perm), run_oracle(labels.copy(), perm))


def test_inner_product_and_expectation(Q, O):
    """PURE-STATE-EXPECTATION (app/src/api/expectation.lisp:78-91) on the device: <prepared | OP prepared>."""
    n = 14
    a, b = rand_state(n, 1), rand_state(n, 2)
    va, vb = Q.DeviceVector(1 << n), Q.DeviceVector(1 << n)
    va.upload(a)
    vb.upload(b)
    got = va.inner_product(vb)
    ref = O.inner_product(a, b)
    assert abs(got - ref) < 1e-13
    assert abs(va.inner_product(va) - 1.0) < 1e-13
    va.close()
    vb.close()
    # <Z0>, <X1>, <Z0 Z3> on a prepared state, against the oracle run of the same programs
    prep = "H 0\nCNOT 0 1\nRX(0.3) 2\nRY(1.1) 3\nCZ 2 3\nH 4"
    ops = ["Z 0", "X 1", "Z 0\nZ 3", "I 0"]
    got = Q.perform_expectation(prep, ops, 5)
    from qvm_b200.quil import parse_quil
    psi = O.zero_state(5)
    for ins in parse_quil(prep).instructions:
        O.apply_matrix(psi, G.gate_matrix(ins.name, ins.params), tuple(ins.qubits))
    for op, g in zip(ops, got):
        phi = psi.copy()
        for ins in parse_quil(op).instructions:
            O.apply_matrix(phi, G.gate_matrix(ins.name, ins.params), tuple(ins.qubits))
        assert abs(g - O.inner_product(psi, phi).real) < 1e-13
    assert abs(got[3] - 1.0) < 1e-13


def test_probabilities_export(Q, O):
    """PERFORM-PROBABILITIES on the device: |psi_i|^2 bit for bit as the oracle's PROBABILITY (explicitly rounded)."""
    n = 15
    psi = rand_state(n, 6)
    vec = Q.DeviceVector(1 << n)
    vec.upload(psi)
    ref = O.probabilities(psi)
    assert np.array_equal(vec.probabilities(), ref)
    assert np.array_equal(vec.probabilities(1000, 777), ref[1000:1777])
    assert Q.probabilities_octets(vec.probabilities(0, 4)) == Q.probabilities_octets(ref[:4])
    vec.close()


# ---------------------------------------------------------------- lazy reset (SET-TO-ZERO-STATE as a flag)
@pytest.mark.parametrize("n,basis", [(12, 0), (14, 0), (14, 5), (15, (1 << 15) - 1), (16, 0x8421), (18, 1 << 17)])
def test_lazy_reset_first_pass_synthesises_its_tiles(Q, O, n, basis):
    """set_basis_state only records the basis index; the first compiled pass that follows builds its tiles instead of loading
    them (zero tiles are written back as zeros).  Same amplitudes as the oracle run from the explicit basis vector."""
    from qvm_b200 import _lib
    rng = np.random.default_rng(n * 977 + basis)
    circ = CC.qft_circuit(range(n)) + random_circuit(n, 40, rng, max_dense=3)
    ref = np.zeros(1 << n, dtype=np.complex128)
    ref[basis] = 1.0
    run_oracle(ref, circ)
    for mode in ("apply_gates", "tape", "unfused"):
        vec = Q.DeviceVector(1 << n)
        junk = rand_state(n, 3)
        vec.upload(junk)                      # the buffer holds something else: a skipped reset would show
        vec.set_basis_state(basis)
        l0 = _lib.launch_count()
        if mode == "apply_gates":
            vec.apply_gates(circ, fuse=True)
        elif mode == "tape":
            tape = Q.Tape(n, circ, fuse=True)
            vec.run_tape(tape)
            tape.close()
        else:
            vec.apply_gates(circ, fuse=False)  # first pass runs on the interpreter: the vector is written first
        assert _lib.launch_count() > l0
        assert_close(vec.download(), ref)
        vec.close()


def test_lazy_reset_is_invisible_to_every_reader(Q, O):
    n = 14
    for basis in (0, 9, (1 << n) - 2):
        want = np.zeros(1 << n, dtype=np.complex128)
        want[basis] = 1.0
        u = np.random.default_rng(1).random(64)

        def fresh():
            v = Q.DeviceVector(1 << n)
            v.upload(rand_state(n, 8))
            v.set_basis_state(basis)
            return v

        v = fresh(); assert np.array_equal(v.download(), want); v.close()
        v = fresh(); assert v.norm2() == 1.0; v.close()
        for q in (0, 3, n - 1):
            v = fresh(); assert v.prob_excited(q) == float((basis >> q) & 1); v.close()
        v = fresh(); assert (v.sample(u, strict=False) == basis).all(); v.close()
        v = fresh(); p = v.probabilities(); assert np.array_equal(p, np.abs(want) ** 2); v.close()
        v = fresh(); w = Q.DeviceVector(1 << n); w.copy_from(v); assert np.array_equal(w.download(), want); v.close(); w.close()
        v = fresh(); v.collapse(0, basis & 1, 1.0); assert np.array_equal(v.download(), want); v.close()
        v = fresh(); v.upload(want[:8] * 0 + 2.0, offset=0); got = v.download()      # partial upload lands on the written vector
        chk = want.copy(); chk[:8] = 2.0
        assert np.array_equal(got, chk); v.close()
        # two resets in a row, then gates
        v = fresh(); v.set_zero_state(); v.apply_gates([(G.gate_matrix("H"), (q,)) for q in range(n)], fuse=True)
        assert_close(v.download(), np.full(1 << n, 2.0 ** (-n / 2), dtype=np.complex128)); v.close()


def test_schedule_cache_hits_only_on_identical_programs(Q, O):
    """qvmcuda_apply_gates keeps its last few schedules per state (the reference compiles a loaded program once and runs it
    many times): a repeated gate list is served from the cache, a list that differs in one angle, one qubit or the starting
    layout is not."""
    n = 13
    rng = np.random.default_rng(77)
    base = CC.qft_circuit(range(n)) + random_circuit(n, 30, rng, max_dense=3)
    variants = [base]
    v = list(base)
    v[5] = (G.gate_matrix("RZ", [0.123]), (4,))
    variants.append(v)
    v = list(base)
    v[7] = (v[7][0], tuple((q + 1) % n for q in v[7][1]))
    variants.append(v)
    variants += [random_circuit(n, 25, rng, max_dense=3) for _ in range(4)]     # more programs than cache slots
    psi = rand_state(n, 5)
    vec = Q.DeviceVector(1 << n)
    for rep in range(3):
        for circ in variants:
            vec.upload(psi)
            vec.apply_gates(circ, fuse=True)
            ref = run_oracle(psi.copy(), circ)
            assert_close(vec.download(), ref)
    # same program twice in a row without re-upload (state differs, schedule identical), then from an absorbed-swap layout
    vec.upload(psi)
    vec.apply_gates(base, fuse=True)
    vec.apply_gates(base, fuse=True)
    ref = run_oracle(run_oracle(psi.copy(), base), base)
    assert_close(vec.download(), ref)
    vec.upload(psi)
    vec.apply_gates(base, fuse=True, absorb_swaps=True)     # leaves a permuted layout behind
    vec.apply_gates(base, fuse=True)                        # same list, different starting layout: must not reuse the tape above
    assert_close(vec.download(), ref)
    vec.close()
