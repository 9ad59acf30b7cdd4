"""Pins the oracle (oracle/qvm_oracle.c) against the reference's own known-answer tests, transcribed in
tests/golden/reference_kats.json (every entry cites the reference test file:line).  CPU only."""
import json
import math
import os

import numpy as np

from helpers import rand_state, rand_unitary
from oracle import oracle as O
from qvm_b200 import circuits as CC
from qvm_b200 import gates as G
from qvm_b200.quil import parse_quil

KATS = json.load(open(os.path.join(os.path.dirname(__file__), "golden", "reference_kats.json")))
SAMPLER_KATS = [(c["probs"], [tuple(d) for d in c["draws"]]) for c in KATS["sampler"]["cases"]]


def run_quil(n, lines, density=False):
    prog = parse_quil("\n".join(lines))
    st = O.zero_density(n) if density else O.zero_state(n)
    for ins in prog.instructions:
        m = prog.gate_matrix(ins)
        if density:
            O.density_apply_unitary(st, n, m, ins.qubits)
        else:
            O.apply_matrix(st, m, ins.qubits)
    return st


def test_bit_math():
    k = KATS["inject_bit"]
    for n, want in k["cases"]:
        assert O.inject_bit(k["x"], n) == want
    k = KATS["eject_bit"]
    for x, n in k["cases"]:
        assert O.eject_bit(x, n) == k["target"]
    k = KATS["index_to_address"]
    for q, addr in enumerate(k["addresses"]):
        assert O.index_to_address(k["index"], q, k["state"]) == addr
    k = KATS["nat_tuple"]
    assert list(O._nt(k["args"])) == k["stored"]


def test_matvec_kat():
    k = KATS["avx_matvec"]
    m = np.array([[complex(*e) for e in row] for row in k["matrix"]])
    v = np.array([complex(*e) for e in k["vector"]])
    O.apply_matrix(v, m, (0,))
    assert np.array_equal(v, np.array([complex(*e) for e in k["result"]]))


def test_qft2_truth_table():
    k = KATS["qft2"]
    for t, out in zip(k["inputs"], k["outputs"]):
        psi = O.zero_state(2)
        psi[:] = 0
        psi[t] = 1
        for m, q in CC.qft_circuit([0, 1]):
            O.apply_matrix(psi, m, q)
        np.testing.assert_allclose(psi, [complex(*e) for e in out], atol=1e-15)


def test_gate_programs():
    b = KATS["bell_pure"]
    psi = run_quil(2, b["program"])
    for i in b["indices"]:
        assert abs(psi[i] - b["amplitude"]) < 1e-15
    d = KATS["bell_density"]
    rho = run_quil(2, d["program"], density=True)
    for i in d["indices"]:
        assert abs(rho[i] - d["entry"]) < 1e-15
    assert abs(rho.sum() - 2.0) < 1e-14
    for n in range(1, KATS["bell_n"]["n_max"] + 1):
        psi = O.zero_state(n)
        for m, q in CC.bell_circuit(n):
            O.apply_matrix(psi, m, q)
        assert abs(abs(psi[0]) ** 2 - 0.5) < 1e-15 and abs(abs(psi[-1]) ** 2 - 0.5) < 1e-15
    for n in range(1, KATS["hadamard_n"]["n_max"] + 1):
        psi = O.zero_state(n)
        for m, q in CC.hadamard_circuit(n):
            O.apply_matrix(psi, m, q)
        np.testing.assert_allclose(np.abs(psi) ** 2, 2.0 ** -n, rtol=1e-13)
    assert abs(abs(run_quil(2, KATS["cnot_from_cz"]["program"])[-1]) ** 2 - 1) < 1e-14
    assert abs(abs(run_quil(2, KATS["swap"]["program"])[KATS["swap"]["index"]]) ** 2 - 1) < 1e-15
    assert abs(abs(run_quil(1, KATS["full_rotation"]["program"])[0]) ** 2 - 1) < 1e-14
    assert abs(abs(run_quil(3, KATS["inversion"]["program"])[-1]) ** 2 - 1) < 1e-15


def test_qubit_ordering_against_kron():
    """tests/gate-tests.lisp:28-44 pins compiled == interpreted on CNOT 2 0 / CSWAP 1 3 2; check the
    oracle against an independent dense construction (first Quil argument = most significant bit)."""
    n = 4
    rng = np.random.default_rng(0)
    for name, qs in [("CNOT", (2, 0)), ("CSWAP", (1, 3, 2)), ("CCNOT", (3, 0, 2)), ("SWAP", (1, 3))]:
        m = G.gate_matrix(name)
        psi = rand_state(n, 1)
        want = np.zeros_like(psi)
        k = len(qs)
        for i in range(1 << n):
            sub = sum(((i >> q) & 1) << (k - 1 - j) for j, q in enumerate(qs))
            for r in range(1 << k):
                if m[r, sub] != 0:
                    o = i
                    for j, q in enumerate(qs):
                        o = (o & ~(1 << q)) | (((r >> (k - 1 - j)) & 1) << q)
                    want[o] += m[r, sub] * psi[i]
        got = O.apply_matrix(psi.copy(), m, qs)
        np.testing.assert_allclose(got, want, atol=1e-15)
    # permutation gates: psi'[i] = psi[perm[i]] equals multiplying by the permutation matrix for involutions
    psi = rand_state(n, 2)
    a = O.apply_permutation(psi.copy(), [0, 1, 2, 3, 4, 6, 5, 7], (1, 3, 2))
    b = O.apply_matrix(psi.copy(), G.gate_matrix("CSWAP"), (1, 3, 2))
    np.testing.assert_allclose(a, b, atol=0)


def test_cdf_and_sampler_kats():
    np.testing.assert_allclose(O.cdf(np.array([1, 1, 1, 0], dtype=np.complex128))[:3], [1.0, 2.0, 3.0], atol=1e-15)
    v = np.full(4, math.sqrt(0.5), dtype=np.complex128)
    np.testing.assert_allclose(O.cdf(v)[:3], [0.5, 1.0, 1.5], atol=1e-15)
    v8 = np.sqrt(np.full(8, 0.1)).astype(np.complex128)
    for i in range(8):
        p = sum([0.1] * i) if i else 0.0          # the reference loops p from 0.0 by 0.1 (:192-194)
        assert O.sample_bisect(v8, p) == i
    for probs, cases in SAMPLER_KATS:
        v = np.sqrt(np.array(probs)).astype(np.complex128)
        for p, want in cases:
            assert O.sample_bisect(v, p) == want
            assert int(O.sample_tree(v, [p], True)[0]) == want


def test_samplers_agree_on_random_states():
    for n in (1, 4, 10, 12):
        a = rand_state(n, n)
        u = np.random.default_rng(2024).random(5000)
        seq = O.sample_multiple(a, u)
        tree = O.sample_tree(a, u, False)
        dist = O.sample_as_distribution(a, np.sort(u))
        cdf = O.cdf(a)
        for i in np.nonzero(seq != tree)[0]:
            lo, hi = sorted((int(seq[i]), int(tree[i])))
            assert np.abs(cdf[lo:hi + 1] - u[i]).min() < 1e-12
        assert (np.searchsorted(cdf, u, side="left").clip(0, a.size - 1) == seq).all()
        # strict rule == bisection sampler up to boundary draws
        bis = np.array([O.sample_bisect(a, p) for p in u[:200]])
        t2 = O.sample_tree(a, u[:200], True)
        assert (bis == t2).mean() > 0.99
        assert dist.size == u.size


def test_measurement_semantics():
    # measure-all on the 8 basis states (tests/measurement-tests.lisp:213-237)
    for i in range(8):
        psi = run_quil(3, [("X" if (i >> q) & 1 else "I") + f" {q}" for q in range(3)])
        b = O.measure_all(psi, 0.3)
        assert b == i and psi[i] == 1
    # interpreted vs compiled MEASURE rules (measurement.lisp:93-105, compile-gate.lisp:231-254)
    psi = run_quil(1, ["H 0"])
    assert O.measure(psi.copy(), 0, 0.3) == 1 and O.measure(psi.copy(), 0, 0.7) == 0
    assert O.measure_compiled(psi.copy(), 0, 0.3) == 0 and O.measure_compiled(psi.copy(), 0, 0.7) == 1
    c = psi.copy()
    O.measure(c, 0, 0.3)
    np.testing.assert_allclose(c, [0, 1], atol=1e-15)
    z = O.zero_state(2)
    assert O.measure(z, 1, 0.0) == 0          # p1 = 0 is deterministic even for r = 0


def test_density_kats():
    k = KATS["density_force_measurement_1q"]
    rho = run_quil(1, k["program"], density=True)
    O.density_force_measurement(rho, 1, k["force"][1], k["force"][0], k["force"][2])
    assert abs(rho[0] + rho[3] - 1) < 1e-15 and abs(rho[3] - 1) < 1e-15
    k = KATS["density_force_measurement_4q"]
    rho = run_quil(4, k["program"], density=True)
    assert abs(np.trace(rho.reshape(16, 16)) - 1) < 1e-14
    p = O.density_prob_excited(rho, 4, 3)
    O.density_force_measurement(rho, 4, 3, 1, p)
    assert abs(np.trace(rho.reshape(16, 16)) - 1) < 1e-14
    for p in KATS["density_purity"]["ps"]:
        rho = np.diag([1 - p, p]).astype(np.complex128).ravel().copy()
        O.density_apply_unitary(rho, 1, G.gate_matrix("H"), (0,))
        m = rho.reshape(2, 2)
        assert abs(np.trace(m @ m).real - ((1 - p) ** 2 + p ** 2)) < 1e-14
    rho = run_quil(1, ["H 0"], density=True)
    O.density_measure_discard(rho, 1, 0)
    np.testing.assert_allclose(rho.reshape(2, 2), np.diag([0.5, 0.5]), atol=1e-15)


def test_density_matches_dense_algebra():
    """Independent check of the vec(rho) conventions: rho' = sum_j K rho K^dagger on embedded operators."""
    n = 3
    rng = np.random.default_rng(3)
    psi = rand_state(n, 9)
    rho = np.outer(psi, psi.conj())
    vec = rho.ravel().copy()

    def embed(m, qs):
        k = len(qs)
        full = np.zeros((1 << n, 1 << n), dtype=np.complex128)
        for i in range(1 << n):
            sub = sum(((i >> q) & 1) << (k - 1 - j) for j, q in enumerate(qs))
            for r in range(1 << k):
                o = i
                for j, q in enumerate(qs):
                    o = (o & ~(1 << q)) | (((r >> (k - 1 - j)) & 1) << q)
                full[o, i] += m[r, sub]
        return full

    U = rand_unitary(2, rng)
    O.density_apply_unitary(vec, n, U, (2, 0))
    E = embed(U, (2, 0))
    rho = E @ rho @ E.conj().T
    np.testing.assert_allclose(vec.reshape(8, 8), rho, atol=1e-14)
    kr = G.depolarizing_kraus_map(0.3)
    O.density_apply_kraus(vec, n, kr, (1,))
    rho = sum(embed(k, (1,)) @ rho @ embed(k, (1,)).conj().T for k in kr)
    np.testing.assert_allclose(vec.reshape(8, 8), rho, atol=1e-14)
    assert abs(O.density_prob_excited(vec, n, 1) - sum(rho[i, i].real for i in range(8) if i & 2)) < 1e-14


def test_kraus_builders():
    # tests/basic-noise-qvm-tests.lisp:186-260 check completeness of the generated maps
    for p in (0.1, 0.5, 0.9):
        G.check_kraus_ops(G.depolarizing_kraus_map(p))
    G.check_kraus_ops(G.damping_kraus_map(5.0, 2.0))
    G.check_kraus_ops(G.dephasing_kraus_map(3.0, 1.0))
    kk = G.kraus_kron(G.depolarizing_kraus_map(0.2), G.damping_kraus_map(4.0, 1.0))
    assert len(kk) == 8 and kk[0].shape == (4, 4)
    G.check_kraus_ops(kk)
    assert np.allclose(G.kraus_kron([], [G.gate_matrix("X")])[0], np.kron(np.eye(2), G.gate_matrix("X")))


def test_multithreaded_baseline_equals_serial():
    n = 16
    rng = np.random.default_rng(5)
    a = rand_state(n, 1)
    b = a.copy()
    for _ in range(10):
        k = int(rng.integers(1, 4))
        q = tuple(int(x) for x in rng.choice(n, k, replace=False))
        m = rand_unitary(k, rng)
        O.apply_matrix(a, m, q)
        O.apply_matrix(b, m, q, threads=4)
    assert np.array_equal(a, b)


def test_inner_product_is_the_reference_loop():
    """app/src/api/expectation.lisp:79-84: sum of (conjugate a_i) * b_i; <psi|Z0|psi> of a Bell pair is 0 and
    <psi|psi> = 1 (the expectation API asserts a vanishing imaginary part, :73)."""
    from oracle import oracle as O
    rng = np.random.default_rng(5)
    a = rng.standard_normal(64) + 1j * rng.standard_normal(64)
    b = rng.standard_normal(64) + 1j * rng.standard_normal(64)
    assert abs(O.inner_product(a, b) - np.vdot(a, b)) < 1e-13
    bell = np.zeros(4, dtype=np.complex128)
    bell[0] = bell[3] = math.sqrt(0.5)
    z0 = bell.copy()
    z0[3] = -z0[3]
    assert abs(O.inner_product(bell, z0)) < 1e-15
    assert abs(O.inner_product(bell, bell) - 1.0) < 1e-15


def test_probabilities_and_wire_formats():
    """:wavefunction / :probabilities replies (app/src/handle-request.lisp:135-176) are streams of big-endian IEEE
    doubles (WRITE-64-BE, app/src/utilities.lisp:40-87); the Bell pair of tests/state-representation-tests.lisp:30-41."""
    import struct
    from oracle import oracle as O
    from qvm_b200 import qvm as Q
    bell = np.zeros(4, dtype=np.complex128)
    bell[0] = bell[3] = math.sqrt(0.5)
    p = O.probabilities(bell)
    assert np.allclose(p, [0.5, 0, 0, 0.5], atol=1e-15)
    wf = Q.wavefunction_octets(bell)
    assert len(wf) == 16 * 4
    assert wf[:8] == struct.pack(">d", math.sqrt(0.5)) and wf[8:16] == struct.pack(">d", 0.0)
    assert wf[48:56] == struct.pack(">d", math.sqrt(0.5))
    po = Q.probabilities_octets(p)
    assert po == b"".join(struct.pack(">d", x) for x in p)
    rng = np.random.default_rng(3)
    a = rng.standard_normal(32) + 1j * rng.standard_normal(32)
    assert np.array_equal(O.probabilities(a), a.real * a.real + a.imag * a.imag)
