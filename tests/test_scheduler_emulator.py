"""Host logic: the gate scheduler's tile programs, interpreted on the CPU through the same op code
the CUDA kernel compiles (tests/support/qv_emulator.cpp), must reproduce the oracle."""
import numpy as np
import pytest

from helpers import assert_close, rand_state, random_circuit, run_emulator, run_oracle
from qvm_b200 import circuits as CC


@pytest.mark.parametrize("n,tile_bits", [(1, 12), (2, 12), (3, 12), (5, 12), (8, 4), (9, 5), (10, 6), (13, 12), (15, 12)])
@pytest.mark.parametrize("fuse", [True, False])
@pytest.mark.parametrize("reg_bits", [3, 4])
def test_qft_matches_oracle(n, tile_bits, fuse, reg_bits):
    circ = CC.qft_circuit(range(n))
    a = rand_state(n)
    b = a.copy()
    steps, desc, _ = run_emulator(a, n, circ, fuse=fuse, tile_bits=tile_bits, reg_bits=reg_bits)
    run_oracle(b, circ)
    assert_close(a, b)
    if not fuse:
        assert steps >= len(circ)


@pytest.mark.parametrize("seed", range(24))
def test_random_circuits_match_oracle(seed):
    reg_bits = (3, 4, 0)[seed % 3]
    rng = np.random.default_rng(1000 + seed)
    n = int(rng.integers(1, 14))
    tile_bits = int(rng.integers(3, 13))
    circ = random_circuit(n, int(rng.integers(5, 60)), rng, max_dense=4)
    a = rand_state(n, seed)
    b = a.copy()
    run_emulator(a, n, circ, fuse=True, tile_bits=tile_bits, reg_bits=reg_bits)
    run_oracle(b, circ)
    assert_close(a, b)
    c = rand_state(n, seed)
    run_emulator(c, n, circ, fuse=False, tile_bits=tile_bits, reg_bits=reg_bits)
    assert_close(c, b)


def test_fusion_packs_qft_into_few_passes():
    n = 16
    circ = CC.qft_circuit(range(n))
    a = rand_state(n)
    steps, desc, _ = run_emulator(a, n, circ, fuse=True, tile_bits=12)
    assert steps <= 4, desc


def test_absorbed_swaps_relabel_qubits():
    n = 10
    circ = CC.qft_circuit(range(n))
    a = rand_state(n)
    b = a.copy()
    _, _, l2p = run_emulator(a, n, circ, fuse=True, tile_bits=8, absorb_swaps=True)
    run_oracle(b, circ)
    # physical index bit l2p[q] holds logical qubit q: un-permute and compare
    idx = np.arange(1 << n)
    phys = np.zeros_like(idx)
    for q in range(n):
        phys |= ((idx >> q) & 1) << int(l2p[q])
    assert_close(a[phys], b)


def test_hadamard_and_bell_known_answers():
    # tests/gate-tests.lisp:13-26 (H^n uniform) and :75-89 (Bell first/last probability 1/2)
    from oracle import oracle as O
    for n in range(1, 11):
        a = O.zero_state(n)
        run_emulator(a, n, CC.hadamard_circuit(n))
        assert np.allclose(np.abs(a) ** 2, 2.0 ** -n, atol=1e-14)
        b = O.zero_state(n)
        run_emulator(b, n, CC.bell_circuit(n))
        assert abs(abs(b[0]) ** 2 - 0.5) < 1e-14 and abs(abs(b[-1]) ** 2 - 0.5) < 1e-14


@pytest.mark.parametrize("reg_bits", [3, 4])
def test_trailing_permutation_gates_fold_into_the_write_back(reg_bits):
    """X / CNOT / SWAP at the end of a pass are affine maps of the tile-local index: the scheduler folds them into
    the store addressing (no arithmetic).  Mixed with gates they do and do not commute with."""
    from qvm_b200 import gates as G
    n = 11
    rng = np.random.default_rng(77)
    for trial in range(6):
        circ = [(G.gate_matrix("H"), (int(q),)) for q in rng.choice(n, 4, replace=False)]
        circ += [(G.gate_matrix("CPHASE", [0.4]), (0, 5)), (G.gate_matrix("RX", [0.3]), (2,))]
        for _ in range(12):
            a, b = (int(x) for x in rng.choice(n, 2, replace=False))
            kind = int(rng.integers(0, 4))
            if kind == 0:
                circ.append((G.gate_matrix("CNOT"), (a, b)))
            elif kind == 1:
                circ.append((G.gate_matrix("SWAP"), (a, b)))
            elif kind == 2:
                circ.append((G.gate_matrix("X"), (a,)))
            else:
                circ.append((G.gate_matrix("CZ"), (a, b)))       # diagonal: commutes with controls only
        a0 = rand_state(n, trial)
        b0 = a0.copy()
        steps, desc, _ = run_emulator(a0, n, circ, fuse=True, tile_bits=12, reg_bits=reg_bits)
        run_oracle(b0, circ)
        assert_close(a0, b0)
        assert "store_perm" in desc, desc
    # a pure SWAP network needs no rounds at all
    circ = [(G.gate_matrix("SWAP"), (q, n - 1 - q)) for q in range(n // 2)]
    a0 = rand_state(n, 5)
    b0 = a0.copy()
    steps, desc, _ = run_emulator(a0, n, circ, fuse=True, tile_bits=12, reg_bits=reg_bits)
    run_oracle(b0, circ)
    assert np.array_equal(a0, b0)
    assert "rounds=0" in desc, desc


def test_wide_diagonal_gate():
    """A diagonal on more than 8 qubits runs as its own element-wise pass (it used to be densified and rejected above 11
    qubits at launch time); its wires may sit on rank bits of a sharded state."""
    import helpers as H
    n = 14
    rng = np.random.default_rng(21)
    for k in (9, 12):
        qs = tuple(int(x) for x in rng.choice(n, k, replace=False))
        d = np.exp(1j * rng.uniform(0, 6.28, size=1 << k))
        circ = H.random_circuit(n, 10, rng) + [(np.diag(d), qs)] + H.random_circuit(n, 10, rng)
        psi = H.rand_state(n, k)
        ref = H.run_oracle(psi.copy(), circ)
        x = psi.copy()
        H.run_emulator(x, n, circ)
        H.assert_close(x, ref)
        y = psi.copy()
        _, _, _, l2p = H.run_emulator_sharded(y, n, 4, circ, remap_pull=True)
        H.assert_close(H.unpermute(y, l2p), ref)


def test_too_wide_dense_gate_fails_before_anything_runs():
    import helpers as H
    n = 14
    rng = np.random.default_rng(3)
    u = H.rand_unitary(1, rng)
    big = np.eye(1 << 12, dtype=np.complex128)
    big[0, 0] = 0; big[0, -1] = 1; big[-1, -1] = 0; big[-1, 0] = 1      # mixes all 12 qubits
    psi = H.rand_state(n, 1)
    x = psi.copy()
    with pytest.raises(RuntimeError, match="more than 11 mixing qubits"):
        H.run_emulator(x, n, [(u, (0,)), (big, tuple(range(12)))])
    assert (x == psi).all()      # the schedule is rejected as a whole: the first gate has not been applied


# ---------------------------------------------------------------- swap routing
def _swap_heavy_circuit(n, n_gates, rng):
    """Random circuit in which every third gate or so is an exact SWAP (also back-to-back and repeated pairs)."""
    from qvm_b200 import gates as G
    base = random_circuit(n, n_gates, rng, max_dense=3)
    out = []
    for g in base:
        out.append(g)
        while n >= 2 and rng.integers(0, 3) == 0:
            a, b = rng.choice(n, 2, replace=False)
            out.append((G.gate_matrix("SWAP"), (int(a), int(b))))
    return out


@pytest.mark.parametrize("seed", range(24))
def test_swap_routing_matches_oracle_and_ends_canonical(seed):
    """Absorbed SWAP gates + the permutation executed by the passes' write-backs: same state as the oracle, canonical layout
    (l2p = identity), whatever the tile size; also from a non-canonical starting layout (the tape then canonicalises it)."""
    rng = np.random.default_rng(4200 + seed)
    n = int(rng.integers(2, 15))
    tile_bits = int(rng.integers(3, 13))
    circ = _swap_heavy_circuit(n, int(rng.integers(4, 50)), rng)
    a = rand_state(n, seed)
    b = a.copy()
    _, desc, l2p = run_emulator(a, n, circ, fuse=True, tile_bits=tile_bits, route_swaps=1, reg_bits=(3, 4, 0)[seed % 3])
    run_oracle(b, circ)
    assert list(l2p) == list(range(n)), desc
    assert_close(a, b)
    # cost-model choice: same answer, canonical too
    c = rand_state(n, seed)
    _, _, l2p = run_emulator(c, n, circ, fuse=True, tile_bits=tile_bits, route_swaps=-1)
    assert list(l2p) == list(range(n))
    assert_close(c, b)
    # start from a permuted layout (as left behind by an absorb_swaps run)
    perm = rng.permutation(n).astype(np.int32)
    logical = rand_state(n, 77 + seed)
    idx = np.arange(1 << n)
    phys = np.zeros_like(idx)
    for q in range(n):
        phys |= ((idx >> q) & 1) << int(perm[q])
    d = np.empty_like(logical)
    d[phys] = logical                      # physical bit perm[q] holds logical qubit q
    e = logical.copy()
    _, desc, l2p = run_emulator(d, n, circ, fuse=True, tile_bits=tile_bits, route_swaps=1, l2p_in=perm)
    run_oracle(e, circ)
    assert list(l2p) == list(range(n)), desc
    assert_close(d, e)


def test_swap_routing_index_tracer():
    """dqvm's debug wavefunction psi_i = i (dqvm/tests/program-tests.lisp:14-19) through permutation-only circuits: routed
    SWAP / CNOT / X networks move integer labels exactly."""
    from qvm_b200 import gates as G
    n = 13
    rng = np.random.default_rng(5)
    circ = []
    for _ in range(60):
        a, b = (int(x) for x in rng.choice(n, 2, replace=False))
        circ.append([(G.gate_matrix("SWAP"), (a, b)), (G.gate_matrix("SWAP"), (a, b)), (G.gate_matrix("CNOT"), (a, b)),
                     (G.gate_matrix("X"), (a,))][int(rng.integers(0, 4))])
    psi = np.arange(1 << n).astype(np.complex128)
    ref = run_oracle(psi.copy(), circ)
    _, desc, l2p = run_emulator(psi, n, circ, fuse=True, tile_bits=6, route_swaps=1)
    assert list(l2p) == list(range(n)), desc
    assert np.array_equal(psi, ref)


def test_swap_routing_runs_qft30_in_five_passes():
    """The 30-qubit QFT (15 trailing SWAPs = a bit reversal): 4 gate passes + 2 permutation passes without routing, 5 passes
    with it -- the schedule the bench runs.  Schedule only (no 16 GiB state here)."""
    from qvm_b200 import qvm
    t = qvm.Tape(30, CC.qft_circuit(range(30)), fuse=True)
    try:
        assert t.info()["passes"] == 5, t.describe()
        assert "routed_transpositions=15" in t.describe()
    finally:
        t.close()
