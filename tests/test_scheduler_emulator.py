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
