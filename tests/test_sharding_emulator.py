"""Host logic of the multi-GPU path: the sharded schedules (remap passes, peer tiles, rank-dependent
diagonals/controls) interpreted on the CPU for 2/4/8 emulated ranks must reproduce the oracle.
Mirrors dqvm's own strategy of testing the address algebra without MPI ranks
(dqvm/tests/distributed-qvm-tests.lisp:52-98, dqvm/tests/program-tests.lisp:61-113)."""
import numpy as np
import pytest

from helpers import assert_close, rand_state, random_circuit, run_emulator_sharded, run_oracle, unpermute
from qvm_b200 import circuits as CC


@pytest.mark.parametrize("world", [2, 4, 8])
@pytest.mark.parametrize("fuse", [True, False])
@pytest.mark.parametrize("pull", [False, True])
@pytest.mark.parametrize("absorb", [False, True])
def test_sharded_qft(world, fuse, pull, absorb):
    n, tile_bits = 13, 7
    circ = CC.qft_circuit(range(n))
    a = rand_state(n)
    ref = run_oracle(a.copy(), circ)
    steps, peer_steps, desc, l2p = run_emulator_sharded(a, n, world, circ, fuse=fuse, tile_bits=tile_bits, remap_pull=pull,
                                                        absorb_swaps=absorb)
    assert peer_steps >= 1, desc
    assert_close(unpermute(a, l2p), ref)


@pytest.mark.parametrize("seed", range(8))
def test_sharded_random_circuits(seed):
    rng = np.random.default_rng(900 + seed)
    world = int(rng.choice([2, 4, 8]))
    n = int(rng.integers(12, 15))
    circ = random_circuit(n, 50, rng, max_dense=4)
    a = rand_state(n, seed)
    ref = run_oracle(a.copy(), circ)
    steps, peer_steps, desc, l2p = run_emulator_sharded(a, n, world, circ, fuse=True, tile_bits=7, remap_pull=bool(seed & 1),
                                                        absorb_swaps=bool(seed & 2))
    assert_close(unpermute(a, l2p), ref)


@pytest.mark.parametrize("pull", [False, True])
def test_index_tracer_through_remaps(pull):
    """dqvm's debug wavefunction psi_i = i (dqvm/tests/program-tests.lisp:14-19): permutation-only circuits
    move integer labels exactly, so any address-algebra slip shows up as a wrong integer."""
    from qvm_b200 import gates as G
    n, world = 12, 4
    rng = np.random.default_rng(3)
    circ = []
    for _ in range(40):
        a, b, c = (int(x) for x in rng.choice(n, 3, replace=False))
        circ.append([(G.gate_matrix("SWAP"), (a, b)), (G.gate_matrix("CNOT"), (a, b)), (G.gate_matrix("CCNOT"), (a, b, c)),
                     (G.gate_matrix("X"), (a,))][int(rng.integers(0, 4))])
    psi = np.arange(1 << n).astype(np.complex128)
    ref = run_oracle(psi.copy(), circ)
    _, peer_steps, desc, l2p = run_emulator_sharded(psi, n, world, circ, fuse=True, tile_bits=6, remap_pull=pull)
    assert peer_steps >= 1 and ((("REMAP(pull)" in desc) or ("PULL+TILE" in desc)) == pull), desc
    assert np.array_equal(unpermute(psi, l2p), ref)


def test_diagonal_gates_on_global_qubits_need_no_exchange():
    from qvm_b200 import gates as G
    n, world = 12, 8
    circ = [(G.gate_matrix("H"), (q,)) for q in range(9)]
    circ += [(G.gate_matrix("CZ"), (11, 3)), (G.gate_matrix("CPHASE", [0.3]), (10, 11)), (G.gate_matrix("RZ", [0.7]), (9,)),
             (G.gate_matrix("CNOT"), (10, 2)), (G.gate_matrix("T"), (11,))]
    a = rand_state(n)
    ref = run_oracle(a.copy(), circ)
    steps, peer_steps, desc, l2p = run_emulator_sharded(a, n, world, circ, fuse=True, tile_bits=6)
    assert peer_steps == 0, desc
    assert_close(unpermute(a, l2p), ref)


@pytest.mark.parametrize("world", [2, 8])
@pytest.mark.parametrize("pull", [False, True])
def test_repeated_application_carries_the_layout(world, pull):
    """bench.py's sharded step applies the same circuit again and again on whatever layout the previous run left
    behind (SWAPs on rank bits are relabelings, remaps move qubits): three QFTs in a row, layout carried over,
    must equal three oracle QFTs."""
    n = 13
    circ = CC.qft_circuit(range(n))
    a = rand_state(n, 11)
    ref = a.copy()
    l2p = np.arange(n, dtype=np.int32)
    relabeled = 0
    for it in range(3):
        ref = run_oracle(ref, circ)
        # the shards hold the state in PHYSICAL order; hand the emulator the current layout and take the new one back
        import ctypes as C
        from helpers import emulator, flatten_circuit
        ks, qf, mf = flatten_circuit(circ)
        desc = C.create_string_buffer(1 << 16)
        emu = emulator()
        emu.qvtest_run_sharded.restype = C.c_int
        emu.qvtest_set_remap_pull(int(pull))
        emu.qvtest_set_reg_bits(0)
        p = lambda x: x.ctypes.data_as(C.c_void_p)
        rc = emu.qvtest_run_sharded(p(a), n, world, len(circ), p(ks), p(qf), p(mf), 1, 7, 0, p(l2p), desc, len(desc))
        assert rc >= 0, desc.value.decode()
        relabeled += desc.value.decode().count("relabeled_swaps")
        assert_close(unpermute(a, l2p), ref)
    assert relabeled >= 1      # at least one run met a SWAP on a rank bit
    assert not (l2p == np.arange(n)).all()


def test_remap_hoisting_folds_the_exchange_into_a_gate_pass():
    """Fused pulls: an exchange a later gate needs is done as part of the current pass when that pass then holds everything it
    held and more (an exchange pass is NVLink-bound whatever it computes).  Hoisted schedules reproduce the oracle; the QFT on
    8 emulated ranks needs fewer steps than with hoisting off."""
    from helpers import emulator
    hoisted = 0
    for world, n, tb in [(2, 13, 7), (4, 14, 6), (8, 16, 8), (8, 14, 7)]:
        circ = CC.qft_circuit(range(n))
        a = rand_state(n)
        ref = run_oracle(a.copy(), circ)
        steps, peer_steps, desc, l2p = run_emulator_sharded(a, n, world, circ, fuse=True, tile_bits=tb, remap_pull=True, absorb_swaps=True)
        assert_close(unpermute(a, l2p), ref)
        hoisted += "hoisted_remaps" in desc
    assert hoisted >= 2
    rng = np.random.default_rng(11)
    for seed in range(12):
        world = int(rng.choice([2, 4, 8]))
        n = int(rng.integers(12, 16))
        circ = random_circuit(n, 80, rng, max_dense=4)
        a = rand_state(n, seed)
        ref = run_oracle(a.copy(), circ)
        _, _, desc, l2p = run_emulator_sharded(a, n, world, circ, fuse=True, tile_bits=int(rng.integers(6, 9)), remap_pull=True,
                                               absorb_swaps=bool(seed & 1))
        assert_close(unpermute(a, l2p), ref)
