"""The reference's benchmark inputs (bench/*.quil, committed as tests/golden/bench_circuits.json by
tests/golden/make_golden.py): oracle known answers on CPU, CUDA parity on GPU."""
import numpy as np
import pytest

from helpers import assert_close, load_bench_circuit, rand_state, run_emulator, run_oracle


def test_20H_known_answer_oracle_and_schedule():
    circ, _, n = load_bench_circuit("20H")
    assert n == 20 and len(circ) == 20
    from oracle import oracle as O
    psi = run_oracle(O.zero_state(n), circ)
    np.testing.assert_allclose(psi, 2.0 ** -10, rtol=1e-13)      # all amplitudes = 2^-10
    a = O.zero_state(n)
    steps, desc, _ = run_emulator(a, n, circ, fuse=True)
    assert steps <= 3, desc
    assert_close(a, psi)


def test_5x4x25_shape():
    circ, _, n = load_bench_circuit("5x4x25")
    assert n == 20 and len(circ) == 231           # 20 H, 97 CPHASE(pi), 44 RX, 50 RY, 20 T
    diag = sum(1 for m, _ in circ if np.count_nonzero(m - np.diag(np.diag(m))) == 0)
    assert diag == 117


@pytest.mark.gpu
def test_5x4x25_parity_and_sampling():
    from oracle import oracle as O
    from qvm_b200 import qvm
    circ, _, n = load_bench_circuit("5x4x25")
    ref = run_oracle(O.zero_state(n), circ)
    for fuse in (True, False):
        vec = qvm.DeviceVector(1 << n)
        vec.set_zero_state()
        vec.apply_gates(circ, fuse=fuse)
        assert_close(vec.download(), ref)
        if fuse:
            u = np.random.default_rng(2024).random(100000)          # 10^5 shots
            got = vec.sample(u, strict=False)
            assert (got == O.sample_tree(vec.download(), u, False)).all()
        vec.close()


@pytest.mark.gpu
def test_entangle_25_ghz_and_measures():
    from qvm_b200 import qvm
    circ, measures, n = load_bench_circuit("entangle-25")
    assert n == 25 and [m[0] for m in measures] == [0, 23]
    for seed in range(4):
        q = qvm.make_qvm(n, seed=seed)
        q.state.vec.apply_gates(circ, fuse=True)
        head, tail = q.state.vec.download(0, 2), q.state.vec.download((1 << n) - 2, 2)
        assert abs(head[0] - 2 ** -0.5) < 1e-14 and abs(tail[1] - 2 ** -0.5) < 1e-14 and abs(head[1]) == 0
        b0 = q.measure(0)
        b23 = q.measure(23)          # deterministic second outcome: exercises the p1 = 0 / p1 = 1 branches
        assert b0 == b23
        idx = (1 << n) - 1 if b0 else 0
        assert abs(q.state.vec.download(idx, 1)[0] - 1) < 1e-13
        assert abs(q.state.vec.norm2() - 1) < 1e-13


@pytest.mark.gpu
def test_qaoa_8q_dense_256x256_gates():
    import os
    from oracle import oracle as O
    from qvm_b200 import gates as G
    from qvm_b200 import qvm
    z = np.load(os.path.join(os.path.dirname(__file__), "golden", "bench_qaoa_8q.npz"))
    circ = [(G.gate_matrix("H"), (q,)) for q in range(8)]
    circ += [(z["UC"], tuple(range(8))), (z["UB"], tuple(range(8)))]          # bench/qaoa_8q.quil:526-527
    ref = run_oracle(O.zero_state(8), circ)
    vec = qvm.DeviceVector(1 << 8)
    vec.set_zero_state()
    vec.apply_gates(circ, fuse=True)
    assert_close(vec.download(), ref)
    assert abs(vec.norm2() - 1) < 1e-12
    vec.close()
