"""The N > 1 path on CPU: two processes, torch.distributed `gloo`, ShardedState driving the TEST-ONLY
emulator engine (shards in POSIX shared memory).  Checks the protocol the GPUs follow: identical
schedules on every rank, barriers around peer passes, scalar reductions, sharded sampling/measurement."""
import os
import socket
import sys
import traceback

import numpy as np
import pytest

HERE = os.path.dirname(os.path.abspath(__file__))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, q):
    try:
        sys.path.insert(0, HERE)
        sys.path.insert(0, os.path.dirname(HERE))
        sys.path.insert(0, os.path.join(HERE, "support"))
        import torch.distributed as dist
        dist.init_process_group("gloo", init_method=f"tcp://127.0.0.1:{port}", rank=rank, world_size=world)
        import helpers
        from emu_engine import EmuShardEngine
        from oracle import oracle as O
        from qvm_b200 import circuits as CC
        from qvm_b200 import gates as G
        from qvm_b200.dist import ShardedState

        n = 11
        n_local = n - (world.bit_length() - 1)
        st = ShardedState(n, dist, engine_factory=lambda: EmuShardEngine(n_local, rank, world, dist, tag=port))
        psi = helpers.rand_state(n, 21)
        st.scatter_logical(psi)
        rng = np.random.default_rng(5)
        circ = CC.qft_circuit(range(n)) + helpers.random_circuit(n, 40, rng, max_dense=3)
        st.apply_gates(circ, fuse=True)
        ref = helpers.run_oracle(psi.copy(), circ)
        got = st.gather_logical()
        helpers.assert_close(got, ref)
        assert st.peer_steps >= 1
        assert abs(st.norm2() - 1) < 1e-12
        for qb in range(n):
            assert abs(st.prob_excited(qb) - O.prob_excited(ref, qb)) < 1e-12
        # sampling: every rank returns the same logical indices, distributed like |psi|^2
        u = np.random.default_rng(7).random(4000)
        idx = st.sample(u, strict=False)
        # bit-exact against the oracle's restatement of the sharded summation order (physical, rank-major indices)
        for strict in (False, True):
            want = O.sample_tree_sharded(st.gather_physical(), world, u, strict)
            assert (st.sample_physical(u, strict) == want).all()
        probs = np.abs(ref) ** 2
        hist = np.bincount(idx.astype(np.int64), minlength=1 << n) / u.size
        assert np.abs(hist - probs).sum() < 0.9          # coarse: 2048 bins, 4000 shots
        assert probs[idx.astype(np.int64)].min() > 0
        # raw dump in dqvm's file layout (ordered amplitudes, 16 bytes each) and back
        import tempfile
        path = os.path.join(tempfile.gettempdir(), f"qvm_dump_{port}.bin")
        st.save_wavefunction(path)
        if rank == 0:
            helpers.assert_close(np.fromfile(path, dtype=np.complex128), ref)
        dist.barrier()
        lay_before = st.layout().copy()
        st.load_wavefunction(path)
        helpers.assert_close(st.gather_logical(), ref)
        dist.barrier()
        if rank == 0:
            os.remove(path)
        # straight from the collective reset: ranks != 0 are declared all-zero until the first exchange (the emulator engine checks
        # the claim where the CUDA path relies on it); an upload in between must withdraw it
        zero = np.zeros(1 << n, dtype=np.complex128)
        zero[0] = 1.0
        st.set_zero_state()
        assert st._zero_ranks == ((1 << world) - 1) & ~1
        st.apply_gates(circ, fuse=True)
        assert st._zero_ranks == 0
        helpers.assert_close(st.gather_logical(), helpers.run_oracle(zero.copy(), circ))
        st.set_zero_state()
        st.scatter_logical(psi)
        assert st._zero_ranks == 0
        # back to the layout the run left behind (the checks below pick a rank-selecting qubit)
        st.set_zero_state()
        st.scatter_logical(psi)
        st.apply_gates(circ, fuse=True)
        assert (st.layout() == lay_before).all()
        # measurement of a qubit that currently selects the rank and of a local one
        lay = st.layout()
        for qb in (int(np.argmax(lay)), int(np.argmin(lay))):
            p1 = O.prob_excited(ref, qb)
            bit = st.measure(qb, 0.37)
            assert bit == (1 if 0.37 <= p1 else 0)
            O.force_measurement(ref, qb, bit, p1)
            helpers.assert_close(st.gather_logical(), ref)
        st.close()
        dist.destroy_process_group()
        q.put((rank, "ok"))
    except Exception:
        q.put((rank, traceback.format_exc()))


@pytest.mark.timeout(300)
def test_two_rank_gloo_sharded_state():
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    results = [q.get(timeout=240) for _ in procs]
    for p in procs:
        p.join(timeout=30)
    for rank, msg in results:
        assert msg == "ok", f"rank {rank}: {msg}"
