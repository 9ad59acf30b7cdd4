"""Multi-GPU parity (-m gpu, needs >= 2 GPUs on the box): two ranks over NCCL, each owning one shard;
the tile kernel reaches the peer shard over NVLink (CUDA IPC).  Compared with the oracle on rank 0."""
import os
import socket
import sys
import traceback

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
HERE = os.path.dirname(os.path.abspath(__file__))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, q, inplace):
    try:
        if inplace:
            os.environ["QVM_REMAP_INPLACE"] = "1"     # remaps as in-place peer passes of the tile kernel
        else:
            os.environ.pop("QVM_REMAP_INPLACE", None)  # remaps as out-of-place pulls into the alternate buffer
        sys.path.insert(0, HERE)
        sys.path.insert(0, os.path.dirname(HERE))
        import torch
        import torch.distributed as dist
        torch.cuda.set_device(rank)
        dist.init_process_group("nccl", init_method=f"tcp://127.0.0.1:{port}", rank=rank, world_size=world,
                                device_id=torch.device("cuda", rank))
        import helpers
        from oracle import oracle as O
        from qvm_b200 import circuits as CC
        from qvm_b200.dist import ShardedState

        n = 19 + (world.bit_length() - 1)
        st = ShardedState(n, dist, device=rank)
        assert st.engine.remap_pull == (not inplace)
        psi = helpers.rand_state(n, 21)
        st.scatter_logical(psi)
        rng = np.random.default_rng(5)
        circ = CC.qft_circuit(range(n)) + helpers.random_circuit(n, 60, rng, max_dense=3) + CC.random_layer_circuit(n, 2, 1)
        # absorb: exact SWAP gates as relabelings of the layout (the default for sharded states) or executed as gates
        for fuse, absorb in ((True, False), (False, True), (True, True)):
            st.set_zero_state()
            st.scatter_logical(psi)
            st.apply_gates(circ, fuse=fuse, absorb_swaps=absorb)
            got = st.gather_logical()
            if rank == 0:
                ref = helpers.run_oracle(psi.copy(), circ)
                helpers.assert_close(got, ref)
        assert st.peer_steps >= 1
        # straight from the collective reset: the shards of ranks != 0 are known to hold only zeros (never written in pull mode),
        # the first exchange pass does not fetch them
        zero = np.zeros(1 << n, dtype=np.complex128)
        zero[0] = 1.0
        zref = helpers.run_oracle(zero.copy(), circ)
        for fuse in (True, False):
            st.set_zero_state()
            st.apply_gates(circ, fuse=fuse)
            got = st.gather_logical()
            if rank == 0:
                helpers.assert_close(got, zref)
        st.set_zero_state()
        assert abs(st.norm2() - 1) == 0 and st.prob_excited(n - 1) == 0.0       # readers of a never-written shard
        st.set_zero_state()
        st.scatter_logical(psi)
        st.apply_gates(circ, fuse=True)
        assert abs(st.norm2() - 1) < 1e-12
        ref = helpers.run_oracle(psi.copy(), circ)
        for qb in (0, 7, n - 2, n - 1):
            assert abs(st.prob_excited(qb) - O.prob_excited(ref, qb)) < 1e-12
        u = np.random.default_rng(7).random(20000)
        idx = st.sample(u, strict=False).astype(np.int64)
        # bit-exact against the oracle's restatement of the sharded summation order (physical, rank-major indices)
        for strict in (False, True):
            want = O.sample_tree_sharded(st.gather_physical(), world, u, strict)
            assert (st.sample_physical(u, strict) == want).all()
        probs = np.abs(ref) ** 2
        assert probs[idx].min() > 0
        # coarse distribution check on the top 3 logical qubits
        for qb in (n - 1, n - 2, 0):
            frac = ((idx >> qb) & 1).mean()
            assert abs(frac - O.prob_excited(ref, qb)) < 0.02
        lay = st.layout()
        for qb in (int(np.argmax(lay)), int(np.argmin(lay))):
            p1 = O.prob_excited(ref, qb)
            bit = st.measure(qb, 0.37)
            assert bit == (1 if 0.37 <= p1 else 0)
            O.force_measurement(ref, qb, bit, p1)
            got = st.gather_logical()
            if rank == 0:
                helpers.assert_close(got, ref)
        st.close()
        dist.destroy_process_group()
        q.put((rank, "ok"))
    except Exception:
        q.put((rank, traceback.format_exc()))


@pytest.mark.parametrize("inplace", [False, True])
@pytest.mark.parametrize("world", [2, 4, 8])
def test_sharded_state_over_nvlink(world, inplace):
    import torch
    if torch.cuda.device_count() < world:
        pytest.skip(f"needs {world} GPUs")
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q, inplace)) for r in range(world)]
    for p in procs:
        p.start()
    results = [q.get(timeout=600) for _ in procs]
    for p in procs:
        p.join(timeout=60)
    for rank, msg in results:
        assert msg == "ok", f"rank {rank}: {msg}"


@pytest.mark.parametrize("want_alt", [True, False])
def test_single_process_shard_group(want_alt):
    """One host process, one shard per device (qvmcuda_shard_attach_local: peer access instead of IPC handles) -- the shape of a
    single Lisp image driving several GPUs.  Same schedules and exchange passes as the one-process-per-GPU path."""
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    import helpers
    from oracle import oracle as O
    from qvm_b200 import circuits as CC
    from qvm_b200.dist import LocalShardGroup
    world = 4 if torch.cuda.device_count() >= 4 else 2
    n = 18 + (world.bit_length() - 1)
    grp = LocalShardGroup(n, list(range(world)), want_alt=want_alt)
    psi = helpers.rand_state(n, 33)
    rng = np.random.default_rng(9)
    circ = CC.qft_circuit(range(n)) + helpers.random_circuit(n, 50, rng, max_dense=3)
    ref = helpers.run_oracle(psi.copy(), circ)
    for fuse, absorb in ((True, True), (True, False), (False, True)):
        grp.scatter_logical(psi)
        grp.apply_gates(circ, fuse=fuse, absorb_swaps=absorb)
        helpers.assert_close(grp.gather_logical(), ref)
    assert grp.peer_steps >= 1
    assert abs(grp.norm2() - 1) < 1e-12
    for qb in (0, 5, n - 1):
        assert abs(grp.prob_excited(qb) - O.prob_excited(ref, qb)) < 1e-12
    # from the reset state: zero shards are not fetched by the first exchange
    zero = np.zeros(1 << n, dtype=np.complex128)
    zero[0] = 1.0
    grp.set_zero_state()
    grp.apply_gates(circ)
    helpers.assert_close(grp.gather_logical(), helpers.run_oracle(zero, circ))
    # twice in a row from the layout the first run leaves behind
    grp.scatter_logical(psi)
    grp.apply_gates(circ)
    grp.apply_gates(circ)
    helpers.assert_close(grp.gather_logical(), helpers.run_oracle(ref.copy(), circ))
    grp.close()
