"""Regression guards for the SHAPE of the schedules the bench lines run (no GPU, no state: schedules only).  The numbers of HBM
passes / steps per circuit are what the measured times in BASELINE.md rest on."""
import ctypes as C

import numpy as np
import pytest

from qvm_b200 import _lib as L
from qvm_b200 import circuits as CC
from qvm_b200 import qvm


@pytest.mark.parametrize("n,passes", [(20, 3), (24, 4), (26, 4), (28, 5), (30, 5), (32, 6)])
def test_qft_pass_counts_single_device(n, passes):
    """Fused QFT on one device, canonical layout in and out (swap routing decides between executing and routing the SWAPs)."""
    t = qvm.Tape(n, CC.qft_circuit(range(n)), fuse=True)
    try:
        assert t.info()["passes"] <= passes, t.describe()
        assert t.describe().rstrip().splitlines()[-1].split()[1:] == [str(q) for q in range(n)]      # l2p: identity
    finally:
        t.close()


def _plan(n, world, rank, gates, l2p, pull=1):
    ks, qf, mf = L.flatten_gates(gates)
    h = C.c_void_p()
    L.check(L.lib().qvmcuda_shard_plan(n, world, rank, pull, L.ptr(l2p), len(gates), L.ptr(ks), L.ptr(qf), L.ptr(mf),
                                       L.FUSE | L.ABSORB_SWAPS, C.byref(h)))
    steps = C.c_int()
    L.check(L.lib().qvmcuda_tape_num_steps(h, C.byref(steps)))
    peer = 0
    for i in range(steps.value):
        f = C.c_uint32()
        L.check(L.lib().qvmcuda_tape_step_flags(h, i, C.byref(f)))
        peer += f.value & 1
    L.lib().qvmcuda_tape_destroy(h)
    return steps.value, peer


@pytest.mark.parametrize("world", [2, 4, 8])
def test_sharded_qft_is_three_local_passes_and_one_exchange(world):
    """The N > 1 bench lines: QFT on 32 + log2(world) qubits, 32 local qubits per rank, fused pulls, every SWAP a relabeling, the
    exchange hoisted into a gate pass.  From the canonical layout (a run from |0...0>) two exchanges are needed, in steady state
    one; every rank plans the same number of steps, and the layouts cycle with period two."""
    n = 32 + world.bit_length() - 1
    gates = CC.qft_circuit(range(n))
    per_rank = []
    for rank in (0, world - 1):
        l2p = np.arange(n, dtype=np.int32)
        seen = []
        for run in range(5):
            steps, peer = _plan(n, world, rank, gates, l2p)
            seen.append((steps, peer, tuple(l2p)))
        per_rank.append([(s, p) for s, p, _ in seen])
        assert seen[0][:2] == (4, 2)                       # from the canonical layout
        for s, p, _ in seen[1:]:
            assert (s, p) == (4, 1)                        # steady state: 3 local passes + 1 exchange
        assert seen[1][2] == seen[3][2] and seen[2][2] == seen[4][2]      # layouts repeat with period two
    assert per_rank[0] == per_rank[1]


def test_plan_rejects_bad_geometry():
    l2p = np.arange(10, dtype=np.int32)
    g = CC.qft_circuit(range(10))
    ks, qf, mf = L.flatten_gates(g)
    h = C.c_void_p()
    for world, rank in ((3, 0), (4, 4), (0, 0)):
        assert L.lib().qvmcuda_shard_plan(10, world, rank, 1, L.ptr(l2p), len(g), L.ptr(ks), L.ptr(qf), L.ptr(mf), L.FUSE, C.byref(h)) != 0
