"""TEST-ONLY shard engine for qvm_b200.dist.ShardedState: the rank's shard is a numpy array in POSIX
shared memory, peer passes reach the other ranks' shards through the same shared memory (standing in
for NVLink peer access), steps are interpreted by tests/support/qv_emulator.cpp.  Used by the gloo
world-size-2 tests; never imported by the product."""
from __future__ import annotations

import ctypes as C
from multiprocessing import shared_memory

import numpy as np

import helpers


class EmuShardEngine:
    def __init__(self, n_local, rank, world, dist, tag, tile_bits=6):
        self.n_local, self.rank, self.world, self.tile_bits = n_local, rank, world, tile_bits
        self.n_total = n_local + (world.bit_length() - 1)
        self.emu = helpers.emulator()
        self.emu.qvtest_shard_compile.restype = C.c_void_p
        self.shm = shared_memory.SharedMemory(create=True, size=16 << n_local, name=f"qvemu_{tag}_{rank}")
        dist.barrier()
        self.peers_shm = [self.shm if r == rank else shared_memory.SharedMemory(name=f"qvemu_{tag}_{r}") for r in range(world)]
        self.shards = [np.ndarray(1 << n_local, dtype=np.complex128, buffer=s.buf) for s in self.peers_shm]
        self.local = self.shards[rank]
        self.local[:] = 0
        self.l2p = np.arange(self.n_total, dtype=np.int32)
        self.ptrs = (C.c_void_p * world)(*[a.ctypes.data for a in self.shards])

    def compile(self, gates, fuse=True, absorb_swaps=False):
        ks, qf, mf = helpers.flatten_circuit(gates)
        err = C.create_string_buffer(512)
        p = lambda a: a.ctypes.data_as(C.c_void_p)
        t = self.emu.qvtest_shard_compile(self.n_total, self.world, self.rank, len(gates), p(ks), p(qf), p(mf), int(fuse),
                                          self.tile_bits, int(absorb_swaps), p(self.l2p), err, len(err))
        if not t:
            raise RuntimeError(err.value.decode())
        return C.c_void_p(t)

    def set_zero_ranks(self, mask):
        """The host's claim "these ranks hold only zeros" (qvmcuda_shard_set_zero_ranks): the emulator does not use it, it CHECKS
        it at the first exchange step -- the moment the CUDA path would skip fetching those shards."""
        self.zero_ranks = int(mask)

    def num_steps(self, tape): return self.emu.qvtest_shard_num_steps(tape)
    def step_flags(self, tape, i): return self.emu.qvtest_shard_step_flags(tape, i)
    def run_step(self, tape, i):
        if getattr(self, "zero_ranks", 0) and (self.step_flags(tape, i) & 1):
            for r in range(self.world):
                if self.zero_ranks >> r & 1:
                    assert not self.shards[r].any(), f"rank {r} was declared all-zero but holds amplitudes"
            self.zero_ranks = 0
        self.emu.qvtest_shard_run_step(tape, i, self.ptrs, self.rank)
    def commit(self, tape): self.emu.qvtest_shard_l2p(tape, self.l2p.ctypes.data_as(C.c_void_p))
    def free_tape(self, tape): self.emu.qvtest_shard_free(tape)
    def synchronize(self): pass
    def layout(self): return self.l2p.copy()

    def set_basis_local(self, index):
        self.local[:] = 0
        self.local[index] = 1
        self.l2p = np.arange(self.n_total, dtype=np.int32)

    def clear(self):
        self.local[:] = 0
        self.l2p = np.arange(self.n_total, dtype=np.int32)

    def _pbit(self, q): return int(self.l2p[q])
    def norm2(self): return float((np.abs(self.local) ** 2).sum())

    def prob_excited(self, q):
        p = self._pbit(q)
        if p >= self.n_local:
            return self.norm2() if (self.rank >> (p - self.n_local)) & 1 else 0.0
        idx = np.arange(self.local.size)
        return float((np.abs(self.local[(idx >> p) & 1 == 1]) ** 2).sum())

    def collapse(self, q, keep, inv):
        p = self._pbit(q)
        if p >= self.n_local:
            self.local[:] = self.local * inv if ((self.rank >> (p - self.n_local)) & 1) == keep else 0
            return
        idx = np.arange(self.local.size)
        m = ((idx >> p) & 1) == keep
        self.local[~m] = 0
        self.local[m] *= inv

    def sample_total(self):
        from oracle import oracle as O
        return O.sample_tree_total(np.ascontiguousarray(self.local))

    def sample_local(self, u, strict, base):
        from oracle import oracle as O
        return O.sample_tree_base(np.ascontiguousarray(self.local), u, strict, base)

    def download(self): return np.array(self.local)
    def upload(self, a): self.local[:] = a

    def close(self):
        del self.local, self.shards, self.ptrs
        for r, s in enumerate(self.peers_shm):
            s.close()
        try:
            self.shm.unlink()
        except FileNotFoundError:
            pass
