"""Multi-GPU sharding of one state vector: one process per GPU (torchrun), state split on its top
log2(P) physical index bits (the layout of dqvm, dqvm/src/global-addresses.lisp:99-151, with one
block = the whole shard).

torch.distributed is plumbing only: it carries the 64-byte CUDA IPC handles of the shards, the
barriers around peer passes and a few scalars.  The amplitude traffic itself is done by the tile kernel
(qv_tile_kernel<.., PEERS=true>), which loads and stores peer shards directly over NVLink while it
applies the gates of the pass -- remap and compute in ONE kernel, no NCCL on the data path.

A shard ENGINE needs: shard_compile / num_steps / step_flags / run_step / commit / layout /
synchronize plus the local reductions.  `CudaShardEngine` is the product; tests/support has a CPU
emulator engine (shared-memory shards) used only by the gloo world-size-2 tests.
"""
from __future__ import annotations

import ctypes as C
import json
import math
import os
import time
from typing import List, Sequence

import numpy as np

STEP_PEER = 1
STEP_REMAP = 2


def _log2(n: int) -> int:
    b = n.bit_length() - 1
    if n < 1 or (1 << b) != n:
        raise ValueError("world size must be a power of two")
    return b


class CudaShardEngine:
    """One rank's shard on its GPU, through libqvmcuda's C ABI."""

    def __init__(self, n_local: int, rank: int, world: int, device: int, dist):
        from . import _lib as L
        from .qvm import DeviceVector
        self.L = L
        self.vec = DeviceVector(1 << n_local, device)
        handle = np.zeros(64, dtype=np.uint8)
        L.check(L.lib().qvmcuda_shard_export(self.vec.handle, L.ptr(handle)))
        handles: List[bytes] = [None] * world
        dist.all_gather_object(handles, handle.tobytes())
        allh = np.frombuffer(b"".join(handles), dtype=np.uint8).copy()
        L.check(L.lib().qvmcuda_shard_attach(self.vec.handle, rank, world, L.ptr(allh)))
        self.n_total = n_local + _log2(world)
        # Optional alternate buffer: remaps become out-of-place pulls (each amplitude crosses NVLink once).
        # Every rank must succeed, otherwise all fall back to the in-place exchange through the tile kernel.
        self.remap_pull = False
        if not os.environ.get("QVM_REMAP_INPLACE"):
            alt = np.zeros(64, dtype=np.uint8)
            ok = L.lib().qvmcuda_shard_export_alt(self.vec.handle, L.ptr(alt)) == 0
            alts = [None] * world
            dist.all_gather_object(alts, alt.tobytes() if ok else None)
            if all(a is not None for a in alts):
                allalt = np.frombuffer(b"".join(alts), dtype=np.uint8).copy()
                L.check(L.lib().qvmcuda_shard_attach_alt(self.vec.handle, L.ptr(allalt)))
                self.remap_pull = True

    def compile(self, gates, fuse=True, absorb_swaps=False):
        L = self.L
        ks, qf, mf = L.flatten_gates(gates)
        h = C.c_void_p()
        flags = (L.FUSE if fuse else 0) | (L.ABSORB_SWAPS if absorb_swaps else 0)
        L.check(L.lib().qvmcuda_shard_compile(self.vec.handle, len(gates), L.ptr(ks), L.ptr(qf), L.ptr(mf), flags, C.byref(h)))
        return h

    def num_steps(self, tape) -> int:
        n = C.c_int()
        self.L.check(self.L.lib().qvmcuda_tape_num_steps(tape, C.byref(n)))
        return n.value

    def step_flags(self, tape, i) -> int:
        f = C.c_uint32()
        self.L.check(self.L.lib().qvmcuda_tape_step_flags(tape, i, C.byref(f)))
        return f.value

    def step_info(self, tape, i) -> np.ndarray:
        a = np.zeros(8, dtype=np.int64)
        self.L.check(self.L.lib().qvmcuda_tape_step_info(tape, i, self.L.ptr(a)))
        return a

    def run_step(self, tape, i): self.L.check(self.L.lib().qvmcuda_tape_run_step(self.vec.handle, tape, i))
    def commit(self, tape): self.L.check(self.L.lib().qvmcuda_tape_commit(self.vec.handle, tape))
    def free_tape(self, tape): self.L.lib().qvmcuda_tape_destroy(tape)
    def synchronize(self): self.vec.synchronize()

    def layout(self) -> np.ndarray:
        out = np.zeros(self.n_total, dtype=np.int32)
        self.L.check(self.L.lib().qvmcuda_state_layout(self.vec.handle, self.L.ptr(out), self.n_total))
        return out

    def set_basis_local(self, index): self.vec.set_basis_state(index)
    def clear(self): self.L.check(self.L.lib().qvmcuda_shard_clear(self.vec.handle))
    def set_zero_ranks(self, mask): self.L.check(self.L.lib().qvmcuda_shard_set_zero_ranks(self.vec.handle, int(mask)))
    def norm2(self): return self.vec.norm2()
    def prob_excited(self, q): return self.vec.prob_excited(q)
    def collapse(self, q, keep, inv): self.vec.collapse(q, keep, inv)
    def sample_total(self): return self.vec.sample_total()
    def sample_local(self, u, strict, base): return self.vec.sample_shard(u, strict, base)
    def download(self): return self.vec.download()
    def upload(self, a): self.vec.upload(a)
    def close(self): self.vec.close()


class ShardedState:
    """The state vector of n qubits spread over `world` ranks."""

    def __init__(self, n_qubits: int, dist, engine_factory=None, device: int = 0):
        self.dist = dist
        self.rank = dist.get_rank()
        self.world = dist.get_world_size()
        self.g = _log2(self.world)
        self.n = n_qubits
        self.n_local = n_qubits - self.g
        if engine_factory is None:
            engine_factory = lambda: CudaShardEngine(self.n_local, self.rank, self.world, device, dist)
        self.engine = engine_factory()
        self.peer_steps = 0
        self.steps = 0
        # NVLink accounting of the exchange passes (bench: fraction of the NVLink roofline): seconds spent in passes that
        # read peer shards and the bytes this rank pulled over NVLink in them
        self.peer_seconds = 0.0
        self.peer_bytes = 0.0
        # ranks whose shard is known to hold only zeros (right after a collective reset; every rank keeps the same mask because
        # every method that changes content is collective): their amplitudes are not fetched by the first exchange pass
        self._zero_ranks = 0
        self.set_zero_state()

    # ---- plumbing -------------------------------------------------------------------------------
    def _barrier(self):
        self.engine.synchronize()
        self.dist.barrier()

    def _allreduce_sum(self, x: float) -> float:
        vals = [None] * self.world
        self.dist.all_gather_object(vals, float(x))
        s = 0.0
        for v in vals:          # fixed order: identical on every rank
            s += v
        return s

    # ---- state ------------------------------------------------------------------------------------
    def set_zero_state(self):
        self._barrier()
        if self.rank == 0:
            self.engine.set_basis_local(0)
        else:
            self.engine.clear()
        self._zero_ranks = ((1 << self.world) - 1) & ~1
        # content replaced: layout resets inside the engine (set_basis_state); barrier so that no peer pass
        # starts before every shard is initialised
        self._barrier()

    def apply_gates(self, gates, fuse: bool = True, absorb_swaps: bool = True):
        """absorb_swaps (default on): a sharded state never returns to the canonical qubit layout -- indices are mapped through
        `layout()` on the host, dqvm's "record the permutation instead of undoing it" -- so an exact SWAP gate only exchanges
        the two qubits' entries in the layout; no amplitude moves.  Off: SWAPs between local qubits run as gates (folded into
        a pass's write-back where they trail it), SWAPs touching a rank bit are still relabelings."""
        if hasattr(self.engine, "set_zero_ranks"):
            self.engine.set_zero_ranks(self._zero_ranks)     # before compile: it decides whether an all-zero shard is ever written
        tape = self.engine.compile(gates, fuse=fuse, absorb_swaps=absorb_swaps)
        trace = os.environ.get("QVM_DIST_TRACE") and self.rank == 0
        try:
            n = self.engine.num_steps(tape)
            for i in range(n):
                flags = self.engine.step_flags(tape, i)
                peer = flags & STEP_PEER
                if peer:
                    self._barrier()       # every shard must be complete before anyone reads it remotely
                if trace:
                    self.engine.synchronize()
                if peer or trace:
                    t0 = time.perf_counter()      # the stream is idle here (barrier / synchronize above)
                self.engine.run_step(tape, i)
                if trace:
                    self.engine.synchronize()
                    print(f"[dist] step {i}: {'REMAP' if flags & STEP_REMAP else 'PEER' if peer else 'LOCAL'} "
                          f"{1e3 * (time.perf_counter() - t0):.2f} ms", flush=True)
                if peer:
                    self.engine.synchronize()
                    self.peer_seconds += time.perf_counter() - t0
                    info = self.engine.step_info(tape, i) if hasattr(self.engine, "step_info") else None
                    if info is not None:
                        shard_bytes = 16.0 * (1 << self.n_local)
                        frac = 1.0 - 2.0 ** (-int(info[3]))
                        # a pull pass reads (1 - 2^-pairs) of the new shard from peers; an in-place pass pulls that
                        # share of its tiles and pushes it back (counted once: ingress)
                        self.peer_bytes += shard_bytes * frac
                    self._barrier()       # remote writes must have landed before local work resumes
                    self.peer_steps += 1
                    self._zero_ranks = 0  # after an exchange every shard may hold amplitudes
                self.steps += 1
            self.engine.commit(tape)
        finally:
            self.engine.free_tape(tape)

    def layout(self) -> np.ndarray:
        return self.engine.layout()

    def norm2(self) -> float:
        return self._allreduce_sum(self.engine.norm2())

    def prob_excited(self, q: int) -> float:
        """Sum of the shards' partial probabilities (dqvm: MPI_Allreduce, dqvm/src/measurement.lisp:24-36)."""
        return self._allreduce_sum(self.engine.prob_excited(q))

    def measure(self, q: int, r: float) -> int:
        """MEASURE with the reference's rule (src/measurement.lisp:93-105); r is the host-drawn uniform
        (identical on every rank)."""
        p1 = self.prob_excited(q)
        cbit = 0 if p1 == 0.0 else (1 if r <= p1 else 0)
        inv = 1.0 / math.sqrt(p1) if cbit == 1 else 1.0 / math.sqrt(1.0 - p1)
        self.engine.collapse(q, cbit, inv)
        return cbit

    def sample(self, uniforms: Sequence[float], strict: bool = False) -> np.ndarray:
        """Multi-shot sampling (SAMPLE-WAVEFUNCTION-MULTIPLE-TIMES, src/measurement.lisp:246-288, on a partitioned CDF).
        Returns LOGICAL basis state indices.  The CDF runs in physical index order (rank-major).  Summation order
        (restated as orc_sample_tree_sharded, bit-exact): every shard's mass in its own tree order, the totals added
        left to right, a draw resolved by the first shard whose inclusive prefix hits it, with the exclusive prefix
        as the descent's starting accumulator (NOT subtracted from the draw: that would round differently)."""
        u = np.ascontiguousarray(uniforms, dtype=np.float64)
        phys = self.sample_physical(u, strict)
        l2p = self.layout()
        logical = np.zeros_like(phys)
        for q in range(self.n):
            logical |= ((phys >> np.uint64(int(l2p[q]))) & np.uint64(1)) << np.uint64(q)
        return logical

    def sample_physical(self, u: np.ndarray, strict: bool = False) -> np.ndarray:
        """Physical (rank-major) indices of the draws, identical on every rank."""
        totals = [None] * self.world
        self.dist.all_gather_object(totals, float(self.engine.sample_total()))
        prefix = np.zeros(self.world + 1)
        for r in range(self.world):
            prefix[r + 1] = prefix[r] + totals[r]       # left to right, one rounding per shard
        hit = (prefix[None, 1:] > u[:, None]) if strict else (prefix[None, 1:] >= u[:, None])
        owner = np.where(hit.any(axis=1), hit.argmax(axis=1), self.world - 1)
        mine = np.nonzero(owner == self.rank)[0]
        phys = np.zeros(u.size, dtype=np.uint64)
        if mine.size:
            local = self.engine.sample_local(u[mine], strict, float(prefix[self.rank]))
            phys[mine] = local + (np.uint64(self.rank) << np.uint64(self.n_local))
        parts = [None] * self.world
        self.dist.all_gather_object(parts, (mine, phys[mine]))
        for idx, vals in parts:
            phys[idx] = vals
        return phys

    def gather_physical(self) -> np.ndarray:
        """Whole state in PHYSICAL index order (shards back to back) on every rank (tests / small states only)."""
        self._barrier()
        parts = [None] * self.world
        self.dist.all_gather_object(parts, self.engine.download())
        return np.concatenate(parts)

    def gather_logical(self) -> np.ndarray:
        """Whole state in logical index order on every rank (tests / small states only)."""
        self._barrier()
        parts = [None] * self.world
        self.dist.all_gather_object(parts, self.engine.download())
        phys = np.concatenate(parts)
        l2p = self.layout()
        idx = np.arange(phys.size)
        src = np.zeros_like(idx)
        for q in range(self.n):
            src |= ((idx >> q) & 1) << int(l2p[q])
        return phys[src]

    def save_wavefunction(self, filename: str):
        """SAVE-WAVEFUNCTION of the distributed QVM (dqvm/src/distributed-qvm.lisp:118-137): the file is the ordered
        amplitudes as consecutive (re, im) native doubles, every rank writing its own entries at byte offset
        16 * logical address (dqvm: MPI_File_write_at per amplitude; here one strided write per rank through a memory map).
        The logical address of physical index p comes from the current qubit layout."""
        self._barrier()
        local = self.engine.download()
        if self.rank == 0:
            with open(filename, "wb") as f:
                f.truncate(16 << self.n)
        self.dist.barrier()
        l2p = self.layout()
        phys = (np.arange(local.size, dtype=np.uint64) | (np.uint64(self.rank) << np.uint64(self.n_local)))
        logical = np.zeros_like(phys)
        for q in range(self.n):
            logical |= ((phys >> np.uint64(int(l2p[q]))) & np.uint64(1)) << np.uint64(q)
        mm = np.memmap(filename, dtype=np.complex128, mode="r+", shape=(1 << self.n,))
        mm[logical.astype(np.int64)] = local
        mm.flush()
        del mm
        self.dist.barrier()

    def load_wavefunction(self, filename: str):
        """LOAD-WAVEFUNCTION (dqvm/src/distributed-qvm.lisp:141-150): the ordered amplitudes back into the shards (the layout
        is reset to the identity first)."""
        self.set_zero_state()
        mm = np.memmap(filename, dtype=np.complex128, mode="r", shape=(1 << self.n,))
        lo = self.rank << self.n_local
        self.engine.upload(np.ascontiguousarray(mm[lo: lo + (1 << self.n_local)]))
        self._zero_ranks = 0
        del mm
        self._barrier()

    def scatter_logical(self, psi: np.ndarray):
        """Load a full logical state (identity layout required: call right after set_zero_state)."""
        assert (self.layout() == np.arange(self.n)).all()
        self._barrier()
        lo = self.rank << self.n_local
        self.engine.upload(np.ascontiguousarray(psi[lo: lo + (1 << self.n_local)]))
        self._zero_ranks = 0
        self._barrier()

    def close(self):
        self._barrier()
        self.engine.close()


# --------------------------------------------------------------------------------------- bench (N > 1)
class LocalShardGroup:
    """All shards of one state in ONE host process, one device each (`qvmcuda_shard_attach_local`): the shape of a single
    Lisp image driving several GPUs.  Same schedules, kernels and exchange passes as ShardedState; the barriers around peer
    steps are stream synchronisations of every shard instead of collective barriers.  Also what lets ncu profile an exchange
    pass (it cannot follow IPC-mapped memory across processes)."""

    def __init__(self, n_qubits: int, devices: Sequence[int], want_alt: bool = True):
        from . import _lib as L
        from .qvm import DeviceVector
        self.L = L
        self.world = len(devices)
        self.g = _log2(self.world)
        self.n = n_qubits
        self.n_local = n_qubits - self.g
        self.vecs = [DeviceVector(1 << self.n_local, d) for d in devices]
        arr = (C.c_void_p * self.world)(*[v.handle for v in self.vecs])
        L.check(L.lib().qvmcuda_shard_attach_local(arr, self.world, int(want_alt)))
        self.steps = self.peer_steps = 0
        self._zero_ranks = 0
        self.set_zero_state()

    def _sync(self):
        for v in self.vecs:
            v.synchronize()

    def set_zero_state(self):
        self._sync()
        self.vecs[0].set_basis_state(0)
        for v in self.vecs[1:]:
            self.L.check(self.L.lib().qvmcuda_shard_clear(v.handle))
        self._zero_ranks = ((1 << self.world) - 1) & ~1
        self._sync()

    def layout(self) -> np.ndarray:
        out = np.zeros(self.n, dtype=np.int32)
        self.L.check(self.L.lib().qvmcuda_state_layout(self.vecs[0].handle, self.L.ptr(out), self.n))
        return out

    def apply_gates(self, gates, fuse: bool = True, absorb_swaps: bool = True):
        L = self.L
        ks, qf, mf = L.flatten_gates(gates)
        flags = (L.FUSE if fuse else 0) | (L.ABSORB_SWAPS if absorb_swaps else 0)
        tapes = []
        for v in self.vecs:
            L.check(L.lib().qvmcuda_shard_set_zero_ranks(v.handle, self._zero_ranks))
        for v in self.vecs:
            h = C.c_void_p()
            L.check(L.lib().qvmcuda_shard_compile(v.handle, len(gates), L.ptr(ks), L.ptr(qf), L.ptr(mf), flags, C.byref(h)))
            tapes.append(h)
        try:
            n = C.c_int()
            L.check(L.lib().qvmcuda_tape_num_steps(tapes[0], C.byref(n)))
            for i in range(n.value):
                f = C.c_uint32()
                L.check(L.lib().qvmcuda_tape_step_flags(tapes[0], i, C.byref(f)))
                peer = f.value & STEP_PEER
                if peer:
                    self._sync()          # every shard complete before anyone reads it remotely
                for v, t in zip(self.vecs, tapes):
                    L.check(L.lib().qvmcuda_tape_run_step(v.handle, t, i))
                if peer:
                    self._sync()          # remote reads / writes landed before local work resumes
                    self.peer_steps += 1
                    self._zero_ranks = 0
                self.steps += 1
            for v, t in zip(self.vecs, tapes):
                L.check(L.lib().qvmcuda_tape_commit(v.handle, t))
        finally:
            for t in tapes:
                L.lib().qvmcuda_tape_destroy(t)

    def norm2(self) -> float:
        return float(sum(v.norm2() for v in self.vecs))

    def prob_excited(self, q: int) -> float:
        return float(sum(v.prob_excited(q) for v in self.vecs))

    def scatter_logical(self, psi: np.ndarray):
        """Canonical layout: shard r holds amplitudes r * 2^n_local ... (tests / small states)."""
        self.set_zero_state()
        for r, v in enumerate(self.vecs):
            v.upload(np.ascontiguousarray(psi[r << self.n_local:(r + 1) << self.n_local]))
        self._zero_ranks = 0

    def gather_logical(self) -> np.ndarray:
        self._sync()
        phys = np.concatenate([v.download() for v in self.vecs])
        l2p = self.layout()
        idx = np.arange(phys.size)
        src = np.zeros_like(idx)
        for q in range(self.n):
            src |= ((idx >> q) & 1) << int(l2p[q])
        return phys[src]

    def close(self):
        for v in self.vecs:
            v.close()


def _main_attr(name):
    """bench.py runs as __main__ under torchrun: borrow its helpers (ClockSampler, measured_peaks) when present."""
    import sys
    return getattr(sys.modules.get("__main__"), name, None)


def _start_clock_sampler(rank: int, local_rank: int):
    try:
        cls = _main_attr("ClockSampler")
        return cls(local_rank) if (rank == 0 and cls is not None) else None
    except Exception:
        return None


def _stop_clock_sampler(sampler):
    try:
        return sampler.stop() if sampler is not None else None
    except Exception:
        return None


def _sharded_roofline(local_qubits: int, step_seconds: float, passes_per_step: float, peer_passes_per_step: float,
                      peer_seconds_per_step: float, peer_bytes_per_step: float):
    """Per-GPU figures for the sharded run.  HBM: every pass reads and writes the rank's whole shard (32 B per amplitude,
    SURVEY 8d), so achieved = 32 * 2^local_qubits * passes / step time.  NVLink: the exchange passes pull
    (1 - 2^-pairs) of the shard from peer memory; their ingress rate is set against the measured peer-copy rate (770 GB/s per direction, B200_PROFILING.md) and the nominal 900 GB/s."""
    try:
        peaks = _main_attr("measured_peaks")
        peak, src = peaks() if peaks is not None else (6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)")
        traffic_fn = _main_attr("measured_traffic")
        traffic = traffic_fn() if traffic_fn is not None else {}
        bytes_per_pass = 32.0 * float(1 << local_qubits)
        achieved = bytes_per_pass * passes_per_step / step_seconds / 1e9 if step_seconds > 0 else None
        ratio = traffic.get("fused_pass_dram_bytes", 0) / traffic.get("algorithmic_bytes", 1) if traffic else 0
        nv = None
        if peer_seconds_per_step > 0:
            gbs = peer_bytes_per_step / peer_seconds_per_step / 1e9
            nv = {"bound": "nvlink", "achieved": gbs, "peak": 770.0, "unit": "GB/s", "frac": gbs / 770.0,
                  "nominal_peak": 900.0, "frac_of_nominal": gbs / 900.0,
                  "bytes_pulled_per_gpu_per_step": peer_bytes_per_step, "exchange_ms_per_step": 1e3 * peer_seconds_per_step,
                  "exchange_passes_per_step": peer_passes_per_step,
                  "peak_source": "measured peer copy, 770 GB/s per direction per GPU (B200_PROFILING.md); nominal NVLink 5: 900",
                  "what": "ingress of the passes that load through a qubit remap from peer shards while applying their gates"}
        local = None
        n_local_passes = passes_per_step - peer_passes_per_step
        if n_local_passes > 0 and step_seconds > peer_seconds_per_step:
            la = bytes_per_pass * n_local_passes / (step_seconds - peer_seconds_per_step) / 1e9
            local = {"achieved": la, "frac": la / peak, "passes_per_step": n_local_passes,
                     "ms_per_pass": 1e3 * (step_seconds - peer_seconds_per_step) / n_local_passes,
                     "what": "the passes that touch only the rank's own shard (HBM-bound): step time minus the exchange passes"}
        return {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s",
                "frac": (achieved / peak) if achieved else None, "local_passes": local,
                "traffic": (bytes_per_pass * ratio) if ratio else None,
                "traffic_source": (traffic.get("source", "") + "; DRAM bytes per pass scale with the shard") if ratio else None,
                "kernel": "compiled gate passes (qvj_kernel) + qv_tile_kernel per GPU, average over the passes of a step "
                          "(exchange passes are NVLink-ingress-bound, see nvlink)",
                "peak_source": src, "bytes_per_launch": bytes_per_pass, "launches_per_step": passes_per_step,
                "nvlink": nv}
    except Exception:
        return None


def sharded_parity_check(dist_mod, rank: int, world: int, device: int, n_small: int = 20) -> dict:
    """A small instance of the SAME sharded code path on the SAME ranks, checked against the CPU oracle before anything
    is timed: amplitudes within the north-star tolerance, sampled indices bit-exact, in both exchange modes (fused pull
    remaps into an alternate buffer / in-place peer passes)."""
    from . import circuits
    out = {"qubits": n_small, "tolerance": "1e-12 rel / 1e-14 abs, sampler indices identical", "modes": {}}
    rng = np.random.default_rng(12345)
    psi = rng.standard_normal(1 << n_small) + 1j * rng.standard_normal(1 << n_small)
    psi = np.ascontiguousarray(psi / np.linalg.norm(psi))
    circ = circuits.qft_circuit(range(n_small)) + circuits.random_layer_circuit(n_small, 2, 3)
    ref = None
    u = np.random.default_rng(2024).random(2000)
    ok_all = True
    for mode in ("pull", "inplace"):
        if mode == "inplace":
            os.environ["QVM_REMAP_INPLACE"] = "1"
        try:
            st = ShardedState(n_small, dist_mod, device=device)
            st.scatter_logical(psi)
            st.apply_gates(circ, fuse=True)
            got = st.gather_logical()
            phys = st.gather_physical()
            idx = st.sample_physical(u, False)
            res = {"exchange_passes": st.peer_steps}
            # ... and straight from the collective reset (the known-zero-shards path: lazily cleared shards, the first exchange
            # pass does not fetch them)
            st.set_zero_state()
            st.apply_gates(circ, fuse=True)
            got0 = st.gather_logical()
            if rank == 0:
                from oracle import oracle as O      # the checker, not the thing measured
                if ref is None:
                    ref = psi.copy()
                    for m, q in circ:
                        O.apply_matrix(ref, m, q)
                err = np.abs(got - ref)
                res["max_abs_err"] = float(err.max())
                res["amplitudes_ok"] = bool((err <= 1e-14 + 1e-12 * np.abs(ref)).all())
                res["sampler_bit_exact"] = bool((idx == O.sample_tree_sharded(phys, world, u, False)).all())
                ref0 = np.zeros(1 << n_small, dtype=np.complex128)
                ref0[0] = 1.0
                for m, q in circ:
                    O.apply_matrix(ref0, m, q)
                res["amplitudes_ok_from_reset"] = bool((np.abs(got0 - ref0) <= 1e-14 + 1e-12 * np.abs(ref0)).all())
                ok_all = ok_all and res["amplitudes_ok"] and res["sampler_bit_exact"] and res["amplitudes_ok_from_reset"]
            st.close()
            out["modes"][mode] = res
        finally:
            os.environ.pop("QVM_REMAP_INPLACE", None)
    out["ok"] = bool(ok_all)
    return out


def bench_sharded(args, rank: int, world: int, local_rank: int):
    """Weak scaling: every rank keeps a 2^L shard (L = args.local_qubits, default 32 = 64 GiB: 8 GPUs hold the 35-qubit
    state BASELINE.json names); the circuit is the QFT on L + log2(world) qubits.
    value = 30-qubit-equivalent gates/s = gates * 2^(n - 30) / s (amplitude updates per second / 2^30), so the series
    continues the N = 1 line (plain gates/s at 30 qubits)."""
    import torch
    import torch.distributed as dist

    from . import _lib, circuits

    # the ranks of one box share its host cores: split them between the ranks' pass-compiler pools
    os.environ.setdefault("QVMCUDA_JIT_THREADS", str(max(1, (os.cpu_count() or 8) // world)))
    torch.cuda.set_device(local_rank)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    g = _log2(world)
    L = args.local_qubits
    n = L + g
    parity = None
    if not args.no_parity_check:
        parity = sharded_parity_check(dist, rank, world, local_rank)
    st = ShardedState(n, dist, device=local_rank)
    stream = torch.cuda.Stream()
    st.engine.vec.set_stream(stream.cuda_stream)
    gates = circuits.qft_circuit(range(n))
    wl = f"qft-{n}"

    def step():
        st.apply_gates(gates, fuse=True)

    def timed(fn, steps):
        """CUDA events on the stream the kernels run on, bracketed by barrier + synchronize; max over ranks."""
        st._barrier()
        torch.cuda.synchronize()
        ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        ev0.record(stream)
        for _ in range(steps):
            fn()
        ev1.record(stream)
        st._barrier()
        torch.cuda.synchronize()
        t = torch.tensor([ev0.elapsed_time(ev1) / 1e3], device="cuda", dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    def settle(fn, max_iters):
        """Untimed preparation.  A sharded state is never moved back to the canonical qubit layout (the host maps indices through
        the layout instead), so consecutive runs of the same circuit start from different layouts and the schedule -- and with
        it the set of compiled passes -- only repeats after a few runs.  Run until one whole circuit needs no new kernel."""
        iters = 0
        for _ in range(max_iters):
            before = _lib.jit_stats()
            fn()
            st._barrier()
            after = _lib.jit_stats()
            new = torch.tensor([after["compiled"] + after["disk_hits"] - before["compiled"] - before["disk_hits"]], device="cuda",
                               dtype=torch.int64)
            dist.all_reduce(new, op=dist.ReduceOp.MAX)
            iters += 1
            if int(new.item()) == 0:
                break
        return iters

    j0 = _lib.jit_stats()
    t_prep = time.perf_counter()
    prep_iters = settle(step, 16)           # untimed: schedules, compiles the passes (pass compiler), first touch
    prep_s = time.perf_counter() - t_prep
    for _ in range(args.warmup):
        step()
    l0 = _lib.launch_count()
    p0, s0, ps0, pb0 = st.peer_steps, st.steps, st.peer_seconds, st.peer_bytes
    sampler = _start_clock_sampler(rank, local_rank)     # nvidia-smi clocks during the timed region (rank 0)
    jt0 = _lib.jit_stats()
    dt = timed(step, args.steps)
    jt1 = _lib.jit_stats()
    launches = _lib.launch_count() - l0
    clocks = _stop_clock_sampler(sampler)
    passes_per_step = (st.steps - s0) / args.steps
    peer_per_step = (st.peer_steps - p0) / args.steps
    peer_s = (st.peer_seconds - ps0) / args.steps
    peer_b = (st.peer_bytes - pb0) / args.steps
    norm2 = st.norm2()

    # ---- e2e: through the public API with host buffers every step: reset, gate arrays -> schedule -> device, run,
    #      1000-shot sample and one probability back on the host
    u = np.random.default_rng(2024).random(1000)

    def e2e_step():
        st.set_zero_state()
        st.apply_gates(gates, fuse=True)
        st.sample(u, strict=False)
        st.prob_excited(0)

    e2e_step()
    e2e_steps = max(1, min(args.steps, 3))
    st._barrier()
    t0 = time.perf_counter()
    for _ in range(e2e_steps):
        e2e_step()
    st._barrier()
    te = torch.tensor([time.perf_counter() - t0], device="cuda", dtype=torch.float64)
    dist.all_reduce(te, op=dist.ReduceOp.MAX)
    e2e_dt = float(te.item())
    ks, qf, mf = _lib.flatten_gates(gates)
    h2d = ks.nbytes + qf.nbytes + mf.nbytes + u.nbytes
    d2h = u.size * 8 + 8 * 2       # sampled indices, the shard total for the sampler prefix and one probability partial

    # ---- BASELINE configs[4]: random 1q (RZ.RY.RZ) / CZ layers on all n qubits, sharded
    c5 = None
    if args.c5_layers > 0:
        rgates = circuits.random_layer_circuit(n, args.c5_layers, seed=0)

        def rstep():
            st.apply_gates(rgates, fuse=True)

        # Every run starts from |0...0> in the canonical layout (set_zero_state resets it), so the schedule -- and the set of
        # compiled passes -- is the same for every run: one untimed run compiles them, the timed one finds them in the cache.
        def rrun():
            st.set_zero_state()
            rstep()

        c5_prep = settle(rrun, 3)           # untimed
        st.set_zero_state()
        p1, s1, ps1, pb1 = st.peer_steps, st.steps, st.peer_seconds, st.peer_bytes
        j_before = _lib.jit_stats()
        rdt = timed(rstep, 1)
        j_after = _lib.jit_stats()
        rp = st.steps - s1
        c5 = {"workload": f"random 1q(RZ.RY.RZ)/CZ circuit, {args.c5_layers} layers on {n} qubits, seed 0, from |0...0>", "gates": len(rgates),
              "seconds_per_circuit": rdt, "gates_per_s": len(rgates) / rdt, "value_30q_equivalent": len(rgates) * 2.0 ** (n - 30) / rdt,
              "hbm_passes": rp, "exchange_passes": st.peer_steps - p1,
              "roofline": _sharded_roofline(L, rdt, rp, st.peer_steps - p1, st.peer_seconds - ps1, st.peer_bytes - pb1),
              "bound_note": "layers of general 1q gates are bound by the FP64 pipe on B200, not by HBM (see DESIGN.md); the HBM figure "
                            "is reported for comparison with the QFT line",
              "prepare_circuits": c5_prep, "kernels_compiled_in_timed_region": j_after["compiled"] - j_before["compiled"],
              "norm2": st.norm2()}
    j1 = _lib.jit_stats()
    if rank == 0:
        scale = 2.0 ** (n - 30)
        value = len(gates) * args.steps * scale / dt
        print(json.dumps({
            "metric": "gates/s", "value": value, "unit": "gates/s", "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": 1e3 * dt / args.steps, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": f"{wl} (examples/qft.lisp qft-circuit, {len(gates)} gates) sharded over {world} GPUs: {n} qubits, "
                                   f"{L} local qubits = {16 << L} B per GPU, gate fusion on",
                       "value_definition": "gates/s x 2^(n-30): amplitude updates per second / 2^30 (30-qubit-equivalent gates/s); "
                                           "equals plain gates/s on the 30-qubit N=1 line",
                       "raw_gates_per_s": len(gates) * args.steps / dt, "circuit_seconds": dt / args.steps,
                       "hbm_passes_per_step": passes_per_step, "exchange_passes_per_step": peer_per_step,
                       "exchange": "tile kernel P2P loads over NVLink (IPC-mapped shards), remap fused into the gate pass; "
                                   "torch.distributed (NCCL) for barriers and scalars only",
                       "swap_gates": "exact SWAP gates are relabelings of the sharded state's qubit layout (never undone: indices are "
                                     "mapped on the host); the N = 1 line executes them (canonical layout after every circuit)",
                       "l2_policy": "shard (64 GiB) is far larger than the 126 MB L2; no flush needed",
                       "timing": "CUDA events on the kernels' stream, bracketed by barrier + cudaDeviceSynchronize, max over ranks",
                       "pass_compiler": {"prepare_seconds": prep_s, "prepare_circuits": prep_iters,
                                         "kernels_compiled": j1["compiled"] - j0["compiled"],
                                         "kernels_compiled_in_timed_region": jt1["compiled"] - jt0["compiled"],
                                         "disk_cache_hits": j1["disk_hits"] - j0["disk_hits"], "compile_ms_total": j1["compile_ms"] - j0["compile_ms"]}},
            "roofline": _sharded_roofline(L, dt / args.steps, passes_per_step, peer_per_step, peer_s, peer_b),
            "parity_check": parity, "c5_random": c5,
            "clocks": clocks,
            "gpu_launches": int(launches), "norm2": norm2,
            "e2e": {"value": len(gates) * e2e_steps * scale / e2e_dt, "unit": "gates/s", "h2d_bytes_per_step": int(h2d),
                    "d2h_bytes_per_step": int(d2h), "steps": e2e_steps,
                    "what": "reset + apply_gates(host gate arrays: scheduled, uploaded and run) + 1000-shot sample + prob_excited, "
                            "host wall clock, max over ranks"},
        }))
    st.close()
    dist.destroy_process_group()
