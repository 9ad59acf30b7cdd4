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

    def run_step(self, tape, i): self.L.check(self.L.lib().qvmcuda_tape_run_step(self.vec.handle, tape, i))
    def commit(self, tape): self.L.check(self.L.lib().qvmcuda_tape_commit(self.vec.handle, tape))
    def free_tape(self, tape): self.L.lib().qvmcuda_tape_destroy(tape)
    def synchronize(self): self.vec.synchronize()

    def layout(self) -> np.ndarray:
        out = np.zeros(self.n_total, dtype=np.int32)
        self.L.check(self.L.lib().qvmcuda_state_layout(self.vec.handle, self.L.ptr(out), self.n_total))
        return out

    def set_basis_local(self, index): self.vec.set_basis_state(index)
    def clear(self):
        self.vec.set_basis_state(0)
        self.vec.scale(0.0)
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
        # content replaced: layout resets inside the engine (set_basis_state); barrier so that no peer pass
        # starts before every shard is initialised
        self._barrier()

    def apply_gates(self, gates, fuse: bool = True, absorb_swaps: bool = False):
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
                    t0 = time.perf_counter()
                self.engine.run_step(tape, i)
                if trace:
                    self.engine.synchronize()
                    print(f"[dist] step {i}: {'REMAP' if flags & STEP_REMAP else 'PEER' if peer else 'LOCAL'} "
                          f"{1e3 * (time.perf_counter() - t0):.2f} ms", flush=True)
                if peer:
                    self._barrier()       # remote writes must have landed before local work resumes
                    self.peer_steps += 1
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

    def scatter_logical(self, psi: np.ndarray):
        """Load a full logical state (identity layout required: call right after set_zero_state)."""
        assert (self.layout() == np.arange(self.n)).all()
        self._barrier()
        lo = self.rank << self.n_local
        self.engine.upload(np.ascontiguousarray(psi[lo: lo + (1 << self.n_local)]))
        self._barrier()

    def close(self):
        self._barrier()
        self.engine.close()


# --------------------------------------------------------------------------------------- bench (N > 1)
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


def _sharded_roofline(local_qubits: int, step_seconds: float, passes_per_step: float, peer_passes_per_step: float):
    """Per-GPU figure for the sharded run: every pass reads and writes the rank's whole shard (32 B per amplitude,
    SURVEY 8d), so achieved = 32 * 2^local_qubits * passes / step time.  Pull passes are bounded by NVLink ingress,
    not by HBM: their count is reported beside it."""
    try:
        peaks = _main_attr("measured_peaks")
        peak, src = peaks() if peaks is not None else (6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)")
        bytes_per_pass = 32.0 * float(1 << local_qubits)
        achieved = bytes_per_pass * passes_per_step / step_seconds / 1e9 if step_seconds > 0 else None
        return {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s",
                "frac": (achieved / peak) if achieved else None, "traffic": None,
                "kernel": "qv_tile_kernel per GPU, average over the passes of a step (pull passes are NVLink-ingress-bound)",
                "peak_source": src, "bytes_per_launch": bytes_per_pass, "launches_per_step": passes_per_step,
                "pull_passes_per_step": peer_passes_per_step}
    except Exception:
        return None



def bench_sharded(args, rank: int, world: int, local_rank: int):
    """Weak scaling: every rank keeps a 2^args.qubits shard; the circuit is the QFT on
    args.qubits + log2(world) qubits.  value = 30-qubit-equivalent gates/s = gates * 2^(n - args.qubits) / s,
    i.e. amplitude updates per second divided by 2^args.qubits, so N = 1 is the plain gates/s."""
    import torch
    import torch.distributed as dist

    from . import _lib, circuits

    torch.cuda.set_device(local_rank)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    g = _log2(world)
    n = args.qubits + g
    st = ShardedState(n, dist, device=local_rank)
    if getattr(args, "workload", "qft") == "random":
        gates = circuits.random_layer_circuit(n, args.layers, seed=0)
        wl = f"random 1q(RZ.RY.RZ)/CZ circuit, {args.layers} layers, {n} qubits (SURVEY 8d C5)"
    else:
        gates = circuits.qft_circuit(range(n))
        wl = f"qft-{n}"

    def step():
        st.apply_gates(gates, fuse=True, absorb_swaps=False)

    for _ in range(args.warmup):
        step()
    st._barrier()
    torch.cuda.synchronize()
    l0 = _lib.launch_count()
    p0, s0 = st.peer_steps, st.steps
    sampler = _start_clock_sampler(rank, local_rank)     # nvidia-smi clocks during the timed region (rank 0)
    t0 = time.perf_counter()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ev0.record()
    for _ in range(args.steps):
        step()
    st._barrier()
    ev1.record()
    torch.cuda.synchronize()
    dt_local = time.perf_counter() - t0
    t = torch.tensor([dt_local], device="cuda", dtype=torch.float64)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    dt = float(t.item())
    launches = _lib.launch_count() - l0
    clocks = _stop_clock_sampler(sampler)
    norm2 = st.norm2()
    if rank == 0:
        scale = 2.0 ** (n - args.qubits)
        value = len(gates) * args.steps * scale / dt
        peer_per_step = (st.peer_steps - p0) / args.steps
        print(json.dumps({
            "metric": "gates/s", "value": value, "unit": "gates/s", "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": 1e3 * dt / args.steps, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": f"{wl} sharded over {world} GPUs ({args.qubits} local qubits = {16 << args.qubits} B per GPU), gate fusion on",
                       "value_definition": f"gates/s x 2^(n-{args.qubits}): amplitude updates per second / 2^{args.qubits}; equals plain gates/s at N=1",
                       "raw_gates_per_s": len(gates) * args.steps / dt,
                       "hbm_passes_per_step": (st.steps - s0) / args.steps, "peer_passes_per_step": peer_per_step,
                       "exchange": "tile kernel P2P loads/stores over NVLink (IPC-mapped shards); torch.distributed barriers only",
                       "timing": "host wall clock bracketed by barrier + cudaDeviceSynchronize, max over ranks"},
            "roofline": _sharded_roofline(args.qubits, dt / args.steps, (st.steps - s0) / args.steps, peer_per_step),
            "clocks": clocks,
            "gpu_launches": int(launches), "norm2": norm2,
            "e2e": {"value": value, "unit": "gates/s", "h2d_bytes_per_step": int(sum(np.asarray(m).nbytes for m, _ in gates)),
                    "d2h_bytes_per_step": 0, "what": "apply_gates from host gate arrays (scheduled, uploaded and run inside the timed region)"},
        }))
    st.close()
    dist.destroy_process_group()
