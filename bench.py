#!/usr/bin/env python
"""bench.py -- headline benchmark of the QVM hot path on B200.

Metric (BASELINE.json): gates/s (and HBM GB/s) of gate application on a 30-qubit state, workload
`configs[1]`: the 30-qubit QFT of examples/qft.lisp (480 gates), gate fusion on; the unfused
per-gate-pass bandwidth is reported beside it.  A "step" = one run of the whole circuit over the
device-resident 16 GiB state.

    python bench.py --gpus N --steps K --warmup W            # our arm
    python bench.py --impl reference --gpus N --steps K ...   # the reference's CPU path (oracle port, all host threads)

Under torchrun (N > 1) every rank owns one 64 GiB shard (32 local qubits; the top log2(N) qubits select the rank), weak
scaling: 32 + log2(N) qubits in total, i.e. the 35-qubit state of BASELINE.json on 8 GPUs (qvm_b200/dist.py).

The CPU arm (`cpu_baseline`, `--impl reference`) times baseline/cpu_port.c -- an AVX2/FMA + OpenMP port of the reference's
1q/2q kernels and its contiguous-range work split -- on ALL host cores (OMP_NUM_THREADS is ignored: torchrun sets it to 1).
The reference itself (Common Lisp) needs SBCL, which the image lacks.  The CPU arm applies one gate per pass, as the
reference's transitions do; the GPU arm is reported both ways (fused passes = headline, one gate per pass = `unfused`), and
`ratios` spells out which pair every quotient compares.
"""
from __future__ import annotations

import argparse
import json
import math
import os
import subprocess
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

ALGO_BYTES_PER_AMP = 32   # one gate pass reads 16 B and writes 16 B per amplitude


def measured_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        try:
            return float(json.load(open(path))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


def measured_traffic():
    """dram__bytes_read.sum + dram__bytes_write.sum per launch from the committed ncu --set full capture."""
    path = os.path.join(ROOT, "profiles", "r02_traffic.json")
    try:
        return json.load(open(path))
    except Exception:
        return {}


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index: int):
        self.proc = None
        self.path = f"/tmp/qvm_bench_clocks_{os.getpid()}.csv"
        try:
            self.f = open(self.path, "w")
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", "100", "-i", str(gpu_index)], stdout=self.f, stderr=subprocess.DEVNULL)
        except Exception:
            self.proc = None

    def stop(self):
        out = {"sm_mhz": None, "sm_max_mhz": None, "reasons": []}
        if self.proc is None:
            return out
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        self.f.close()
        sm, mx, reasons = [], [], set()
        for line in open(self.path):
            p = [x.strip() for x in line.split(",")]
            if len(p) < 9:
                continue
            try:
                sm.append(float(p[1]))
                mx.append(float(p[2]))
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), p[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        if sm:
            out = {"sm_mhz": float(np.median(sm)), "sm_max_mhz": float(max(mx)), "reasons": sorted(reasons),
                   "samples": len(sm)}
        try:
            os.remove(self.path)
        except OSError:
            pass
        return out


def qft_gates(n):
    from qvm_b200 import circuits
    return circuits.qft_circuit(range(n))


# ------------------------------------------------------------------------------- CPU arm
def cpu_sample(n_qubits: int, budget_s: float, max_gates: int = 96):
    """Time the CPU port (baseline/cpu_port.c: AVX2/FMA kernels, all host cores, contiguous ranges like
    lparallel:pdotimes) on the first gates of the same circuit, one full pass per gate, bounded by budget_s."""
    from baseline import cpu_port as P
    threads = P.all_cores()
    gates = qft_gates(n_qubits)
    psi = P.zero_state(n_qubits, threads)                               # parallel first touch, untimed
    P.apply_gate(psi, gates[0][0], gates[0][1], threads)                # warm-up, untimed
    done, t0 = 0, time.perf_counter()
    for m, q in gates[1:1 + max_gates]:
        P.apply_gate(psi, m, q, threads)
        done += 1
        if time.perf_counter() - t0 > budget_s:
            break
    dt = time.perf_counter() - t0
    return {"value": done / dt, "unit": "gates/s", "cores": threads, "kind": "port",
            "hbm_equivalent_gbs": done * ALGO_BYTES_PER_AMP * (1 << n_qubits) / dt / 1e9,
            "sample": f"gates 2..{done + 1} of the {n_qubits}-qubit QFT, one pass per gate (the reference's per-gate transitions), "
                      f"{dt:.1f} s, baseline/cpu_port.c (AVX2/FMA port of src/impl/sbcl-avx-vops.lisp + OpenMP contiguous ranges), "
                      f"{threads} threads = all host cores; the reference itself needs SBCL (absent)"}


def run_reference(args):
    """The reference's CPU path on the host cores (rank 0 only; the other ranks exit at once)."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from baseline import cpu_port as P
    world = max(1, args.gpus)
    n_ours = args.qubits if world == 1 else args.local_qubits + int(math.log2(world))
    # The host cannot hold the 33..35-qubit states of the N > 1 line (128 GiB..512 GiB): the CPU runs the 30-qubit circuit and
    # reports the same size-normalised unit (amplitude updates per second / 2^30), which for a memory-bound CPU kernel does
    # not depend on the state size.
    n = min(n_ours, args.qubits)
    threads = P.all_cores()
    gates = qft_gates(n)
    psi = P.zero_state(n, threads)
    per_step = args.ref_gates_per_step
    pos = 0

    def step():
        nonlocal pos
        for _ in range(per_step):
            m, q = gates[pos % len(gates)]
            P.apply_gate(psi, m, q, threads)
            pos += 1

    for _ in range(args.warmup):
        step()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        step()
    dt = time.perf_counter() - t0
    val = per_step * args.steps / dt * 2.0 ** (n - 30)
    sample = (f"{per_step} consecutive gates of the {n}-qubit QFT per step, one pass per gate, {threads} threads (all host cores, "
              f"OMP_NUM_THREADS ignored), baseline/cpu_port.c (AVX2/FMA + OpenMP port; the reference needs SBCL)")
    if n != n_ours:
        sample += f"; stands for the {n_ours}-qubit workload of the GPU arm in the size-normalised unit (gates/s x 2^(n-30))"
    print(json.dumps({
        "impl": "reference", "metric": "gates/s", "value": val, "unit": "gates/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * dt / args.steps, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": f"qft-{n} (examples/qft.lisp qft-circuit), PURE-STATE-QVM, complex double, one pass per gate (no fusion)"},
        "cpu_baseline": {"value": val, "unit": "gates/s", "cores": threads, "kind": "port", "sample": sample},
        "e2e": {"value": val, "unit": "gates/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }))


# ------------------------------------------------------------------------------- GPU arm
def run_ours(args):
    import torch

    from qvm_b200 import _lib, qvm

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if world > 1:
        from qvm_b200 import dist
        return dist.bench_sharded(args, rank, world, local_rank)

    n = args.qubits
    torch.cuda.set_device(local_rank)
    stream = torch.cuda.Stream()
    vec = qvm.DeviceVector(1 << n, device=local_rank)
    vec.set_stream(stream.cuda_stream)
    gates = qft_gates(n)
    tape = qvm.Tape(n, gates, fuse=True)
    info = tape.info()
    peak, peak_src = measured_peaks()
    bytes_per_pass = ALGO_BYTES_PER_AMP * (1 << n)

    def timed(fn, steps, warmup):
        for _ in range(warmup):
            fn()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        l0 = _lib.launch_count()
        with torch.cuda.stream(stream):
            e0.record(stream)
            for _ in range(steps):
                fn()
            e1.record(stream)
        torch.cuda.synchronize()
        return e0.elapsed_time(e1) / 1e3, _lib.launch_count() - l0

    # ---- untimed: the pass compiler turns the fused passes into kernels (in-tree disk cache, else NVRTC)
    j0 = _lib.jit_stats()
    t_prep = time.perf_counter()
    vec.set_zero_state()
    vec.run_tape(tape)
    vec.synchronize()
    prep_s = time.perf_counter() - t_prep
    j1 = _lib.jit_stats()

    # ---- headline: fused tape, state resident in HBM (the state the preparation run left behind: a reset here would be lazy and
    #      make the first pass of the first step a cheaper write-only pass)
    clocks = ClockSampler(local_rank)
    dt, launches = timed(lambda: vec.run_tape(tape), args.steps, args.warmup)
    clk = clocks.stop()
    value = len(gates) * args.steps / dt
    pass_s = dt / (args.steps * info["passes"])
    traffic = measured_traffic()
    scale = (1 << n) / float(1 << 30)          # the ncu capture is at 30 qubits; DRAM bytes scale with the state
    roofline = {"bound": "hbm", "achieved": bytes_per_pass / pass_s / 1e9, "peak": peak, "unit": "GB/s",
                "frac": bytes_per_pass / pass_s / 1e9 / peak,
                "traffic": traffic.get("fused_pass_dram_bytes", 0) * scale or None,
                "kernel": "fused passes of the timed region: qvj_kernel (compiled passes, 27-212 gates per launch) and qv_tile_kernel "
                          "(the two SWAP-only passes); see profiles/r02_*.md",
                "peak_source": peak_src, "traffic_source": traffic.get("source"),
                "bytes_per_launch": bytes_per_pass, "launches_per_step": info["passes"]}

    # ---- unfused: every gate its own HBM pass (bounded sample of the same circuit)
    sample = gates[: args.unfused_gates]
    tape_u = qvm.Tape(n, sample, fuse=False)
    dtu, _ = timed(lambda: vec.run_tape(tape_u), 1, 1)
    pass_u = dtu / tape_u.info()["passes"]
    unfused = {"gates_per_s": len(sample) / dtu, "hbm_gbs": bytes_per_pass / pass_u / 1e9,
               "frac_of_peak": bytes_per_pass / pass_u / 1e9 / peak, "ms_per_gate_pass": 1e3 * pass_u,
               "sample": f"first {len(sample)} gates of the circuit, one tile-kernel launch per gate"}
    # the same kernel with ONE 1q/2q gate per launch: the HBM-bound case BASELINE.json's 70 % target is about
    roofline_unfused = {"bound": "hbm", "achieved": unfused["hbm_gbs"], "peak": peak, "unit": "GB/s",
                        "frac": unfused["frac_of_peak"],
                        "traffic": traffic.get("single_gate_pass_dram_bytes", 0) * scale or None,
                        "kernel": "qv_tile_kernel, one gate per launch (gate fusion off)", "peak_source": peak_src,
                        "bytes_per_launch": bytes_per_pass, "launches": tape_u.info()["passes"]}

    # ---- e2e: through the public API with host buffers: reset, program (host arrays) -> device, run,
    #      10^3-shot sample + one probability back to the host.
    u = np.random.default_rng(2024).random(1000)
    ks, qf, mf = _lib.flatten_gates(gates)
    h2d = ks.nbytes + qf.nbytes + mf.nbytes + u.nbytes + info["table_bytes"]
    d2h = u.size * 8 + 8

    def e2e_step():
        vec.set_zero_state()
        vec.apply_gates(gates, fuse=True)
        vec.sample(u, strict=False)
        vec.prob_excited(0)

    for _ in range(2):
        e2e_step()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        e2e_step()
    torch.cuda.synchronize()
    e2e_dt = time.perf_counter() - t0
    e2e = {"value": len(gates) * args.steps / e2e_dt, "unit": "gates/s", "h2d_bytes_per_step": int(h2d),
           "d2h_bytes_per_step": int(d2h),
           "what": "qvm reset + apply_gates(host gate arrays, scheduled and uploaded inside) + 1000-shot sample + prob readback"}

    vec.close()
    cpu = None
    ratios = None
    if not args.no_cpu_baseline:
        cpu = cpu_sample(n, args.cpu_budget)
        ratios = {"gpu_one_gate_per_pass_vs_cpu_one_gate_per_pass": unfused["gates_per_s"] / cpu["value"],
                  "gpu_fused_vs_cpu_one_gate_per_pass": value / cpu["value"],
                  "note": "like for like is the first (both sides one HBM/DRAM pass per gate); the second also contains the "
                          "scheduler's packing of 480 gates into a handful of passes (config.hbm_passes_per_step), which a same-support matrix fusion on the CPU side "
                          "would not give the QFT (no two consecutive gates act on the same qubits)"}

    print(json.dumps({
        "metric": "gates/s", "value": value, "unit": "gates/s", "n_gpus": 1, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": 1e3 * dt / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f64", "data": "synthetic",
        "config": {"workload": f"qft-{n} (examples/qft.lisp qft-circuit, {len(gates)} gates), PURE-STATE-QVM, complex double, gate fusion on",
                   "state_bytes": 16 << n, "l2_policy": "state (16 GiB at 30 qubits) is far larger than the 126 MB L2; no flush needed",
                   "hbm_passes_per_step": info["passes"],
                   "pass_compiler": {"policy": j1["policy"], "prepare_seconds": prep_s, "kernels_compiled": j1["compiled"] - j0["compiled"],
                                     "disk_cache_hits": j1["disk_hits"] - j0["disk_hits"], "compile_ms_total": j1["compile_ms"] - j0["compile_ms"]}},
        "roofline": roofline, "roofline_unfused": roofline_unfused, "unfused": unfused, "cpu_baseline": cpu, "ratios": ratios, "e2e": e2e, "gpu_launches": int(launches),
        "clocks": clk,
    }))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--qubits", type=int, default=30)
    ap.add_argument("--unfused-gates", type=int, default=60)
    ap.add_argument("--cpu-budget", type=float, default=15.0)
    ap.add_argument("--ref-gates-per-step", type=int, default=4)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--local-qubits", type=int, default=32, help="N>1: qubits per shard (32 = 64 GiB per GPU; 8 GPUs = 35 qubits)")
    ap.add_argument("--c5-layers", type=int, default=10, help="N>1: layers of the random 1q/CZ circuit timed beside the QFT (0 = skip)")
    ap.add_argument("--no-parity-check", action="store_true", help="N>1: skip the small sharded parity check before the timed region")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
