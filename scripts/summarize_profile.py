"""Turn an ncu launch list (csv) + a `--set full` report into a small tracked summary under profiles/.

usage: python scripts/summarize_profile.py <launches.csv> <report.ncu-rep> <out.md> [title]
"""
import csv
import subprocess
import sys

launches, rep, out = sys.argv[1:4]
title = sys.argv[4] if len(sys.argv) > 4 else out

METRICS = [
    ("gpu__time_duration.sum", "duration"),
    ("dram__bytes_read.sum", "dram read"),
    ("dram__bytes_write.sum", "dram write"),
    ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram % of ncu peak"),
    ("launch__registers_per_thread", "regs/thread"),
    ("launch__occupancy_limit_registers", "occ limit regs (CTAs)"),
    ("launch__occupancy_limit_shared_mem", "occ limit smem (CTAs)"),
    ("sm__warps_active.avg.pct_of_peak_sustained_active", "achieved occupancy %"),
    ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue active %"),
    ("sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active", "fp64 pipe active %"),
    ("smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio", "stall long_scoreboard"),
    ("smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio", "stall short_scoreboard"),
    ("smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio", "stall barrier"),
    ("smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio", "stall mio_throttle"),
    ("smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio", "stall lg_throttle"),
    ("smsp__average_warps_issue_stalled_wait_per_issue_active.ratio", "stall wait (fixed latency)"),
    ("smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio", "stall no_instruction"),
    ("smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio", "stall math_pipe_throttle"),
    ("smsp__inst_executed.sum", "warp instructions"),
    ("sm__inst_executed.avg.per_cycle_active", "IPC per SM"),
    ("l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "smem bank conflicts"),
    ("l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "smem wavefronts"),
]

lines = [f"# {title}", ""]
rows = [r for r in csv.reader(open(launches)) if len(r) > 5]
hdr = rows[0]
ik, iv, im, ii = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Name"), hdr.index("ID")
byid = {}
for r in rows[1:]:
    d = byid.setdefault(r[ii], {"name": r[ik].split("(")[0].replace("void ", "")[:60]})
    d[r[im]] = float(r[iv].replace(",", ""))
lines += ["## Launch list (`ncu --metrics gpu__time_duration.sum,smsp__inst_executed.sum --clock-control none`; cold-cache, "
          "serialised: compare shares)", "", "| # | kernel | duration | warp instructions |", "|---|---|---|---|"]
tot = {}
for i, d in byid.items():
    t = d.get("gpu__time_duration.sum", 0.0)
    n = d.get("smsp__inst_executed.sum")
    lines.append(f"| {i} | `{d['name']}` | {t / 1e6:.3f} ms | {'' if n is None else f'{n / 1e9:.3f} G'} |")
    tot[d["name"]] = tot.get(d["name"], 0.0) + t
all_t = sum(tot.values())
lines += ["", "| kernel | total | share |", "|---|---|---|"]
for k, v in sorted(tot.items(), key=lambda kv: -kv[1]):
    lines.append(f"| `{k}` | {v / 1e6:.2f} ms | {100 * v / all_t:.1f} % |")

raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rr = list(csv.reader(raw.splitlines()))
h, units = rr[0], rr[1]
idx = {n: i for i, n in enumerate(h)}
lines += ["", "## `ncu --set full` capture (per launch, same command)", ""]
names = [r[idx["Kernel Name"]].split("(")[0].replace("void ", "")[:40] for r in rr[2:]]
lines.append("| metric | " + " | ".join(f"#{i}" for i in range(len(names))) + " | unit |")
lines.append("|---|" + "---|" * (len(names) + 1))
for m, label in METRICS:
    if m not in idx:
        continue
    vals = []
    for r in rr[2:]:
        v = r[idx[m]]
        try:
            vals.append(f"{float(v):.4g}")
        except ValueError:
            vals.append(v)
    lines.append(f"| {label} | " + " | ".join(vals) + f" | {units[idx[m]]} |")
lines += ["", "kernels: " + ", ".join(f"#{i} `{n}`" for i, n in enumerate(names)), ""]
open(out, "w").write("\n".join(lines))
print("\n".join(lines[-30:]))
