#!/usr/bin/env python
"""Summarise a gpurun round's ncu output into profiles/<tag>_ncu_summary.md.

    python scripts/ncu_summary.py <tag>      # reads gpurun_out/<tag>_launches.csv, gpurun_out/<tag>_prof.ncu-rep
"""
import collections
import csv
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
tag = sys.argv[1]
G = os.path.join(ROOT, "gpurun_out")
out = []

# ---- launch list (gpu__time_duration.sum per launch; cold-cache, serialised: compare shares)
lp = os.path.join(G, f"{tag}_launches.csv")
if os.path.exists(lp):
    rows = [r for r in csv.reader(open(lp)) if len(r) > 5]
    hdr = rows[0]
    i_name, i_val = hdr.index("Kernel Name"), hdr.index("Metric Value")
    agg = collections.OrderedDict()
    for r in rows[1:]:
        n = r[i_name].split("(")[0].replace("void ", "")
        agg.setdefault(n, []).append(float(r[i_val].replace(",", "")))
    ours = {n: v for n, v in agg.items() if n.startswith(("ups::", "tc::", "tma::", "pctc::", "standin::", "dp::"))}
    tot = sum(sum(v) for v in ours.values())
    out.append(f"## Launch list ({tag}_launches.csv: `ncu --metrics gpu__time_duration.sum --clock-control none` over "
               "`python bench.py --steps 2 --warmup 3 --no-cpu --no-e2e --no-n4 --no-scale-workloads`)\n")
    out.append("Library kernels only (torch's RNG/fill kernels that build the synthetic inputs are excluded from the share).\n")
    out.append("| kernel | launches | mean µs | share of path time |\n|---|---|---|---|")
    for n, v in ours.items():
        out.append(f"| `{n}` | {len(v)} | {sum(v) / len(v) / 1e3:.1f} | {sum(v) / tot:.3f} |")
    out.append("")

# ---- full capture
rp = os.path.join(G, f"{tag}_prof.ncu-rep")
if os.path.exists(rp):
    raw = subprocess.run(["ncu", "-i", rp, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    hdr, units = rows[0], rows[1]
    idx = {h: i for i, h in enumerate(hdr)}
    want = [("gpu__time_duration.sum", "duration"),
            ("dram__bytes_read.sum", "DRAM read"), ("dram__bytes_write.sum", "DRAM write"),
            ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "DRAM % of peak"),
            ("sm__throughput.avg.pct_of_peak_sustained_elapsed", "SM % of peak"),
            ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue active %"),
            ("sm__warps_active.avg.pct_of_peak_sustained_active", "warps active %"),
            ("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "tensor pipe active %"),
            ("launch__registers_per_thread", "regs/thread"),
            ("launch__grid_size", "grid"), ("launch__block_size", "block"),
            ("l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "smem wavefronts"),
            ("l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "smem bank conflicts"),
            ("smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio", "stall long_scoreboard"),
            ("smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio", "stall short_scoreboard"),
            ("smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio", "stall mio_throttle"),
            ("smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio", "stall barrier"),
            ("smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio", "stall math_pipe"),
            ("smsp__average_warps_issue_stalled_wait_per_issue_active.ratio", "stall wait")]
    out.append(f"## Full capture ({tag}_prof.ncu-rep: `ncu --set full --clock-control none --import-source on`, one launch "
               "per kernel after warm-up, CUB B=256 S=128 K=16 F=64)\n")
    names = [r[idx["Kernel Name"]].split("(")[0].replace("void ", "") for r in rows[2:]]
    out.append("| metric | " + " | ".join(f"`{n}`" for n in names) + " |")
    out.append("|---|" + "---|" * len(names))
    for key, label in want:
        if key not in idx:
            continue
        u = units[idx[key]]
        cells = []
        for r in rows[2:]:
            v = r[idx[key]]
            try:
                f = float(v.replace(",", ""))
                v = f"{f:.3g}" if abs(f) < 1e5 else f"{f:.4g}"
            except ValueError:
                pass
            cells.append(v)
        out.append(f"| {label} [{u}] | " + " | ".join(cells) + " |")
    out.append("")
    # per-launch DRAM traffic of each kernel for bench.py's roofline.traffic (CUB config only)
    def gb(r, key):
        v = float(r[idx[key]].replace(",", ""))
        u = units[idx[key]].lower()
        return v * {"gbyte": 1e9, "mbyte": 1e6, "kbyte": 1e3, "byte": 1.0}[u]
    traffic = {}
    for n, r in zip(names, rows[2:]):
        base = n.split("<")[0].split("::")[-1]
        traffic[base] = gb(r, "dram__bytes_read.sum") + gb(r, "dram__bytes_write.sum")
    json.dump({"workload": "cub", "B": 256, "source": f"profiles/{tag}_ncu_summary.md", "bytes_per_launch": traffic},
              open(os.path.join(ROOT, "profiles", "ncu_traffic.json"), "w"), indent=1)

# ---- bench lines of the same round
for name in ("auto", "simt", "df", "penn", "cub", "deepfashion", "pennaction", "tpsbwd"):
    bp = os.path.join(G, f"{tag}_bench_{name}.json")
    if os.path.exists(bp) and os.path.getsize(bp):
        try:
            d = json.loads(open(bp).read().strip().splitlines()[-1])
        except Exception:  # noqa: BLE001
            continue
        out.append(f"## bench ({name}): {d['config']['workload']}\n")
        out.append(f"* value {d['value']:.0f} {d['unit']}, {d['ms_per_step']:.4f} ms/step, K4 variant `{d['config'].get('decode_bwd')}`, "
                   f"step roofline frac {d['step_roofline']['frac']:.3f} of measured {d['step_roofline']['peak']} GB/s")
        out.append(f"* per C-ABI call (CUDA events, ms): {d['per_call_ms']}")
        out.append(f"* dominant kernel `{d['roofline']['kernel']}`: {d['roofline']['achieved']:.0f} GB/s algorithmic "
                   f"= {d['roofline']['frac']:.3f} of measured peak; clocks {d['clocks']}")
        if "e2e" in d:
            out.append(f"* e2e {d['e2e']['value']:.0f} {d['unit']} (H2D {d['e2e']['h2d_bytes_per_step'] / 1e6:.1f} MB, "
                       f"D2H {d['e2e']['d2h_bytes_per_step'] / 1e6:.1f} MB per step)")
        if "cpu_baseline" in d:
            out.append(f"* cpu_baseline {d['cpu_baseline']}")
        out.append("")

dst = os.path.join(ROOT, "profiles", f"{tag}_ncu_summary.md")
open(dst, "w").write(f"# ncu / bench summary, round tag `{tag}`\n\n" + "\n".join(out) + "\n")
print(dst)
