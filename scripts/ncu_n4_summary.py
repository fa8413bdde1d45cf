#!/usr/bin/env python
"""Summarise the N4 ncu capture into profiles/<tag>_n4_ncu_summary.md.

    python scripts/ncu_n4_summary.py <tag>     # reads gpurun_out/<tag>_n4_prof.ncu-rep
(`ncu --set full --clock-control none --import-source on -k regex:"inject_conv|parts_conv"` over
`python scripts/bench_inject_conv.py --no-library --once`: one launch of every N4 kernel at the CUB bench shape)
"""
import csv
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
tag = sys.argv[1]
rep = os.path.join(ROOT, "gpurun_out", f"{tag}_n4_prof.ncu-rep")
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
hdr, units = rows[0], rows[1]
idx = {h: i for i, h in enumerate(hdr)}
want = [("gpu__time_duration.sum", "duration"),
        ("dram__bytes_read.sum", "DRAM read"), ("dram__bytes_write.sum", "DRAM write"),
        ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "DRAM %"),
        ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue %"),
        ("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "tensor pipe %"),
        ("l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed", "LSU smem %"),
        ("sm__warps_active.avg.pct_of_peak_sustained_active", "warps active %"),
        ("launch__registers_per_thread", "regs"), ("launch__grid_size", "grid"),
        ("smsp__inst_executed.sum", "warp instr")]
out = [f"# N4 kernels — ncu full capture ({tag}_n4_prof.ncu-rep)\n",
       "`ncu --set full --clock-control none --import-source on -k regex:\"inject_conv|parts_conv\"` over "
       "`python scripts/bench_inject_conv.py --no-library --once` (CUB B=256, 128², K=16, F=64, Co=32; one launch per kernel; "
       "durations are cold-cache and serialised — the CUDA-event numbers are in the bench JSON).\n",
       "| kernel | " + " | ".join(n for _, n in want) + " |", "|---|" + "---|" * len(want)]
seen = set()
for r in rows[2:]:
    name = r[idx["Kernel Name"]].split("(const")[0].split("(float")[0].replace("void ", "").split("::")[-1]
    if name in seen:
        continue
    seen.add(name)
    cells = []
    for m, _ in want:
        if m not in idx:
            cells.append("-")
            continue
        v, u = r[idx[m]], units[idx[m]]
        try:
            f = float(v.replace(",", ""))
            v = f"{f:.3f}" if f < 100 else f"{f:.0f}"
        except ValueError:
            pass
        cells.append(f"{v} {u}".strip() if u not in ("%", "") else v)
    out.append(f"| `{name}` | " + " | ".join(cells) + " |")
out.append("")
open(os.path.join(ROOT, "profiles", f"{tag}_n4_ncu_summary.md"), "w").write("\n".join(out))
print("\n".join(out))
