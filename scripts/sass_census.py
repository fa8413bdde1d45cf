#!/usr/bin/env python
"""Per-kernel census of the SASS mnemonics that prove which hardware path a kernel uses
(B200_PROFILING.md "What proves a Blackwell-native kernel").

    python scripts/sass_census.py [out.md]        # disassembles csrc/libups_b200.so with cuobjdump (no GPU needed)
"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "unsupervised-part-segmentation_b200", "csrc", "libups_b200.so")
WATCH = [("UTCHMMA", "tcgen05.mma (kind::tf32 / f16)"), ("UTCQMMA", "tcgen05.mma (fp8/fp4)"), ("LDTM", "tcgen05.ld"),
         ("STTM", "tcgen05.st"), ("UTCBAR", "tcgen05.commit"), ("UTMALDG", "TMA load"), ("UTMASTG", "TMA store"),
         ("UBLKCP", "bulk copy"), ("SYNCS", "mbarrier"), ("LDGMC", "multimem.ld_reduce"), ("STG.E.128.STRONG.SYS", "sys-scope 16-byte store"),
         ("REDG.E.ADD.F32x2", "8-byte vector reduction"), ("REDG", "global reduction (all)"), ("ATOMG", "global atomic"),
         ("HMMA", "legacy mma.sync"), ("FFMA2", "packed fp32x2 FMA"), ("LDGSTS", "cp.async")]


def main():
    out = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True, check=True).stdout
    ver = subprocess.run(["strings", LIB], capture_output=True, text=True).stdout
    m = re.search(r"ups_b200 [0-9.]+ \(sm_100a, src=[0-9a-f]+\)", ver)
    kernels = collections.OrderedDict()
    cur = None
    for line in out.splitlines():
        f = re.match(r"\s*Function : (\S+)", line)
        if f:
            cur = f.group(1)
            kernels[cur] = collections.Counter()
            continue
        if cur is None:
            continue
        ins = re.match(r"\s+/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d+\s+)?([A-Z][A-Z0-9_.]*)", line)
        if ins:
            op = ins.group(1)
            kernels[cur]["_total"] += 1
            for w, _ in WATCH:
                if op.startswith(w) or (("." in w) and w in op):
                    kernels[cur][w] += 1
    demangle = subprocess.run(["c++filt"], input="\n".join(kernels), capture_output=True, text=True).stdout.splitlines()
    rows = []
    for (mangled, c), name in zip(kernels.items(), demangle):
        short = re.sub(r"\(.*", "", name).replace("void ", "")
        hits = {w: c[w] for w, _ in WATCH if c[w]}
        rows.append((short, c["_total"], hits))
    lines = [f"# SASS census of `{os.path.relpath(LIB, ROOT)}` ({m.group(0) if m else 'version ?'})", "",
             "`cuobjdump -sass` per kernel: total instructions and the mnemonics that identify the hardware path "
             "(B200_PROFILING.md: `UTC*MMA` = tcgen05.mma, `LDTM` = tcgen05.ld, `UTMALDG` = TMA, `LDGMC` = multimem, "
             "`HMMA` = legacy mma.sync).", "",
             "| kernel | SASS instr | " + " | ".join(w for w, _ in WATCH) + " |", "|---|---|" + "---|" * len(WATCH)]
    tot = collections.Counter()
    for short, n, hits in rows:
        for w, v in hits.items():
            tot[w] += v
        if not hits:
            continue
        lines.append(f"| `{short}` | {n} | " + " | ".join(str(hits.get(w, "")) for w, _ in WATCH) + " |")
    lines.append("| **whole library** | " + str(sum(n for _, n, _ in rows)) + " | " + " | ".join(str(tot.get(w, "")) for w, _ in WATCH) + " |")
    lines += ["", "Legend: " + "; ".join(f"`{w}` = {d}" for w, d in WATCH), "",
              f"{len(rows)} kernels in the library; kernels without any of the watched mnemonics are omitted from the table."]
    text = "\n".join(lines) + "\n"
    dst = sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "profiles", "r02_sass_census.md")
    open(dst, "w").write(text)
    print(dst)


if __name__ == "__main__":
    main()
