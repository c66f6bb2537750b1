#!/usr/bin/env python
"""Per-source-line summary of an ncu report: python profiles/ncu_lines.py <rep> <kernel-regex> [top]
Uses `ncu --page source --print-source cuda,sass --csv` (needs -lineinfo builds and --import-source on captures)."""
import csv
import subprocess
import sys

rep, kern = sys.argv[1], sys.argv[2]
top = int(sys.argv[3]) if len(sys.argv) > 3 else 25
txt = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass", "--kernel-name",
                      f"regex:{kern}"], capture_output=True, text=True).stdout
rows = list(csv.reader(txt.splitlines()))
hdr = next(r for r in rows if "Instructions Executed" in r)
ie, sm, src = hdr.index("Instructions Executed"), hdr.index("# Samples"), hdr.index("Source")
lines = []
for r in rows:
    if len(r) == len(hdr) and r[0].isdigit():           # CUDA source lines start with the line number
        try:
            lines.append((int(r[0]), r[src], float(r[ie] or 0), float(r[sm] or 0)))
        except ValueError:
            pass
# a line can appear several times (inlining): merge
agg = {}
for ln, s, i, m in lines:
    a = agg.setdefault((ln, s), [0.0, 0.0])
    a[0] += i
    a[1] += m
ti = sum(v[0] for v in agg.values()) or 1
tm = sum(v[1] for v in agg.values()) or 1
print(f"total warp-instructions {ti:.3e}, samples {tm:.0f}")
for (ln, s), (i, m) in sorted(agg.items(), key=lambda kv: -kv[1][0])[:top]:
    print(f"{i / ti * 100:5.1f}% instr {m / tm * 100:5.1f}% samples  L{ln:<4d} {s.strip()[:100]}")
