#!/usr/bin/env python
"""Per-kernel SASS facts of the shipped library -> profiles/sass_summary.txt:
registers / spills / shared memory (cuobjdump --dump-resource-usage) and counts of the mnemonics that show what the code is
made of: UBLKCP (1-D bulk async copies = the TMA engine), SYNCS (mbarrier), FFMA2 / FMUL2 (packed fp32), ATOMS (shared
atomics), MEMBAR / FENCE, LDL / STL (local memory).  usage: python profiles/sass_summary.py"""
import collections
import os
import re
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "computer-vision-models_b200", "cvmhot", "lib", "libcvmhot.so")
MNEMS = ["UBLKCP", "SYNCS", "FFMA2", "FMUL2", "FADD2", "ATOMS", "MEMBAR", "FENCE", "LDL", "STL", "DFMA", "DMUL", "MUFU", "BAR"]


def demangle(n):
    try:
        return subprocess.run(["c++filt", n], capture_output=True, text=True).stdout.strip() or n
    except Exception:
        return n


res = subprocess.run(["cuobjdump", "--dump-resource-usage", LIB], capture_output=True, text=True).stdout
usage = {}
cur = None
for line in res.splitlines():
    m = re.match(r"\s*Function (\S+):", line)
    if m:
        cur = m.group(1)
    elif cur and "REG:" in line:
        usage[cur] = line.strip()
sass = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True).stdout
counts = collections.defaultdict(collections.Counter)
n_instr = collections.Counter()
cur = None
for line in sass.splitlines():
    m = re.match(r"\s*Function : (\S+)", line)
    if m:
        cur = m.group(1)
        continue
    if cur and re.match(r"\s+/\*[0-9a-f]+\*/", line):
        n_instr[cur] += 1
        for mn in MNEMS:
            if re.search(r"\b" + mn + r"\b|\b" + mn + r"\.", line):
                counts[cur][mn] += 1
out = ["arch: " + ", ".join(sorted(set(re.findall(r"arch = (sm_\w+)", sass))))]
for fn in sorted(n_instr, key=lambda f: demangle(f)):
    name = re.sub(r"\(anonymous namespace\)::", "", demangle(fn))
    name = re.sub(r"\((anonymous namespace::)?\w+Params\)|\(.*\)$", "", name)
    out.append(f"{name[:70]:70s} instr {n_instr[fn]:6d}  {usage.get(fn, '')}")
    out.append("    " + "  ".join(f"{k} {v}" for k, v in counts[fn].items() if v))
txt = "\n".join(out) + "\n"
open(os.path.join(ROOT, "profiles", "sass_summary.txt"), "w").write(txt)
print(txt)
