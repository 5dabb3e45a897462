#!/usr/bin/env python
"""Static evidence (no GPU): per kernel of libfastk_gpu.so registers / stack / static smem (cuobjdump --dump-resource-usage)
and counts of the SASS mnemonics that matter here -- UBLKCP (TMA bulk copy), SYNCS (mbarrier), ATOMS / ATOMG / RED
(shared / global atomics), SHFL / REDUX (warp collectives), BAR (CTA barriers), LDG.E.128 (128-bit global loads)."""
import collections
import re
import subprocess
import sys

so = sys.argv[1] if len(sys.argv) > 1 else "fastk_b200/lib/libfastk_gpu.so"
filt = lambda s: subprocess.run(["c++filt", s], capture_output=True, text=True).stdout.strip()
short = lambda s: re.sub(r"\(.*", "", filt(s)).replace("void ", "").replace("fk::", "")
res = {}
out = subprocess.run(["cuobjdump", "--dump-resource-usage", so], capture_output=True, text=True).stdout
fn = None
for line in out.splitlines():
    m = re.match(r"\s*Function (\S+):", line)
    if m:
        fn = short(m.group(1))
    m = re.search(r"REG:(\d+) STACK:(\d+) SHARED:(\d+)", line)
    if m and fn:
        res[fn] = tuple(int(x) for x in m.groups())
cnt = collections.defaultdict(collections.Counter)
sass = subprocess.run(["cuobjdump", "-sass", so], capture_output=True, text=True).stdout
fn = None
for line in sass.splitlines():
    m = re.search(r"Function : (\S+)", line)
    if m:
        fn = short(m.group(1))
        continue
    m = re.search(r"\b(UBLKCP|SYNCS|ATOMS|ATOMG|RED|SHFL|REDUX|BAR|LDGSTS)\b", line)
    if m and fn:
        cnt[fn][m.group(1)] += 1
    if fn and re.search(r"\bLDG\.E\.(128|ENL2\.128)|LDG\.E\.[A-Z.]*128", line):
        cnt[fn]["LDG.128"] += 1
cols = ["UBLKCP", "SYNCS", "ATOMS", "ATOMG", "RED", "SHFL", "REDUX", "BAR", "LDG.128"]
print("%-58s %4s %5s %6s  " % ("kernel", "regs", "stack", "ssmem") + " ".join("%7s" % c for c in cols))
for fn in sorted(res):
    r = res[fn]
    print("%-58s %4d %5d %6d  " % (fn[:58], r[0], r[1], r[2]) + " ".join("%7d" % cnt[fn][c] for c in cols))
