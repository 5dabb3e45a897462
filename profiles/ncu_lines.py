#!/usr/bin/env python
"""Per-source-line instruction / stall-sample shares of one kernel from an .ncu-rep (needs -lineinfo + --import-source on)."""
import csv
import subprocess
import sys


def main(rep, kernel, top=30):
    out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--print-source", "cuda,sass", "--csv",
                          "--kernel-name", "regex:" + kernel], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    his = [i for i, r in enumerate(rows) if len(r) > 8 and r[0] == "Line No"]
    if not his:
        print("no source page for a kernel matching %r in %s (was it captured with --import-source on and -lineinfo?)" % (kernel, rep))
        return
    hi = his[0]
    h = rows[hi]
    iins, ismp = h.index("Instructions Executed"), h.index("# Samples")
    data, tot, tots = [], 0, 0
    for r in rows[hi + 1:]:
        if len(r) <= iins or r[2] != "-":      # keep only the per-source-line aggregate rows
            continue
        try:
            n, s = int(r[iins]), int(r[ismp])
        except ValueError:
            continue
        tot += n
        tots += s
        data.append((n, s, r[0], r[1].strip()[:120]))
    print("total warp-inst", tot, "samples", tots)
    for n, s, ln, src in sorted(data, key=lambda x: -x[1])[:top]:
        print("%5.1f%% inst %5.1f%% smp  L%-4s | %s" % (100 * n / max(tot, 1), 100 * s / max(tots, 1), ln, src))


if __name__ == "__main__":
    main(sys.argv[1], sys.argv[2], int(sys.argv[3]) if len(sys.argv) > 3 else 30)
