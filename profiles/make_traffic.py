#!/usr/bin/env python3
"""traffic.json from an `ncu --set full` capture: dram__bytes_read.sum + dram__bytes_write.sum per pipeline stage.

usage: make_traffic.py <report.ncu-rep> <kmers of the captured batch> [k]  > traffic.json
Stage = the launches of its kernels inside the capture window; k_sortcount runs FKGPU_D2H_CHUNKS (8) times per step, the
window may hold fewer: its sum is scaled to 8 launches.  bench.py scales every figure linearly with the k-mer count."""
import csv, json, subprocess, sys

rep, kmers = sys.argv[1], int(sys.argv[2])
k = int(sys.argv[3]) if len(sys.argv) > 3 else 40
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
hdr, units = rows[0], rows[1]
ni, ri, wi = hdr.index("Kernel Name"), hdr.index("dram__bytes_read.sum"), hdr.index("dram__bytes_write.sum")
mult = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
stage_of = [("k_super", "super_scan"), ("k_tilepart<1", "super_partition"), ("k_refine<1", "super_partition"),
            ("k_bucket_count", "bucket_count"), ("k_tilepart<2", "entry_partition"), ("k_tilepart<3", "entry_partition"),
            ("k_refine<2", "refine"), ("k_refine<3", "refine"), ("k_sortcount", "sortcount")]
tot, nsort = {}, 0
for r in rows[2:]:
    name = r[ni]
    for pat, st in stage_of:
        if pat in name:
            b = float(r[ri]) * mult[units[ri]] + float(r[wi]) * mult[units[wi]]
            tot[st] = tot.get(st, 0.0) + b
            nsort += st == "sortcount"
            break
if nsort:
    tot["sortcount"] *= 8.0 / nsort
out = {"note": "dram__bytes_read.sum + dram__bytes_write.sum per stage (a stage may be several launches) from the ncu --set full "
               "capture %s (bench.py --genome-mbp 10 --no-cpu --no-e2e; summarised in r2_ncu_top_kernels.txt); bench.py scales it "
               "linearly with the k-mer count of its own batch. sortcount: %d of its 8 launches were inside the capture window, "
               "scaled by 8/%d. The captured batch is small (its 25 M entries mostly live in the 126 MB L2), so the entry stages "
               "under-state the traffic of the full batch." % (rep, nsort, max(nsort, 1)),
       "kmer": k, "kmers": kmers}
out.update({s: int(v) for s, v in tot.items()})
print(json.dumps(out, indent=1))
