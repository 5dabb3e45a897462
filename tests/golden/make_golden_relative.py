#!/usr/bin/env python
"""Regenerates tests/golden/relative/*: RELATIVE profiles (-p:<table>, FastK.c:269-281) as the REFERENCE ITSELF (oracle/_ref)
produces them.  Run in the build container (needs oracle/_ref):

    python tests/golden/make_golden_relative.py

The table is what the reference writes for tests/golden/c1_k40.fa with -k40 -t2 (so its count-1 k-mers are ABSENT and
profile as 0); the query reads are other samples of the same genome, random reads, and edge cases.  Stored: rel_k40.fa
(query), rel_k40.json (table arguments, # reads, # profile parts), rel_k40.prof.npz (profiles decoded by the reference's Profex).
"""
import json
import os
import struct
import subprocess
import sys
import tempfile

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, HERE)
from fastk_b200 import synth  # noqa: E402
from make_golden import REF, profex_decode  # noqa: E402

OUT = os.path.join(HERE, "relative")


def main():
    if not os.path.exists(os.path.join(REF, "FastK")):
        sys.exit("oracle/_ref/FastK missing: run `make -C oracle ref` in the build container")
    os.makedirs(OUT, exist_ok=True)
    genome = synth.random_genome(20_000, 101)                 # the genome behind c1_k40.fa
    reads = synth.sample_reads(genome, 150, 180, 0.004, 777, n_rate=0.002, lower_rate=0.2, len_jitter=120)
    reads += synth.sample_reads(synth.random_genome(5_000, 55), 30, 150, 0.0, 56)
    reads += [b"", b"ACGT", b"N" * 90, b"A" * 120, bytes(b"ACGT"[x] for x in genome[100:139]), bytes(b"ACGT"[x] for x in genome[100:140])]
    src = os.path.join(OUT, "rel_k40.fa")
    synth.write_fasta(reads, src)
    table_src = os.path.join(HERE, "c1_k40.fa")
    meta = dict(k=40, table_src="c1_k40.fa", table_cutoff=2, T=2, nreads=len(reads))
    with tempfile.TemporaryDirectory() as d:
        subprocess.check_call([os.path.join(REF, "FastK"), "-k40", "-t2", "-T4", "-P" + d, "-N" + os.path.join(d, "tab"), table_src],
                              stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
        subprocess.check_call([os.path.join(REF, "FastK"), "-k40", "-p:" + os.path.join(d, "tab"), "-T2", "-P" + d,
                               "-N" + os.path.join(d, "out"), src], stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
        assert not os.path.exists(os.path.join(d, "out.hist")) and not os.path.exists(os.path.join(d, "out.ktab"))
        prof, off = profex_decode(d, "out", len(reads))
        meta["prof_parts"] = struct.unpack("<i", open(os.path.join(d, "out.prof"), "rb").read()[4:8])[0]
    np.savez_compressed(os.path.join(OUT, "rel_k40.prof.npz"), prof=prof, off=off)
    json.dump(meta, open(os.path.join(OUT, "rel_k40.json"), "w"), indent=1, sort_keys=True)
    print("rel_k40 ok:", len(reads), "reads,", len(prof), "profile values,", int((prof > 0).sum()), "non-zero")


if __name__ == "__main__":
    main()
