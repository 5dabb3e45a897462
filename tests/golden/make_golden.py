#!/usr/bin/env python
"""Regenerates tests/golden/*: small inputs + the outputs of the REFERENCE ITSELF (oracle/_ref, built from
/root/reference by oracle/Makefile) on them.  Run in the build container (needs oracle/_ref):

    python tests/golden/make_golden.py

Stored per case:  <case>.fa|.fq  input,  <case>.json  {k, table cutoff, nonzero histogram bins, max_inst,
ktab stub sha/len, nparts, table entries},  <case>.ktab.gz  stub bytes + concatenated hidden-part payloads,
<case>.prof.npz  profiles as decoded by the reference's own Profex (libfastk Fetch_Profile).
"""
import gzip
import hashlib
import json
import os
import re
import struct
import subprocess
import sys
import tempfile

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
from fastk_b200 import synth  # noqa: E402

REF = os.path.join(ROOT, "oracle", "_ref")

CASES = {
    # config[0] of BASELINE.json: 1K synthetic 150 bp reads, k=40
    "c1_k40": dict(k=40, t=1, p=True, fmt="fa", T=4,
                   reads=lambda: synth.sample_reads(synth.random_genome(20_000, 101), 1000, 150, 0.005, 102)),
    "mixed_k21": dict(k=21, t=2, p=True, fmt="fq", T=3,
                      reads=lambda: synth.sample_reads(synth.random_genome(6_000, 201), 400, 100, 0.01, 202, n_rate=0.01,
                                                       lower_rate=0.3, len_jitter=90)
                      + [b"A", b"ACGT" * 5, b"N" * 60, b"acgtn" * 20, b"A" * 200, b"AC" * 100]),
    "long_k63": dict(k=63, t=1, p=True, fmt="fa", T=2,
                     reads=lambda: synth.sample_reads(synth.random_genome(8_000, 301), 30, 3_000, 0.002, 302)),
}


def read_ktab(d, root):
    stub = open(os.path.join(d, root + ".ktab"), "rb").read()
    nparts = struct.unpack("<i", stub[4:8])[0]
    payload, ns = b"", []
    for t in range(1, nparts + 1):
        x = open(os.path.join(d, "." + root + ".ktab.%d" % t), "rb").read()
        ns.append(struct.unpack("<q", x[4:12])[0])
        payload += x[12:]
    return stub, payload, ns


def read_hist(path):
    b = open(path, "rb").read()
    assert len(b) == 262164
    k, lo, hi = struct.unpack("<iii", b[:12])
    ilow, maxinst = struct.unpack("<qq", b[12:28])
    h = np.frombuffer(b[28:], dtype="<i8")
    return k, lo, hi, ilow, maxinst, h


def profex_decode(d, root, nreads):
    out = subprocess.run([os.path.join(REF, "Profex"), os.path.join(d, root), "1-#"], capture_output=True, text=True).stdout
    prof, off = [], [0]
    cur = None
    for line in out.splitlines():
        m = re.match(r"^Read (\d+):", line)
        if m:
            if cur is not None:
                off.append(len(prof))
            cur = int(m.group(1))
            continue
        m = re.match(r"^\s*(\d+):\s*(\d+)\s*$", line)
        if m and cur is not None:
            prof.append(int(m.group(2)))
    off.append(len(prof))
    assert len(off) == nreads + 1, (len(off), nreads)
    return np.array(prof, dtype=np.uint16), np.array(off, dtype=np.int64)


def main():
    if not os.path.exists(os.path.join(REF, "FastK")):
        sys.exit("oracle/_ref/FastK missing: run `make -C oracle ref` in the build container")
    for name, c in CASES.items():
        reads = c["reads"]()
        src = os.path.join(HERE, name + "." + c["fmt"])
        (synth.write_fastq if c["fmt"] == "fq" else synth.write_fasta)(reads, src)
        with tempfile.TemporaryDirectory() as d:
            cmd = [os.path.join(REF, "FastK"), "-k%d" % c["k"], "-t%d" % c["t"], "-T%d" % c["T"], "-P" + d,
                   "-N" + os.path.join(d, "out"), src]
            if c["p"]:
                cmd.insert(3, "-p")
            subprocess.check_call(cmd, stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
            k, lo, hi, ilow, maxinst, h = read_hist(os.path.join(d, "out.hist"))
            stub, payload, ns = read_ktab(d, "out")
            nz = np.nonzero(h)[0]
            meta = dict(k=c["k"], t=c["t"], T=c["T"], fmt=c["fmt"], nreads=len(reads), hist_header=[k, lo, hi, ilow, maxinst],
                        hist_nonzero={str(int(i + 1)): int(h[i]) for i in nz}, ktab_stub_len=len(stub),
                        ktab_stub_sha1=hashlib.sha1(stub).hexdigest(), ktab_part_entries=ns,
                        ktab_payload_sha1=hashlib.sha1(payload).hexdigest(), ktab_payload_len=len(payload))
            with gzip.open(os.path.join(HERE, name + ".ktab.gz"), "wb") as f:
                f.write(struct.pack("<q", len(stub)) + stub + payload)
            if c["p"]:
                prof, off = profex_decode(d, "out", len(reads))
                np.savez_compressed(os.path.join(HERE, name + ".prof.npz"), prof=prof, off=off)
                meta["prof_parts"] = struct.unpack("<i", open(os.path.join(d, "out.prof"), "rb").read()[4:8])[0]
            json.dump(meta, open(os.path.join(HERE, name + ".json"), "w"), indent=1, sort_keys=True)
        print(name, "ok:", meta["ktab_part_entries"], "entries")


if __name__ == "__main__":
    main()
