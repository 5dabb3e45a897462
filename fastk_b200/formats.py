"""Readers of the FastK output files (.hist, .ktab + hidden parts) and the comparison of a device result with them.

Host-side mirror of the libfastk readers (libfastk.c:51-96 Load_Histogram, libfastk.c:843-965 Load/Open_Kmer_Table);
layouts as written by count.c:1896-1909 and table.c:216-217,282-284,329-333,483-498.
Used by bench.py and the tests to byte-compare what the CUDA path produced with what the reference FastK wrote for
the same reads."""
import os
import struct

import numpy as np

HIST_BYTES = 262164


def read_hist(path):
    """-> dict(k, low, high, ilow, max_inst, hist[32768]) ; hist[c] = # distinct k-mers with count c (c >= 1)."""
    b = open(path, "rb").read()
    if len(b) != HIST_BYTES:
        raise ValueError(f"{path}: {len(b)} bytes, expected {HIST_BYTES}")
    k, lo, hi = struct.unpack("<iii", b[:12])
    ilow, maxinst = struct.unpack("<qq", b[12:28])
    h = np.zeros(32768, dtype=np.int64)
    h[1:] = np.frombuffer(b[28:], dtype="<i8")
    return dict(k=k, low=lo, high=hi, ilow=ilow, max_inst=maxinst, hist=h)


def read_ktab_stub(d, root):
    stub = open(os.path.join(d, root + ".ktab"), "rb").read()
    k, nparts, cutoff, ib = struct.unpack("<iiii", stub[:16])
    idx = np.frombuffer(stub[16:], dtype="<i8")
    if len(idx) != 256 ** ib:
        raise ValueError(f"{root}.ktab: index of {len(idx)} entries, expected {256 ** ib}")
    return dict(k=k, nparts=nparts, cutoff=cutoff, ibyte=ib, idx=idx)


def ktab_parts(d, root, stub):
    """yields the (n, kbytes - ibyte + 2) uint8 array of every hidden part, in part order (memory-mapped)."""
    k, ib = stub["k"], stub["ibyte"]
    pw = ((2 * k + 7) >> 3) + 2 - ib
    for t in range(1, stub["nparts"] + 1):
        p = os.path.join(d, "." + root + ".ktab.%d" % t)
        with open(p, "rb") as f:
            pk, n = struct.unpack("<iq", f.read(12))
        if pk != k or os.path.getsize(p) != 12 + n * pw:
            raise ValueError(f"{p}: header (k={pk}, n={n}) does not match its size")
        yield (np.memmap(p, dtype=np.uint8, mode="r", offset=12, shape=(n, pw)) if n else np.zeros((0, pw), np.uint8))


def compare_with_fastk_files(d, root, k, cutoff, hist, max_inst, table):
    """hist [32768] int64, max_inst, table (n, kbytes+2) uint8 rows [key][u16 LE count] in key order  vs  the files a
    FastK run left in directory d under `root`.  -> list of mismatch descriptions (empty = byte-identical content:
    every histogram bin, the max_inst field, the prefix index of the stub and every suffix + count byte of the parts)."""
    bad = []
    h = read_hist(os.path.join(d, root + ".hist"))
    if h["k"] != k:
        bad.append(f".hist k {h['k']} != {k}")
    if not np.array_equal(h["hist"][1:], np.asarray(hist)[1:]):
        nz = np.nonzero(h["hist"][1:] != np.asarray(hist)[1:])[0]
        bad.append(f".hist differs in {len(nz)} bins, first at count {int(nz[0]) + 1}")
    if h["max_inst"] != int(max_inst):
        bad.append(f".hist max_inst {h['max_inst']} != {int(max_inst)}")
    if h["ilow"] != int(np.asarray(hist)[1]):
        bad.append(".hist low-bin field differs")
    if table is None:
        return bad
    stub = read_ktab_stub(d, root)
    ib = stub["ibyte"]
    if stub["k"] != k or stub["cutoff"] != cutoff:
        bad.append(f".ktab stub k/cutoff {stub['k']}/{stub['cutoff']} != {k}/{cutoff}")
    n = table.shape[0]
    if int(stub["idx"][-1]) != n:
        bad.append(f".ktab holds {int(stub['idx'][-1])} entries, the device table {n}")
        return bad
    # prefix index: idx[x] = # entries whose first ibyte key bytes are <= x
    cnt = np.zeros(256 ** ib, dtype=np.int64)
    step = 1 << 24
    for a in range(0, n, step):
        t = table[a:a + step]
        p = np.zeros(len(t), dtype=np.int64)
        for j in range(ib):
            p = (p << 8) | t[:, j]
        cnt += np.bincount(p, minlength=256 ** ib)
    if not np.array_equal(np.cumsum(cnt), stub["idx"]):
        bad.append(".ktab prefix index differs")
    a = 0
    for t, part in enumerate(ktab_parts(d, root, stub), 1):
        b = a + part.shape[0]
        if b > n or not np.array_equal(table[a:b, ib:], part):
            bad.append(f".ktab part {t} payload differs")
        a = b
    if a != n:
        bad.append(".ktab parts hold a different number of entries")
    return bad


def compare_fastk_outputs(d, root_a, root_b, table=True):
    """Two FastK runs' files in directory d (tiers T0 / T1 of DESIGN.md §1): `.hist` byte-identical; `.ktab` stub
    byte-identical and the concatenated payload of the hidden parts byte-identical.  -> list of mismatches."""
    bad = []
    ha, hb = (open(os.path.join(d, r + ".hist"), "rb").read() for r in (root_a, root_b))
    if ha != hb:
        bad.append(".hist files differ")
    if not table:
        return bad
    sa, sb = (open(os.path.join(d, r + ".ktab"), "rb").read() for r in (root_a, root_b))
    if sa != sb:
        bad.append(".ktab stubs differ")
        return bad
    pa = [np.asarray(p) for p in ktab_parts(d, root_a, read_ktab_stub(d, root_a))]
    pb = [np.asarray(p) for p in ktab_parts(d, root_b, read_ktab_stub(d, root_b))]
    ca = np.concatenate(pa) if pa else np.zeros((0, 1), np.uint8)
    cb = np.concatenate(pb) if pb else np.zeros((0, 1), np.uint8)
    if ca.shape != cb.shape or not np.array_equal(ca, cb):
        bad.append(f".ktab hidden-part payloads differ ({ca.shape[0]} vs {cb.shape[0]} records)")
    return bad
