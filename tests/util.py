"""Test helpers: file parsing and golden-fixture access (no product code)."""
import gzip
import json
import os
import struct

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
GOLD = os.path.join(HERE, "golden")
ROOT = os.path.dirname(HERE)


def read_seq_file(path):
    """FASTA (multi-line ok) / FASTQ (4-line) -> list of bytes, as io.c's automaton delivers them."""
    op = gzip.open if path.endswith(".gz") else open
    data = op(path, "rb").read()
    reads = []
    if data[:1] == b"@":
        lines = data.split(b"\n")
        for i in range(0, len(lines) - 1, 4):
            if lines[i][:1] == b"@":
                reads.append(lines[i + 1])
    else:
        cur = None
        for ln in data.split(b"\n"):
            if ln[:1] == b">":
                if cur is not None:
                    reads.append(cur)
                cur = b""
            elif cur is not None:
                cur += ln
        if cur is not None:
            reads.append(cur)
    return reads


def golden_cases():
    return sorted(f[:-5] for f in os.listdir(GOLD) if f.endswith(".json"))


def golden(name):
    meta = json.load(open(os.path.join(GOLD, name + ".json")))
    src = os.path.join(GOLD, name + "." + meta["fmt"])
    blob = gzip.open(os.path.join(GOLD, name + ".ktab.gz"), "rb").read()
    sl = struct.unpack("<q", blob[:8])[0]
    meta["src"] = src
    meta["ktab_stub"] = blob[8:8 + sl]
    meta["ktab_payload"] = blob[8 + sl:]
    hist = np.zeros(32768, dtype=np.int64)
    for k, v in meta["hist_nonzero"].items():
        hist[int(k)] = v
    meta["hist"] = hist
    pz = os.path.join(GOLD, name + ".prof.npz")
    if os.path.exists(pz):
        z = np.load(pz)
        meta["prof"], meta["prof_off"] = z["prof"], z["off"]
    return meta


def read_hist_file(path):
    b = open(path, "rb").read()
    assert len(b) == 262164, len(b)
    k, lo, hi = struct.unpack("<iii", b[:12])
    ilow, maxinst = struct.unpack("<qq", b[12:28])
    h = np.zeros(32768, dtype=np.int64)
    h[1:] = np.frombuffer(b[28:], dtype="<i8")
    return dict(k=k, low=lo, high=hi, ilow=ilow, max_inst=maxinst, hist=h, raw=b)


def read_ktab_files(d, root):
    stub = open(os.path.join(d, root + ".ktab"), "rb").read()
    k, nparts, cutoff, ib = struct.unpack("<iiii", stub[:16])
    payload, ns = b"", []
    for t in range(1, nparts + 1):
        x = open(os.path.join(d, "." + root + ".ktab.%d" % t), "rb").read()
        pk, n = struct.unpack("<iq", x[:12])
        assert pk == k
        pw = ((2 * k + 7) >> 3) + 2 - ib
        assert len(x) == 12 + n * pw, (len(x), n, pw)
        ns.append(n)
        payload += x[12:]
    return dict(k=k, nparts=nparts, cutoff=cutoff, ibyte=ib, stub=stub, payload=payload, part_entries=ns)


def check_parts_on_first_byte_boundaries(kt):
    """README.md:984 / table.c:257: a first-byte value never spans two parts; entries strictly increasing."""
    k, ib = kt["k"], kt["ibyte"]
    kb = (2 * k + 7) >> 3
    idx = np.frombuffer(kt["stub"][16:], dtype="<i8")
    assert len(idx) == 256 ** ib
    assert idx[-1] == sum(kt["part_entries"])
    first_byte_cum = idx[(np.arange(256) + 1) * 256 ** (ib - 1) - 1]
    cum = np.cumsum(kt["part_entries"])
    for c in cum[:-1]:
        assert c in first_byte_cum or c == 0, "table part does not end on a first-byte boundary"
    return kb


def decode_prof_files(d, root, oracle):
    """-> (prof uint16 concat, off int64) decoding every read with the oracle's restatement of Fetch_Profile."""
    stub = open(os.path.join(d, root + ".prof"), "rb").read()
    k, nparts = struct.unpack("<ii", stub[:8])
    prof, off = [], [0]
    total = 0
    for t in range(1, nparts + 1):
        x = open(os.path.join(d, "." + root + ".pidx.%d" % t), "rb").read()
        pk, first, n = struct.unpack("<iqq", x[:20])
        assert pk == k and first == total
        idx = np.frombuffer(x[20:20 + 8 * n], dtype="<i8")
        data = open(os.path.join(d, "." + root + ".prof.%d" % t), "rb").read()
        prev = 0
        for e in idx:
            p = oracle.decode_profile(data[prev:int(e)], cap=1 << 22)
            prof.append(p)
            off.append(off[-1] + len(p))
            prev = int(e)
        total += n
    return (np.concatenate(prof) if prof else np.zeros(0, np.uint16)), np.array(off, dtype=np.int64), nparts


def golden_relative(name="rel_k40"):
    """relative-profile fixture (-p:<table>): the reference's own profiles of <name>.fa against the table it built for
    table_src at table_cutoff (tests/golden/make_golden_relative.py)"""
    d = os.path.join(GOLD, "relative")
    meta = json.load(open(os.path.join(d, name + ".json")))
    meta["src"] = os.path.join(d, name + ".fa")
    meta["table_src"] = os.path.join(GOLD, meta["table_src"])
    z = np.load(os.path.join(d, name + ".prof.npz"))
    meta["prof"], meta["prof_off"] = z["prof"], z["off"]
    return meta


def oracle_relative_profiles(oracle, table_reads, k, cutoff, reads):
    """what -p:<table> must report: per position the (saturated) count the table holds, 0 for k-mers below its cutoff"""
    import ctypes as C
    from fastk_b200.synth import to_block
    bases, boff = to_block(table_reads)
    boff = np.ascontiguousarray(boff, dtype=np.int64)
    t = oracle.lib.fko_count(bases, boff.ctypes.data_as(C.POINTER(C.c_int64)), len(table_reads), k, 0)
    try:
        out = []
        for r in reads:
            buf = np.zeros(max(len(r), 1), dtype=np.uint16)
            pl = oracle.lib.fko_profile(t, r, len(r), 0, buf.ctypes.data_as(C.POINTER(C.c_uint16)))
            p = buf[:pl].copy()
            p[p < cutoff] = 0
            out.append(p)
        return out
    finally:
        oracle.lib.fko_free_table(t)


def merge_tables_oracle(tables, kb):
    """Fastmerge.c:311-331 restated in numpy for the tests: tables = list of (n, kb+2) uint8 record arrays in key order.
    -> (merged records, hist[32768], max_inst part): counts of equal k-mers added and saturated at 32767, histogram of the
    merged counts, and for sums above 32767 the counts of their unsaturated members go to max_inst."""
    allr = np.concatenate([t for t in tables if len(t)]) if any(len(t) for t in tables) else np.zeros((0, kb + 2), np.uint8)
    pad = np.zeros((len(allr), 16), dtype=np.uint8)
    pad[:, :kb] = allr[:, :kb]
    w = pad.view(">u8")
    order = np.lexsort((w[:, 1], w[:, 0]))
    allr = allr[order]
    cnt = allr[:, kb].astype(np.int64) | (allr[:, kb + 1].astype(np.int64) << 8)
    keys = w[order]
    head = np.ones(len(allr), dtype=bool)
    if len(allr) > 1:
        head[1:] = (keys[1:] != keys[:-1]).any(axis=1)
    gid = np.cumsum(head) - 1
    ng = int(gid[-1]) + 1 if len(allr) else 0
    tot = np.bincount(gid, weights=cnt, minlength=ng).astype(np.int64)
    small = np.bincount(gid, weights=np.where(cnt < 0x7fff, cnt, 0), minlength=ng).astype(np.int64)
    sat = np.minimum(tot, 0x7fff)
    out = allr[head].copy()
    out[:, kb] = sat & 0xff
    out[:, kb + 1] = sat >> 8
    hist = np.bincount(sat, minlength=32768).astype(np.int64)
    hist[0] = 0
    return out, hist, int(small[tot > 0x7fff].sum())
