"""CPU: the product's host-side writers (fastk_b200/host/fk_files.c) fed with oracle results must reproduce
the reference's files (golden vectors): .hist bytes, .ktab stub + payload, decoded .prof."""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

import util

ROOT = util.ROOT


@pytest.fixture(scope="module")
def fkfiles(tmp_path_factory):
    d = tmp_path_factory.mktemp("fkfiles")
    so = os.path.join(str(d), "libfkfiles.so")
    subprocess.check_call(["gcc", "-O2", "-shared", "-fPIC", "-o", so, os.path.join(ROOT, "fastk_b200", "host", "fk_files.c"), "-lpthread"])
    L = C.CDLL(so)
    L.fk_write_hist.argtypes = [C.c_char_p, C.c_char_p, C.c_int, C.POINTER(C.c_int64), C.c_int64]
    L.fk_write_ktab.argtypes = [C.c_char_p, C.c_char_p, C.c_int, C.c_int, C.c_int, C.POINTER(C.c_uint8), C.c_int64]
    L.fk_write_prof.argtypes = [C.c_char_p, C.c_char_p, C.c_int, C.c_int, C.POINTER(C.c_int64), C.POINTER(C.c_int64),
                                C.POINTER(C.c_uint16)]
    L.fk_idx_bytes.argtypes = [C.c_int64, C.c_int]
    L.fk_encode_profile.argtypes = [C.POINTER(C.c_uint16), C.c_int64, C.POINTER(C.c_uint8)]
    L.fk_encode_profile.restype = C.c_int64
    L.fk_table_split.argtypes = [C.POINTER(C.c_uint8), C.c_int64, C.c_int, C.c_int, C.POINTER(C.c_int)]
    L.fk_read_ktab.argtypes = [C.c_char_p, C.POINTER(C.c_int), C.POINTER(C.c_int), C.POINTER(C.POINTER(C.c_uint8)), C.POINTER(C.c_int64)]
    L.fk_write_ktab_runs.argtypes = [C.c_char_p, C.c_char_p, C.c_int, C.c_int, C.c_int, C.POINTER(C.POINTER(C.c_uint8)),
                                     C.POINTER(C.c_int64), C.c_int]
    return L


@pytest.mark.parametrize("name,nruns", [("c1_k40", 3), ("mixed_k21", 5), ("long_k63", 2), ("c1_k40", 1)])
def test_table_delivered_as_runs_merges_to_the_reference_table(oracle_lib, fkfiles, name, nruns, tmp_path):
    """A multi-round count hands the table over as sorted runs with disjoint keys (fkgpu_result.run_table): the writer
    merges every part's slices of the runs (the role of table.c:240-313 for NPARTS part files) -- same files as the
    reference's, whatever way the keys were dealt to the runs (here: by a hash, and one run left empty)."""
    g = util.golden(name)
    reads = util.read_seq_file(g["src"])
    r = oracle_lib.count(reads, g["k"], cutoff=g["t"])
    tab = np.ascontiguousarray(r["table"], dtype=np.uint8)
    rng = np.random.default_rng(7)
    owner = rng.integers(0, max(1, nruns - 1), len(tab)) if nruns > 1 else np.zeros(len(tab), dtype=np.int64)   # last run stays empty
    runs = [np.ascontiguousarray(tab[owner == i]) for i in range(nruns)]
    ptrs = (C.POINTER(C.c_uint8) * nruns)(*[x.ctypes.data_as(C.POINTER(C.c_uint8)) for x in runs])
    ns = (C.c_int64 * nruns)(*[len(x) for x in runs])
    d = str(tmp_path).encode()
    assert fkfiles.fk_write_ktab_runs(d, b"w", g["k"], g["t"], g["T"], ptrs, ns, nruns) == 0
    kt = util.read_ktab_files(str(tmp_path), "w")
    assert kt["stub"] == g["ktab_stub"] and kt["payload"] == g["ktab_payload"]
    util.check_parts_on_first_byte_boundaries(kt)


@pytest.mark.parametrize("name", util.golden_cases())
def test_writers_reproduce_reference_files(oracle_lib, fkfiles, name, tmp_path):
    g = util.golden(name)
    reads = util.read_seq_file(g["src"])
    r = oracle_lib.count(reads, g["k"], cutoff=g["t"], profiles=True)
    d = str(tmp_path).encode()
    hist = np.ascontiguousarray(r["hist"], dtype=np.int64)
    assert fkfiles.fk_write_hist(d, b"w", g["k"], hist.ctypes.data_as(C.POINTER(C.c_int64)), r["max_inst"]) == 0
    h = util.read_hist_file(os.path.join(str(tmp_path), "w.hist"))
    assert np.array_equal(h["hist"][1:], g["hist"][1:]) and [h["k"], h["low"], h["high"], h["ilow"], h["max_inst"]] == g["hist_header"]
    tab = np.ascontiguousarray(r["table"], dtype=np.uint8)
    assert fkfiles.fk_write_ktab(d, b"w", g["k"], g["t"], g["T"], tab.ctypes.data_as(C.POINTER(C.c_uint8)), len(tab)) == 0
    kt = util.read_ktab_files(str(tmp_path), "w")
    assert kt["stub"] == g["ktab_stub"] and kt["payload"] == g["ktab_payload"]
    util.check_parts_on_first_byte_boundaries(kt)
    off = np.zeros(len(reads) + 1, dtype=np.int64)
    for i, p in enumerate(r["profiles"]):
        off[i + 1] = off[i] + len(p)
    prof = np.concatenate(r["profiles"]).astype(np.uint16) if len(reads) else np.zeros(1, np.uint16)
    nparts = 3
    rbeg = np.array([len(reads) * t // nparts for t in range(nparts + 1)], dtype=np.int64)
    assert fkfiles.fk_write_prof(d, b"w", g["k"], nparts, rbeg.ctypes.data_as(C.POINTER(C.c_int64)),
                                 off.ctypes.data_as(C.POINTER(C.c_int64)), prof.ctypes.data_as(C.POINTER(C.c_uint16))) == 0
    p2, o2, np_ = util.decode_prof_files(str(tmp_path), "w", oracle_lib)
    assert np_ == nparts and np.array_equal(o2, g["prof_off"]) and np.array_equal(p2, g["prof"])


def test_idx_bytes_rule(fkfiles):
    # count.c:1620-1626
    assert fkfiles.fk_idx_bytes(0x40000 - 1, 40) == 1
    assert fkfiles.fk_idx_bytes(0x40000, 40) == 2
    assert fkfiles.fk_idx_bytes(0x4000000, 40) == 2
    assert fkfiles.fk_idx_bytes(0x4000001, 40) == 3
    assert fkfiles.fk_idx_bytes(0x4000001, 11) == 2
    assert fkfiles.fk_idx_bytes(0x4000001, 7) == 1


def _profile_vectors():
    rng = np.random.default_rng(3)
    yield np.zeros(0, np.uint16)
    yield np.array([0], np.uint16)
    yield np.array([127], np.uint16)
    yield np.array([128], np.uint16)                                   # first count needs the 2-byte form (merge.c:541-562)
    yield np.array([32767] * 5, np.uint16)                             # saturated, zero forward differences
    yield np.zeros(200, np.uint16)                                     # zero runs longer than 63 (count.c:886-899)
    yield np.array([5] * 63 + [6] + [6] * 64 + [7], np.uint16)         # runs of exactly 62 / 63 / 64 equal counts
    yield np.array([100, 131, 100, 69, 100, 132, 100, 68], np.uint16)  # |d| = 31 / 32 on both sides of the 1-byte limit
    yield np.array([0, 32767, 0, 16384, 1, 32766], np.uint16)          # largest 15-bit differences, both signs
    for _ in range(60):
        n = int(rng.integers(1, 400))
        base = rng.integers(0, 32768)
        steps = rng.choice([0, 0, 0, 1, -1, 3, -7, 31, -31, 32, -32, 500, -500, 20000, -20000], n)
        v = np.clip(base + np.cumsum(steps), 0, 32767).astype(np.uint16)
        v[rng.random(n) < 0.05] = 0                                    # N-intervals: counts drop to 0
        yield v


def test_profile_code_round_trip_and_equals_the_restated_encoder(oracle_lib, fkfiles):
    """fk_encode_profile (product, host C) against the oracle's restatement of the reference encoder (count.c:886-921,
    merge.c:534-716) and decoder (libfastk.c:1707-1803): same bytes, and decode(encode(x)) == x."""
    for v in _profile_vectors():
        out = np.zeros(2 * len(v) + 8, np.uint8)
        nb = fkfiles.fk_encode_profile(v.ctypes.data_as(C.POINTER(C.c_uint16)), len(v), out.ctypes.data_as(C.POINTER(C.c_uint8)))
        code = out[:nb].tobytes()
        assert code == oracle_lib.encode_profile(v), list(v[:20])
        dec = oracle_lib.decode_profile(code, cap=len(v) + 4)
        assert len(dec) == len(v) and np.array_equal(dec, v), list(v[:20])


def test_table_split_rule(fkfiles):
    """fk_table_split cuts table parts on first-byte boundaries by the cumulative-threshold rule of MSDsort.c:330-352
    (Appendix C of SURVEY.md), here checked against a direct restatement on the first-byte histogram."""
    rng = np.random.default_rng(4)
    tw = 12
    for nparts in (1, 2, 4, 7):
        first = np.sort((rng.random(5000) ** 2 * 256).astype(np.uint8))       # skewed like canonical k-mers
        ent = np.zeros((len(first), tw), np.uint8)
        ent[:, 0] = first
        beg = (C.c_int * (nparts + 1))()
        fkfiles.fk_table_split(ent.ctypes.data_as(C.POINTER(C.c_uint8)), len(ent), tw, nparts, beg)
        part = np.bincount(first, minlength=256).astype(np.int64) * tw
        asize, n, s, b, want = int(part.sum()), 0, 0, 0, []
        thr = asize // nparts
        for x in range(256):
            s += int(part[x])
            if s >= thr and n < nparts:
                want.append(b)
                n += 1
                thr = asize * (n + 1) // nparts
                b = x + 1
        while len(want) < nparts:
            want.append(256)
        want.append(256)
        got = list(beg)
        assert got[0] == 0 and got[-1] == 256 and all(got[i] <= got[i + 1] for i in range(nparts)), got
        assert got == want, (nparts, got, want)


def test_remove_outputs_only_touches_its_own_files(fkfiles, tmp_path):
    """Clean_Exit's removal of partial outputs (FastK.c:181-221): exactly <root>.{hist,ktab,prof} and the numbered
    hidden parts, whatever characters the directory name holds."""
    d = os.path.join(str(tmp_path), "dir with space;echo")
    os.makedirs(d)
    mine = ["r.hist", "r.ktab", "r.prof", ".r.ktab.1", ".r.ktab.12", ".r.pidx.3", ".r.prof.3"]
    other = ["r.histx", "rr.hist", ".r.ktab.x", ".r.ktab.", ".rr.ktab.1", "keep.txt", ".r.prof.3a"]
    for f in mine + other:
        open(os.path.join(d, f), "w").write("x")
    fkfiles.fk_remove_outputs.argtypes = [C.c_char_p, C.c_char_p]
    fkfiles.fk_remove_outputs(d.encode(), b"r")
    assert sorted(os.listdir(d)) == sorted(other)


@pytest.mark.parametrize("name", util.golden_cases())
def test_table_reader_round_trip(oracle_lib, fkfiles, name, tmp_path):
    """fk_read_ktab (the -p:<table> loader of our host program): stub + hidden parts -> full [key][count] records."""
    g = util.golden(name)
    r = oracle_lib.count(util.read_seq_file(g["src"]), g["k"], cutoff=g["t"])
    tab = np.ascontiguousarray(r["table"], dtype=np.uint8)
    d = str(tmp_path).encode()
    assert fkfiles.fk_write_ktab(d, b"w", g["k"], g["t"], g["T"], tab.ctypes.data_as(C.POINTER(C.c_uint8)), len(tab)) == 0
    for nm in (os.path.join(str(tmp_path), "w"), os.path.join(str(tmp_path), "w.ktab")):
        k, cut, n = C.c_int(), C.c_int(), C.c_int64()
        rec = C.POINTER(C.c_uint8)()
        assert fkfiles.fk_read_ktab(nm.encode(), C.byref(k), C.byref(cut), C.byref(rec), C.byref(n)) == 0
        assert k.value == g["k"] and cut.value == g["t"] and n.value == len(tab)
        back = np.ctypeslib.as_array(rec, shape=(n.value * tab.shape[1],)).reshape(n.value, tab.shape[1])
        assert np.array_equal(back, tab)
    assert fkfiles.fk_read_ktab(os.path.join(str(tmp_path), "missing").encode(), C.byref(k), C.byref(cut), C.byref(rec), C.byref(n)) != 0


def test_compare_fastk_outputs_on_two_reference_runs(tmp_path):
    """formats.compare_fastk_outputs (bench.py's files arm): two runs of the compiled reference on the same FASTA with
    different -M (different hidden-part splits are allowed, T1) compare equal; a different cutoff does not."""
    import subprocess
    import numpy as np
    from fastk_b200 import formats, synth
    ref = os.path.join(ROOT, "oracle", "_ref", "FastK")
    if not os.path.exists(ref):
        pytest.skip("oracle/_ref/FastK not built")
    genome = synth.random_genome(30_000, 5)
    reads = synth.sample_reads(genome, 3000, 150, 0.005, 6)
    fa = str(tmp_path / "r.fasta")
    synth.write_fasta(reads, fa)
    for name, extra in (("a", ["-t1", "-M1"]), ("b", ["-t1", "-M4"]), ("c", ["-t2", "-M1"])):
        subprocess.check_call([ref, "-k40", "-T3", f"-P{tmp_path}", f"-N{tmp_path}/{name}"] + extra + [fa],
                              stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
    assert formats.compare_fastk_outputs(str(tmp_path), "a", "b") == []
    assert formats.compare_fastk_outputs(str(tmp_path), "a", "c") != []
    assert formats.compare_fastk_outputs(str(tmp_path), "a", "c", table=False) == []
