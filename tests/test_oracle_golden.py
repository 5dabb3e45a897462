"""CPU: the oracle (oracle/fastk_oracle.c) against (a) the committed golden vectors produced by the reference
itself and (b), when oracle/_ref is present, the reference run live on extra cases."""
import hashlib
import os
import subprocess

import numpy as np
import pytest

import util
from fastk_b200 import synth


@pytest.mark.parametrize("name", util.golden_cases())
def test_oracle_matches_golden(oracle_lib, name, tmp_path):
    g = util.golden(name)
    reads = util.read_seq_file(g["src"])
    assert len(reads) == g["nreads"]
    got = oracle_lib.count(reads, g["k"], cutoff=g["t"], profiles=True)
    assert np.array_equal(got["hist"][1:], g["hist"][1:])
    assert got["max_inst"] == g["hist_header"][4]
    assert got["hist"][1] == g["hist_header"][3]
    # files written by the oracle's own writers: stub and concatenated payload byte-identical to the reference's
    oracle_lib.run_files([g["src"]], str(tmp_path), "o", g["k"], table=g["t"], profile=True, nparts=g["T"])
    kt = util.read_ktab_files(str(tmp_path), "o")
    assert kt["stub"] == g["ktab_stub"]
    assert hashlib.sha1(kt["payload"]).hexdigest() == g["ktab_payload_sha1"]
    assert kt["payload"] == g["ktab_payload"]
    assert kt["nparts"] == g["T"]
    util.check_parts_on_first_byte_boundaries(kt)
    h = util.read_hist_file(os.path.join(str(tmp_path), "o.hist"))
    assert np.array_equal(h["hist"][1:], g["hist"][1:]) and h["max_inst"] == g["hist_header"][4]
    assert [h["k"], h["low"], h["high"]] == g["hist_header"][:3]
    # profiles: decoded values identical to the reference's Fetch_Profile output
    prof, off, _ = util.decode_prof_files(str(tmp_path), "o", oracle_lib)
    assert np.array_equal(off, g["prof_off"])
    assert np.array_equal(prof, g["prof"])
    for r in range(len(reads)):
        assert np.array_equal(got["profiles"][r], g["prof"][g["prof_off"][r]:g["prof_off"][r + 1]])


def test_profile_code_roundtrip(oracle_lib):
    rng = np.random.default_rng(0)
    for trial in range(200):
        n = int(rng.integers(1, 400))
        kind = trial % 4
        if kind == 0:
            p = rng.integers(0, 32768, n)
        elif kind == 1:
            p = np.repeat(rng.integers(0, 200, max(1, n // 70 + 1)), 70)[:n]
        elif kind == 2:
            p = np.clip(np.cumsum(rng.integers(-40, 41, n)) + 100, 0, 32767)
        else:
            p = np.where(rng.random(n) < 0.1, 0, 32767)
        p = p.astype(np.uint16)
        code = oracle_lib.encode_profile(p)
        assert np.array_equal(oracle_lib.decode_profile(code), p)


REF_CASES = [
    ("bc", dict(k=25, t=1, extra=["-bc12"], bc=12, hoco=False)),
    ("hoco", dict(k=31, t=1, extra=["-c"], bc=0, hoco=True)),
    ("multiline", dict(k=40, t=3, extra=[], bc=0, hoco=False)),
    ("saturate", dict(k=40, t=1, extra=[], bc=0, hoco=False)),
]


@pytest.mark.parametrize("name,c", REF_CASES)
def test_oracle_matches_reference_live(oracle_lib, ref_bin, tmp_path, name, c):
    if ref_bin is None:
        pytest.skip("oracle/_ref not built (no /root/reference on this box and no prebuilt binaries)")
    if name == "saturate":
        rng = np.random.default_rng(5)
        a = bytes(b"ACGT"[x] for x in rng.integers(0, 4, 45))
        b = bytes(b"ACGT"[x] for x in rng.integers(0, 4, 50))
        reads = [a] * 40000 + [a[:42]] * 3 + [b] * 32767
    else:
        reads = synth.sample_reads(synth.random_genome(30_000, 7), 1500, 140, 0.004, 8, n_rate=0.002, len_jitter=60)
        reads += [b"AAAAAAAACCCCCCCCGGGGGGGGTTTTTTTT" * 8, b"ACGT" * 3]
    src = os.path.join(str(tmp_path), "in.fa")
    synth.write_fasta(reads, src, width=70 if name == "multiline" else 0)
    rd = os.path.join(str(tmp_path), "ref")
    od = os.path.join(str(tmp_path), "orc")
    os.makedirs(rd)
    os.makedirs(od)
    subprocess.check_call([os.path.join(ref_bin, "FastK"), "-k%d" % c["k"], "-t%d" % c["t"], "-p", "-T4", "-P" + rd,
                           "-N" + os.path.join(rd, "x")] + c["extra"] + [src],
                          stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
    oracle_lib.run_files([src], od, "x", c["k"], table=c["t"], profile=True, bc=c["bc"], compress=c["hoco"], nparts=4)
    assert open(os.path.join(rd, "x.hist"), "rb").read() == open(os.path.join(od, "x.hist"), "rb").read()
    a, b = util.read_ktab_files(rd, "x"), util.read_ktab_files(od, "x")
    assert a["stub"] == b["stub"] and a["payload"] == b["payload"]
    pa, oa, _ = util.decode_prof_files(rd, "x", oracle_lib)
    pb, ob, _ = util.decode_prof_files(od, "x", oracle_lib)
    assert np.array_equal(oa, ob) and np.array_equal(pa, pb)
    if name == "saturate":
        h = util.read_hist_file(os.path.join(od, "x.hist"))
        assert h["hist"][32767] == 17 and h["max_inst"] == 600446


def test_oracle_relative_profiles_match_reference(oracle_lib):
    """-p:<table> (FastK.c:269-281): the oracle's restatement -- look every canonical k-mer up in the table, 0 if absent --
    against the profiles the reference itself produced (tests/golden/relative, decoded by its Profex)."""
    g = util.golden_relative()
    want = util.oracle_relative_profiles(oracle_lib, util.read_seq_file(g["table_src"]), g["k"], g["table_cutoff"],
                                         util.read_seq_file(g["src"]))
    assert len(want) == g["nreads"] == len(g["prof_off"]) - 1
    for r, p in enumerate(want):
        assert np.array_equal(p, g["prof"][g["prof_off"][r]:g["prof_off"][r + 1]]), f"relative profile of read {r} differs"
