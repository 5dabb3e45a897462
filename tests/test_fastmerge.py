"""Fastmerge.  CPU: the numpy restatement of the reference's merge rule (tests/util.merge_tables_oracle,
Fastmerge.c:311-331) is pinned against the reference's own Fastmerge binary on tables its own FastK wrote.  GPU: the library's
fkgpu_merge_tables and our Fastmerge command line against both."""
import os
import subprocess

import numpy as np
import pytest

import util
from fastk_b200 import formats, synth

OURS = os.path.join(util.ROOT, "fastk_b200", "bin", "Fastmerge")


def part_reads():
    genome = synth.random_genome(15_000, 501)
    # FOUR parts: with three or fewer the reference tool skips the initial heap build (Fastmerge.c:302-304) and writes the first
    # k-mer of every thread's range once per table holding it (observed here: duplicate keys with partial counts in its own
    # output) -- with four or more it merges as specified, and that is the behaviour restated and tested
    parts = [synth.sample_reads(genome, 400 - 30 * i, 150, 0.004, 510 + i, n_rate=0.002) for i in range(4)]
    hot = bytes(b"ACGT"[x] for x in np.random.default_rng(3).integers(0, 4, 60))
    parts[0] += [hot] * 20000          # its k-mers saturate in two of the parts: the max_inst rule of Fastmerge.c:321-327
    parts[1] += [hot] * 30000
    parts[2] += [hot] * 5
    return parts


def ref_tables(ref_bin, d, k, cutoffs):
    """the reference FastK on every part -> [(root, records, hist dict)]"""
    out = []
    for i, (reads, t) in enumerate(zip(part_reads(), cutoffs)):
        fa = os.path.join(d, "part%d.fa" % i)
        synth.write_fasta(reads, fa)
        root = os.path.join(d, "p%d" % i)
        subprocess.check_call([os.path.join(ref_bin, "FastK"), "-k%d" % k, "-t%d" % t, "-T2", "-P" + d, "-N" + root, fa],
                              stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
        stub = formats.read_ktab_stub(d, "p%d" % i)
        kb = (2 * k + 7) >> 3
        ib = stub["ibyte"]
        recs = []
        idx = stub["idx"]
        pre = np.repeat(np.arange(len(idx)), np.diff(np.concatenate(([0], idx))))
        suf = np.concatenate(list(formats.ktab_parts(d, "p%d" % i, stub)))
        full = np.zeros((len(suf), kb + 2), dtype=np.uint8)
        for b in range(ib):
            full[:, b] = (pre >> (8 * (ib - 1 - b))) & 0xff
        full[:, ib:] = suf
        out.append((root, full, formats.read_hist(root + ".hist")))
    return out


def read_merged(d, root, k):
    stub = formats.read_ktab_stub(d, root)
    payload = np.concatenate(list(formats.ktab_parts(d, root, stub)))
    return stub, payload, formats.read_hist(os.path.join(d, root + ".hist"))


def test_merge_rule_matches_reference_fastmerge(ref_bin, tmp_path):
    if ref_bin is None or not os.path.exists(os.path.join(ref_bin, "Fastmerge")):
        pytest.skip("oracle/_ref/Fastmerge not built")
    d, k = str(tmp_path), 40
    parts = ref_tables(ref_bin, d, k, (1, 2, 1, 1))
    subprocess.check_call([os.path.join(ref_bin, "Fastmerge"), "-ht", "-T3", os.path.join(d, "m")] + [p[0] for p in parts],
                          stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
    stub, payload, h = read_merged(d, "m", k)
    kb = (2 * k + 7) >> 3
    want, hist, extra = util.merge_tables_oracle([p[1] for p in parts], kb)
    assert int(stub["idx"][-1]) == len(want) and np.array_equal(payload, want[:, stub["ibyte"]:])
    assert np.array_equal(h["hist"][1:], hist[1:])
    assert h["max_inst"] == extra + sum(p[2]["max_inst"] for p in parts)
    assert stub["cutoff"] == 1


@pytest.mark.gpu
def test_gpu_merge_tables_api(ref_bin, tmp_path):
    from fastk_b200 import FastKGPU
    if ref_bin is None:
        pytest.skip("oracle/_ref not built")
    d, k = str(tmp_path), 40
    parts = ref_tables(ref_bin, d, k, (1, 2, 1, 1))
    kb = (2 * k + 7) >> 3
    want, hist, extra = util.merge_tables_oracle([p[1] for p in parts], kb)
    eng = FastKGPU(k=k, table_cutoff=1)
    try:
        got = eng.merge_tables([p[1] for p in parts])
        assert got.ntable == len(want) and np.array_equal(got.table, want)
        assert np.array_equal(got.hist[1:], hist[1:]) and got.max_inst == extra
        one = eng.merge_tables([parts[0][1]])                 # a single table merges to itself
        assert np.array_equal(one.table, parts[0][1])
    finally:
        eng.close()


@pytest.mark.gpu
@pytest.mark.parametrize("k", [40, 21, 63])
def test_gpu_fastmerge_cli_equals_reference(ref_bin, tmp_path, k):
    if ref_bin is None or not os.path.exists(os.path.join(ref_bin, "Fastmerge")):
        pytest.skip("oracle/_ref/Fastmerge not built")
    assert os.path.exists(OURS), "host program not built"
    d = str(tmp_path)
    parts = ref_tables(ref_bin, d, k, (1, 1, 3, 2))
    subprocess.check_call([os.path.join(ref_bin, "Fastmerge"), "-ht", "-T3", os.path.join(d, "ref")] + [p[0] for p in parts],
                          stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
    r = subprocess.run([OURS, "-ht", "-T3", os.path.join(d, "gpu")] + [p[0] + ".ktab" for p in parts], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    a, b = read_merged(d, "ref", k), read_merged(d, "gpu", k)
    assert np.array_equal(a[0]["idx"], b[0]["idx"]) and a[0]["cutoff"] == b[0]["cutoff"] and a[0]["ibyte"] == b[0]["ibyte"]
    assert np.array_equal(a[1], b[1]), "merged table payload differs from the reference Fastmerge"
    assert open(os.path.join(d, "ref.hist"), "rb").read() == open(os.path.join(d, "gpu.hist"), "rb").read()
