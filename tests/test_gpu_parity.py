"""GPU parity: the CUDA path, called through the C ABI, against the CPU oracle on the same seeded reads.
Bit-exact bar: histogram, max_inst, and the [key][count] table records must be identical."""
import numpy as np
import pytest

from fastk_b200 import FastKGPU
from fastk_b200 import synth

pytestmark = pytest.mark.gpu


def run_gpu(reads, k, cutoff=1, bc=0, nthreads=1, block_bytes=1_000_000, reserve=0):
    g = FastKGPU(k=k, table_cutoff=cutoff, bc_prefix=bc, nthreads=nthreads, reserve_bases=reserve)
    try:
        for i, (bases, boff) in enumerate(synth.blocks(reads, max_bytes=block_bytes)):
            g.ingest(bases, boff.astype(np.int32), tid=i % nthreads)
        return g.finish(fetch_table=True)
    finally:
        g.close()


def check(oracle_lib, reads, k, cutoff=1, bc=0, nthreads=1, **kw):
    want = oracle_lib.count(reads, k, bc_prefix=bc, cutoff=cutoff)
    got = run_gpu(reads, k, cutoff=cutoff, bc=bc, nthreads=nthreads, **kw)
    assert got.nkmers == want["nkmers"]
    assert got.ndistinct == want["ndistinct"]
    assert got.max_inst == want["max_inst"]
    assert np.array_equal(got.hist[1:], want["hist"][1:])
    assert got.ntable == len(want["table"])
    if got.ntable:
        assert np.array_equal(got.table, want["table"])
    return got


@pytest.mark.parametrize("k", [40, 21, 63, 32, 33, 64, 7, 17, 18, 48, 49, 56, 57])
def test_config1_1k_reads(oracle_lib, k):
    genome = synth.random_genome(20_000, 11)
    reads = synth.sample_reads(genome, 1000, 150, 0.005, 12)
    check(oracle_lib, reads, k)


def test_edge_cases(oracle_lib):
    genome = synth.random_genome(5_000, 3)
    reads = synth.sample_reads(genome, 300, 120, 0.01, 4, n_rate=0.01, lower_rate=0.3, len_jitter=100)
    reads += [b"", b"A", b"ACGT" * 9 + b"ACG", b"N" * 200, b"acgtn" * 50, b"A" * 500, b"AC" * 300,
              b"ACGTACGTAC" * 30, b"T" * 40, b"G" * 39]
    check(oracle_lib, reads, 40)
    check(oracle_lib, reads, 21, cutoff=2)
    check(oracle_lib, reads, 40, bc=10)


def test_saturation_known_answer(oracle_lib):
    """40 000 copies of a 45-mer + 3 copies of its first 42 bases + 32 767 copies of a 50-mer."""
    rng = np.random.default_rng(5)
    a = bytes(b"ACGT"[x] for x in rng.integers(0, 4, 45))
    b = bytes(b"ACGT"[x] for x in rng.integers(0, 4, 50))
    reads = [a] * 40000 + [a[:42]] * 3 + [b] * 32767
    got = check(oracle_lib, reads, 40)
    assert got.hist[32767] == 17
    assert got.max_inst == 3 * 40003 + 3 * 40000 + 11 * 32767
    assert (got.table[:, -2].astype(int) | (got.table[:, -1].astype(int) << 8) == 32767).all()


def test_medium_30x(oracle_lib):
    genome = synth.random_genome(200_000, 21)
    reads = synth.sample_reads(genome, 40_000, 150, 0.002, 22)
    check(oracle_lib, reads, 40, nthreads=4)
    check(oracle_lib, reads, 21, cutoff=4, nthreads=3)


@pytest.mark.parametrize("k,bc,reserve_frac", [(40, 0, 1.2), (21, 0, 1.2), (40, 0, 0.25), (63, 0, 1.2), (40, 10, 1.2)])
def test_streamed_front_end(oracle_lib, monkeypatch, k, bc, reserve_frac):
    """reserve_bases > 0: every staging chunk is packed (and, on the super-mer path, scanned into super-mer records) as
    soon as it lands on the device.  Many small chunks from 3 ingest threads; reserve_frac < 1 makes the reservation too
    small, so the stream is abandoned half way and finish packs + scans the whole buffer instead."""
    monkeypatch.setenv("FKGPU_CHUNK_BYTES", str(128 << 10))
    genome = synth.random_genome(150_000, 61)
    reads = synth.sample_reads(genome, 20_000, 150, 0.003, 62, n_rate=0.001, len_jitter=60)
    total = sum(len(r) + 1 for r in reads)
    check(oracle_lib, reads, k, bc=bc, nthreads=3, block_bytes=40_000, reserve=int(total * reserve_frac))


@pytest.mark.parametrize("k,reserve_frac,profile", [(40, 1.3, False), (21, 1.3, False), (40, 0.25, False), (63, 1.3, False), (40, 1.3, True)])
def test_direct_ingest_from_page_locked_blocks(oracle_lib, monkeypatch, k, reserve_frac, profile):
    """DATA_BLOCKs in page-locked host memory are DMA'd straight into device regions (no staging memcpy); each complete
    region is packed + scanned.  3 ingest threads with interleaved regions; reserve_frac < 1 runs out of reserved space
    half way, so later blocks fall back to the staging path and the stream is abandoned."""
    import torch
    monkeypatch.setenv("FKGPU_CHUNK_BYTES", str(128 << 10))
    genome = synth.random_genome(150_000, 71)
    reads = synth.sample_reads(genome, 20_000, 150, 0.003, 72, n_rate=0.001, len_jitter=60)
    total = sum(len(r) + 1 for r in reads)
    nthreads = 3
    want = oracle_lib.count(reads, k, cutoff=1, profiles=profile)
    g = FastKGPU(k=k, table_cutoff=1, profile=profile, nthreads=nthreads, reserve_bases=int(total * reserve_frac))
    try:
        slabs = [torch.empty(64 << 10, dtype=torch.uint8, pin_memory=True) for _ in range(nthreads)]
        for t in range(nthreads):                      # tid-major read order, as io.c delivers it
            mine = reads[len(reads) * t // nthreads: len(reads) * (t + 1) // nthreads]
            for bases, boff in synth.blocks(mine, max_bytes=40_000):
                slab = slabs[t]
                slab[:len(bases)] = torch.frombuffer(bytearray(bases), dtype=torch.uint8)
                b32 = np.ascontiguousarray(boff, dtype=np.int32)
                g.ingest_ptr(slab.data_ptr(), b32.ctypes.data, len(b32) - 1, tid=t)
        got = g.finish(fetch_table=True)
        assert got.nkmers == want["nkmers"] and got.ndistinct == want["ndistinct"] and got.max_inst == want["max_inst"]
        assert np.array_equal(got.hist[1:], want["hist"][1:])
        assert np.array_equal(got.table, want["table"])
        if profile:
            off, prof = g.profiles()
            assert len(off) == len(reads) + 1
            for r in range(len(reads)):
                assert np.array_equal(prof[off[r]:off[r + 1]], want["profiles"][r]), f"profile of read {r} differs"
    finally:
        g.close()


@pytest.mark.parametrize("bbits", [24, 3])
def test_forced_bucket_geometry(bbits):
    """The bucket-id width is a run-time property of the record layout (it grows with the global input at 4 / 8 GPUs:
    24 bits, 13-bit second partition level).  Force the extremes on one GPU -- in a fresh process, the override is read
    once -- and re-run two parity tests: almost-empty buckets (24) and heavily overflowing ones (3: every group
    exceeds the on-chip pool and is re-run in residue classes)."""
    import os
    import subprocess
    import sys
    here = os.path.dirname(os.path.abspath(__file__))
    env = dict(os.environ, FKGPU_BBITS=str(bbits))
    r = subprocess.run([sys.executable, "-m", "pytest", "-q", "-m", "gpu", "-x",
                        os.path.join(here, "test_gpu_parity.py") + "::test_medium_30x",
                        os.path.join(here, "test_gpu_parity.py") + "::test_long_reads_hifi_like"],
                       capture_output=True, text=True, timeout=900, env=env, cwd=os.path.dirname(here))
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-2000:]


def test_everything_spills_to_the_record_pipeline():
    """FKGPU_BIG=48: every bucket group of more than 48 super-mers is handed to the record pipeline (the path giant buckets
    take: high-copy repeats, low-complexity runs), the rest is counted on chip, and both meet again in the entries.  Re-runs
    parity tests in a fresh process (the override is read once): counts, saturation, profiles, streamed ingest."""
    import os
    import subprocess
    import sys
    here = os.path.dirname(os.path.abspath(__file__))
    env = dict(os.environ, FKGPU_BIG="48")
    me = os.path.join(here, "test_gpu_parity.py")
    r = subprocess.run([sys.executable, "-m", "pytest", "-q", "-m", "gpu", "-x", me + "::test_medium_30x", me + "::test_long_reads_hifi_like",
                        me + "::test_saturation_known_answer", me + "::test_profiles_match_oracle", me + "::test_edge_cases",
                        me + "::test_multi_round_count_leaves_disjoint_sorted_runs"],
                       capture_output=True, text=True, timeout=1500, env=env, cwd=os.path.dirname(here))
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-2000:]


def test_repeat_family_and_low_complexity_reads(oracle_lib):
    """A 400 bp repeat unit in 60 000 reads (its k-mers saturate: count > 32767), poly-A and dinucleotide reads, on a random
    background: the repeat's minimizer buckets hold tens of thousands of super-mers each -- far beyond what one CTA
    should grind through -- and leave the chip for the record pipeline; nothing errors out, everything matches the oracle."""
    rng = np.random.default_rng(19)
    unit = bytes(b"ACGT"[x] for x in rng.integers(0, 4, 400))
    flank = lambda: bytes(b"ACGT"[x] for x in rng.integers(0, 4, 30))           # noqa: E731
    reads = [flank() + unit + flank() for _ in range(60_000)]
    reads += [b"A" * 150] * 3000 + [b"AC" * 75] * 2000 + [b"T" * 200] * 500
    reads += synth.sample_reads(synth.random_genome(100_000, 5), 10_000, 150, 0.003, 6)
    g = FastKGPU(k=40, table_cutoff=1, nthreads=2)
    try:
        want = oracle_lib.count(reads, 40, cutoff=1)
        for i, (bases, boff) in enumerate(synth.blocks(reads)):
            g.ingest(bases, boff.astype(np.int32), tid=i % 2)
        got = g.finish(fetch_table=True)
        st = g.last_stats()
        assert st["path"] == 1 and st["spilled_kmers"] > 0, st
        assert got.nkmers == want["nkmers"] and got.ndistinct == want["ndistinct"] and got.max_inst == want["max_inst"]
        assert np.array_equal(got.hist[1:], want["hist"][1:])
        assert np.array_equal(got.table, want["table"])
        assert got.hist[32767] > 300
    finally:
        g.close()


def test_long_reads_hifi_like(oracle_lib):
    genome = synth.random_genome(300_000, 31)
    reads = synth.sample_reads(genome, 400, 15_000, 0.001, 32)
    check(oracle_lib, reads, 40)
    check(oracle_lib, reads, 63)


def test_repeats_force_refinement(oracle_lib):
    """Heavy repeats: oversize work groups take the host-driven MSD refinement path."""
    rng = np.random.default_rng(9)
    unit = bytes(b"ACGT"[x] for x in rng.integers(0, 4, 300))
    reads = [unit * 3] * 3000 + synth.sample_reads(synth.random_genome(50_000, 1), 2000, 150, 0.01, 2)
    check(oracle_lib, reads, 40)
    check(oracle_lib, reads, 25)


def run_gpu_profiles(reads, k, bc=0, nthreads=1, block_bytes=1_000_000):
    g = FastKGPU(k=k, table_cutoff=1, profile=True, bc_prefix=bc, nthreads=nthreads)
    try:
        per_tid = [[] for _ in range(nthreads)]
        # tid-major global order: give tid t a contiguous slice of the reads
        for t in range(nthreads):
            per_tid[t] = reads[len(reads) * t // nthreads: len(reads) * (t + 1) // nthreads]
            for bases, boff in synth.blocks(per_tid[t], max_bytes=block_bytes):
                g.ingest(bases, boff.astype(np.int32), tid=t)
        res = g.finish(fetch_table=True)
        off, prof = g.profiles()
        return res, off, prof
    finally:
        g.close()


@pytest.mark.parametrize("k,bc,nthreads", [(40, 0, 1), (21, 0, 3), (63, 0, 2), (40, 10, 2)])
def test_profiles_match_oracle(oracle_lib, k, bc, nthreads):
    genome = synth.random_genome(30_000, 41)
    reads = synth.sample_reads(genome, 1500, 200, 0.004, 42, n_rate=0.003, lower_rate=0.2, len_jitter=180)
    reads += [b"", b"ACGT", b"N" * 100, b"A" * 300, b"ACGTN" * 40]
    want = oracle_lib.count(reads, k, bc_prefix=bc, cutoff=1, profiles=True)
    res, off, prof = run_gpu_profiles(reads, k, bc=bc, nthreads=nthreads)
    assert np.array_equal(res.table, want["table"])
    assert len(off) == len(reads) + 1
    for r in range(len(reads)):
        assert np.array_equal(prof[off[r]:off[r + 1]], want["profiles"][r]), f"profile of read {r} differs"


@pytest.mark.parametrize("k,reserve_frac", [(40, 1.3), (63, 0.0), (21, 0.3)])
def test_profiles_with_interleaved_chunks(oracle_lib, monkeypatch, k, reserve_frac):
    """The profile output is tid-major while the device read stream holds the chunks in arrival order: three threads whose
    blocks arrive interleaved (so their 64 KB chunks alternate in the stream), ~150 tiles in 16 slices; streamed, plain and
    abandoned-stream front ends.  Long and short reads, so tiles hold many pieces and pieces span many tiles."""
    monkeypatch.setenv("FKGPU_CHUNK_BYTES", str(64 << 10))
    genome = synth.random_genome(120_000, 81)
    reads = synth.sample_reads(genome, 3000, 150, 0.003, 82, n_rate=0.002, len_jitter=100)
    reads += synth.sample_reads(genome, 40, 12_000, 0.002, 83) + [b"", b"ACGT", b"N" * 70]
    nthreads = 3
    total = sum(len(r) + 1 for r in reads)
    want = oracle_lib.count(reads, k, cutoff=1, profiles=True)
    g = FastKGPU(k=k, table_cutoff=1, profile=True, nthreads=nthreads, reserve_bases=int(total * reserve_frac))
    try:
        its = []
        for t in range(nthreads):                      # tid t owns a contiguous slice; blocks are handed over round-robin
            mine = reads[len(reads) * t // nthreads: len(reads) * (t + 1) // nthreads]
            its.append(iter(synth.blocks(mine, max_bytes=30_000)))
        live = list(range(nthreads))
        while live:
            for t in list(live):
                nxt = next(its[t], None)
                if nxt is None:
                    live.remove(t)
                else:
                    g.ingest(nxt[0], nxt[1].astype(np.int32), tid=t)
        got = g.finish(fetch_table=True)
        assert np.array_equal(got.table, want["table"])
        off, prof = g.profiles()
        assert len(off) == len(reads) + 1
        for r in range(len(reads)):
            assert np.array_equal(prof[off[r]:off[r + 1]], want["profiles"][r]), f"profile of read {r} differs"
    finally:
        g.close()


def test_split_long_read_rem_semantics(oracle_lib):
    """A read delivered in pieces with rem > 0 and a k-1 overlap (io.c:296-333) counts and profiles as one read."""
    k = 40
    genome = synth.random_genome(400_000, 51)
    reads = synth.sample_reads(genome, 6, 60_000, 0.001, 52)
    want = oracle_lib.count(reads, k, cutoff=1, profiles=True)
    g = FastKGPU(k=k, table_cutoff=1, profile=True)
    try:
        piece = 25_000
        for r in reads:
            pos = 0
            while True:
                end = min(len(r), pos + piece)
                last = end == len(r)
                bases = r[pos:end] + b"\0"
                g.ingest(bases, np.array([0, len(bases)], dtype=np.int32), tid=0, rem=0 if last else len(r) - end + (k - 1))
                if last:
                    break
                pos = end - (k - 1)
        res = g.finish(fetch_table=True)
        off, prof = g.profiles()
    finally:
        g.close()
    assert res.nkmers == want["nkmers"] and np.array_equal(res.table, want["table"])
    assert res.nreads == len(reads) and res.nbases == sum(len(r) for r in reads)
    assert len(off) == len(reads) + 1
    for r in range(len(reads)):
        assert np.array_equal(prof[off[r]:off[r + 1]], want["profiles"][r])


@pytest.mark.parametrize("k,cutoff,limit_mb", [(40, 1, 96), (21, 2, 64), (63, 1, 128), (40, 0, 64)])
def test_multi_round_count_leaves_disjoint_sorted_runs(oracle_lib, k, cutoff, limit_mb):
    """mem_limit (the host's -M) below the one-round working set: the reads are scanned once, minimizer-bucket ranges are
    counted round by round and every round leaves one sorted run (the reference's NPARTS > 1, count.c:1337).  The runs
    must be strictly increasing, pairwise disjoint, and merge to the oracle's table; histogram and scalars as usual."""
    genome = synth.random_genome(300_000, 81)
    reads = synth.sample_reads(genome, 40_000, 150, 0.004, 82, n_rate=0.001)
    want = oracle_lib.count(reads, k, cutoff=max(cutoff, 1))
    g = FastKGPU(k=k, table_cutoff=cutoff, nthreads=2, mem_limit=limit_mb << 20)
    try:
        for i, (bases, boff) in enumerate(synth.blocks(reads)):
            g.ingest(bases, boff.astype(np.int32), tid=i % 2)
        res = g.finish(fetch_table=True, copy_table=False)
        st = g.last_stats()
        assert st["path"] == 1
        assert res.nkmers == want["nkmers"] and res.ndistinct == want["ndistinct"] and res.max_inst == want["max_inst"]
        assert np.array_equal(res.hist[1:], want["hist"][1:])
        if cutoff == 0:
            assert res.ntable == 0
            return
        assert st["rounds"] == res.nruns and res.nruns > 1, st
        runs = res.view_runs()
        kb = res.kmer_bytes
        assert sum(len(r) for r in runs) == res.ntable == len(want["table"])
        for r in runs:
            keys = [bytes(x) for x in r[:, :kb]]
            assert keys == sorted(set(keys)), "a run is not strictly increasing"
        assert np.array_equal(res.merged_runs(), want["table"])
    finally:
        g.close()


def test_relative_profiles_against_a_loaded_table(oracle_lib):
    """-p:<table>: the table (oracle count of the golden c1_k40 reads at cutoff 2) is loaded, the query reads are only
    packed, and every profile must equal what the REFERENCE produced for the same table and reads (tests/golden/relative)."""
    import util
    g = util.golden_relative()
    tab = oracle_lib.count(util.read_seq_file(g["table_src"]), g["k"], cutoff=g["table_cutoff"])["table"]
    reads = util.read_seq_file(g["src"])
    eng = FastKGPU(k=g["k"], table_cutoff=0, profile=True, nthreads=2)
    try:
        eng.load_profile_table(tab)
        half = len(reads) // 2
        for t, part in enumerate((reads[:half], reads[half:])):
            for bases, boff in synth.blocks(part, max_bytes=20_000):
                eng.ingest(bases, boff.astype(np.int32), tid=t)
        res = eng.finish(fetch_table=False)
        assert res.ntable == 0 and res.nkmers == 0
        off, prof = eng.profiles()
        assert np.array_equal(off, g["prof_off"]) and np.array_equal(prof, g["prof"])
    finally:
        eng.close()


@pytest.mark.parametrize("env", [{"FKGPU_PROF": "legacy"}, {"FKGPU_PROF_LTC": "0"}])
def test_profile_path_switches(env):
    """The profile path has two process-wide switches (read once): FKGPU_PROF=legacy (sorted keys + prefix index + gather) and
    FKGPU_PROF_LTC=0 (lookups without the L2::64B hint).  Each runs the oracle comparisons of this file in a child process."""
    import os
    import subprocess
    import sys
    e = dict(os.environ)
    e.update(env)
    here = os.path.abspath(__file__)
    r = subprocess.run([sys.executable, "-m", "pytest", here, "-q", "-x", "-m", "gpu", "-p", "no:cacheprovider",
                        "-k", "test_profiles_match_oracle or test_relative_profiles_against_a_loaded_table"],
                       env=e, capture_output=True, text=True, cwd=os.path.dirname(os.path.dirname(here)), timeout=900)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    assert " passed" in r.stdout and "failed" not in r.stdout, r.stdout[-2000:]
