"""CPU: the input side of the C host program (fastk_b200/host/fastk_main.c -- the role of io.c) without a GPU.
The program is linked against tests/hoststub/fkgpu_stub.c, a recording stand-in for the library that only stores the
DATA_BLOCKs it is handed; the reads it received must be exactly the reads of the files, in file order, whatever the
format, line width, compression, thread count or read length (long reads arrive in pieces with the k-1 overlap)."""
import gzip
import os
import subprocess

import numpy as np
import pytest

import util
from fastk_b200 import synth

ROOT = util.ROOT
HOST = os.path.join(ROOT, "fastk_b200", "host")


@pytest.fixture(scope="module")
def stub_cli(tmp_path_factory):
    d = str(tmp_path_factory.mktemp("hoststub"))
    exe = os.path.join(d, "FastK_stub")
    subprocess.check_call(["gcc", "-O2", "-Wall", "-I" + os.path.join(ROOT, "include"), "-I" + HOST, "-o", exe,
                           os.path.join(HOST, "fastk_main.c"), os.path.join(HOST, "fk_files.c"),
                           os.path.join(ROOT, "tests", "hoststub", "fkgpu_stub.c"), "-lpthread", "-lz", "-lm"])
    return exe


def run(exe, files, d, args=()):
    out = os.path.join(d, "delivered.txt")
    r = subprocess.run([exe, "-N" + os.path.join(d, "o")] + list(args) + files, capture_output=True, text=True,
                       env=dict(os.environ, FKSTUB_OUT=out))
    assert r.returncode == 0, r.stderr
    reads, meta = [], []
    for line in open(out, "rb").read().split(b"\n")[:-1]:
        if line.startswith(b"#tid"):
            meta.append(dict(zip(line.split()[2::2], [int(x) for x in line.split()[3::2]])))
        else:
            reads.append(line)
    return reads, meta


def hoco(r):
    a = np.frombuffer(r, dtype=np.uint8)
    return r if len(a) == 0 else a[np.concatenate(([True], a[1:] != a[:-1]))].tobytes()


def make_reads(n, length, seed, jitter=0, n_rate=0.002):
    return synth.sample_reads(synth.random_genome(max(3 * length, 50_000), seed), n, length, 0.01, seed + 1,
                              n_rate=n_rate, lower_rate=0.1, len_jitter=jitter)


def write_fastq(reads, path):
    with open(path, "wb") as f:
        for i, r in enumerate(reads):
            f.write(b"@r%d some text\n" % i + r + b"\n+\n" + b"I" * len(r) + b"\n")


@pytest.mark.parametrize("fmt,width,threads", [("fa", 0, 1), ("fa", 70, 4), ("fa", 0, 7), ("fq", 0, 1), ("fq", 0, 4)])
def test_reads_delivered_in_file_order(stub_cli, tmp_path, fmt, width, threads):
    reads = make_reads(12_000, 150, 5, jitter=100) + [b"", b"A", b"ACGTN" * 30]
    src = os.path.join(str(tmp_path), "in." + ("fasta" if fmt == "fa" else "fastq"))
    if fmt == "fa":
        synth.write_fasta(reads, src, width=width)
        if width == 0:
            reads = [r for r in reads]
    else:
        write_fastq(reads, src)
    got, meta = run(stub_cli, [src], str(tmp_path), ["-k40", "-T%d" % threads])
    assert got == reads
    assert len(meta) == threads and all(m[b"maxblock"] <= 1_000_000 and m[b"maxreads"] <= 10_000 for m in meta)
    if threads > 1:
        assert sum(1 for m in meta if m[b"blocks"] > 0) == threads, "every reader thread should own a byte range"


def test_gzip_compress_and_several_files(stub_cli, tmp_path):
    a, b = make_reads(3000, 200, 11, jitter=150), make_reads(2000, 120, 13)
    fa, fb = os.path.join(str(tmp_path), "a.fa"), os.path.join(str(tmp_path), "b.fa")
    synth.write_fasta(a, fa, width=60)
    synth.write_fasta(b, fb)
    got, _ = run(stub_cli, [fa, fb], str(tmp_path), ["-k21", "-T3"])
    assert got == a + b
    got, _ = run(stub_cli, [fa, fb], str(tmp_path), ["-k21", "-T3", "-c"])
    assert got == [hoco(r) for r in a + b]                       # io.c:284-294
    for p in (fa, fb):
        with open(p, "rb") as f, gzip.open(p + ".gz", "wb") as g:
            g.write(f.read())
        os.remove(p)
    got, meta = run(stub_cli, [fa + ".gz", fb + ".gz"], str(tmp_path), ["-k21", "-T4"])
    assert got == a + b and len(meta) == 2                       # one reader thread per .gz file (io.c:2373-2378)


def test_long_reads_arrive_in_overlapping_pieces(stub_cli, tmp_path):
    reads = make_reads(5, 2_600_000, 17, n_rate=0.0) + make_reads(40, 15_000, 19) + make_reads(1, 1_000_100, 23, n_rate=0.0)
    src = os.path.join(str(tmp_path), "long.fa")
    synth.write_fasta(reads, src, width=80)
    for k in (40, 21):
        got, meta = run(stub_cli, [src], str(tmp_path), ["-k%d" % k, "-T2"])
        assert [len(r) for r in got] == [len(r) for r in reads]
        assert got == reads                                      # the stub stitched the pieces on their k-1 overlap
        assert all(m[b"maxblock"] <= 1_000_000 for m in meta)


REF = "/root/reference"


@pytest.fixture(scope="module")
def refhost_stub(tmp_path_factory):
    """The REFERENCE's own driver + io.c + table.c + libfastk.c (compiled where they lie) with fastk_shim.c, linked
    against the recording stand-in: what the reference's input module hands to Distribute_Block, without a GPU."""
    if not os.path.exists(os.path.join(REF, "FastK.c")):
        pytest.skip("needs the reference sources")
    d = str(tmp_path_factory.mktemp("refstub"))
    exe = os.path.join(d, "FastK_refhost_stub")
    srcs = [os.path.join(REF, f) for f in ("FastK.c", "io.c", "table.c", "libfastk.c")]
    subprocess.check_call(["gcc", "-O2", "-w", "-I" + REF, "-I" + os.path.join(REF, "HTSLIB"), "-I" + os.path.join(ROOT, "include"),
                           "-I" + HOST, "-o", exe] + srcs +
                          [os.path.join(HOST, "fastk_shim.c"), os.path.join(HOST, "fk_files.c"),
                           os.path.join(ROOT, "oracle", "ref_thirdparty_stubs.c"),
                           os.path.join(ROOT, "tests", "hoststub", "fkgpu_stub.c"), "-lpthread", "-lz", "-lm"])
    return exe


def test_reference_io_module_through_the_shim(refhost_stub, tmp_path):
    """io.c's own blocks (ITHREADS producers, long reads cut with rem > 0 and the k-1 overlap, io.c:296-333) forwarded
    by fastk_shim.c's Distribute_Block: the stand-in must receive exactly the reads of the file."""
    reads = make_reads(3, 2_300_000, 31, n_rate=0.0) + make_reads(6000, 400, 37, jitter=300)
    src = os.path.join(str(tmp_path), "mix.fasta")
    synth.write_fasta(reads, src, width=100)
    out = os.path.join(str(tmp_path), "delivered.txt")
    r = subprocess.run([refhost_stub, "-k40", "-T4", "-P" + str(tmp_path), "-N" + os.path.join(str(tmp_path), "o"), src],
                       capture_output=True, text=True, env=dict(os.environ, FKSTUB_OUT=out))
    assert r.returncode == 0, r.stderr
    got = [ln for ln in open(out, "rb").read().split(b"\n")[:-1] if not ln.startswith(b"#tid")]
    assert sorted(len(x) for x in got) == sorted(len(x) for x in reads)
    assert got == reads


def test_truncated_gz_is_an_error_not_a_short_input(stub_cli, tmp_path):
    """A corrupt / truncated .gz must end the run with an error (and no outputs), never count the part that could be read."""
    d = str(tmp_path)
    reads = make_reads(3000, 150, 91)
    fa = os.path.join(d, "t.fa")
    synth.write_fasta(reads, fa)
    raw = gzip.compress(open(fa, "rb").read())
    gz = os.path.join(d, "cut.fa.gz")
    open(gz, "wb").write(raw[:len(raw) * 2 // 3])
    r = subprocess.run([stub_cli, "-N" + os.path.join(d, "o"), "-t1", gz], capture_output=True, text=True,
                       env=dict(os.environ, FKSTUB_OUT=os.path.join(d, "delivered.txt")))
    assert r.returncode == 1 and "Error reading" in r.stderr, (r.returncode, r.stderr)
    assert not os.path.exists(os.path.join(d, "o.hist")) and not os.path.exists(os.path.join(d, "o.ktab"))
