"""Seeded synthetic read sets (numpy, host side) shaped like the BASELINE.json configs.

Small / medium sets for the parity tests and the CPU baseline; bench.py builds the large device-resident
sets with the same model (uniform random genome, reads sampled at uniform positions, random strand,
independent substitution errors)."""
import numpy as np

_ACGT = np.frombuffer(b"ACGT", dtype=np.uint8)
_LOW = np.frombuffer(b"acgt", dtype=np.uint8)


def random_genome(size, seed):
    return np.random.default_rng(seed).integers(0, 4, size, dtype=np.uint8)


def sample_reads(genome, nreads, read_len, sub_rate, seed, n_rate=0.0, lower_rate=0.0, len_jitter=0):
    """-> list of bytes objects (ASCII reads).  n_rate: per-base chance of an 'N'; lower_rate: per-read
    chance of being written in lower case; len_jitter: +- uniform length jitter."""
    rng = np.random.default_rng(seed)
    G = len(genome)
    out = []
    for _ in range(nreads):
        L = read_len + (int(rng.integers(-len_jitter, len_jitter + 1)) if len_jitter else 0)
        L = max(1, min(L, G))
        s = int(rng.integers(0, G - L + 1))
        r = genome[s:s + L].copy()
        if sub_rate > 0:
            m = rng.random(L) < sub_rate
            r[m] = (r[m] + rng.integers(1, 4, int(m.sum()))) % 4
        if rng.random() < 0.5:
            r = (3 - r)[::-1]
        a = (_LOW if rng.random() < lower_rate else _ACGT)[r]
        if n_rate > 0:
            a = a.copy()
            a[rng.random(L) < n_rate] = ord("N")
        out.append(a.tobytes())
    return out


def to_block(reads):
    """list of reads -> (bases, boff) laid out like a DATA_BLOCK (FastK.h:87-98): 0-terminated reads,
    boff[i] = start of read i, boff[n] = total bytes."""
    boff = np.zeros(len(reads) + 1, dtype=np.int64)
    for i, r in enumerate(reads):
        boff[i + 1] = boff[i] + len(r) + 1
    bases = b"".join(r + b"\0" for r in reads)
    return bases, boff


def blocks(reads, max_bytes=1_000_000, max_reads=10_000):
    """Cut a read list into DATA_BLOCK sized pieces (io.c:64-66: 1 MB / 10 000 reads), whole reads only."""
    cur, size = [], 0
    for r in reads:
        if cur and (size + len(r) + 1 > max_bytes or len(cur) >= max_reads):
            yield to_block(cur)
            cur, size = [], 0
        cur.append(r)
        size += len(r) + 1
    if cur:
        yield to_block(cur)


def write_fasta(reads, path, width=0):
    with open(path, "wb") as f:
        for i, r in enumerate(reads):
            f.write(b">r%d\n" % i)
            if width:
                for j in range(0, len(r), width):
                    f.write(r[j:j + width] + b"\n")
                if len(r) == 0:
                    f.write(b"\n")
            else:
                f.write(r + b"\n")


def write_fastq(reads, path):
    with open(path, "wb") as f:
        for i, r in enumerate(reads):
            f.write(b"@r%d\n" % i + r + b"\n+\n" + b"I" * len(r) + b"\n")


# ---- bench-scale workloads: one generator for BOTH bench arms and the at-scale parity tests ------------------------------

def workload_rows(genome_bp, nreads, read_len, sub_rate, seed, out=None):
    """Fixed-length reads of the BASELINE.json shape -> uint8 [nreads, read_len+1], every row one ASCII read followed by
    a 0 terminator: the concatenation of the rows is a DATA_BLOCK's `bases` (FastK.h:92).  Deterministic in its
    arguments, so the GPU arm, the reference arm and the parity check of bench.py all see the same reads."""
    rng = np.random.default_rng(seed)
    G, L = int(genome_bp), int(read_len)
    genome = rng.integers(0, 4, G, dtype=np.uint8)
    rows = out if out is not None else np.empty((nreads, L + 1), dtype=np.uint8)
    assert rows.shape == (nreads, L + 1) and rows.dtype == np.uint8
    rows[:, L] = 0
    ar = np.arange(L, dtype=np.int64)
    chunk = max(1, (32 << 20) // L)
    for r0 in range(0, nreads, chunk):
        r1 = min(nreads, r0 + chunk)
        n = r1 - r0
        st = rng.integers(0, G - L + 1, n)
        if L >= 1024:
            blk = np.empty((n, L), dtype=np.uint8)
            for i in range(n):
                blk[i] = genome[st[i]:st[i] + L]
        else:
            blk = genome[st[:, None] + ar[None, :]]
        if sub_rate > 0:
            ns = int(rng.binomial(n * L, sub_rate))
            pos = rng.integers(0, n * L, ns)
            add = rng.integers(1, 4, ns, dtype=np.uint8)
            flat = blk.reshape(-1)
            flat[pos] = (flat[pos] + add) & 3
        flip = rng.random(n) < 0.5
        blk[flip] = (3 - blk[flip])[:, ::-1]
        rows[r0:r1, :L] = _ACGT[blk]
    return rows


def write_rows_fasta(rows, path, first_id=0, append=False):
    """rows of workload_rows -> single-line FASTA with fixed-width headers (one 2-D array, one write)."""
    n, L = rows.shape[0], rows.shape[1] - 1
    hdr = 12                                               # ">r%010d"
    rec = np.empty((n, hdr + 1 + L + 1), dtype=np.uint8)
    ids = np.arange(first_id, first_id + n, dtype=np.int64)
    rec[:, 0] = ord(">")
    rec[:, 1] = ord("r")
    for d in range(10):
        rec[:, 11 - d] = ord("0") + (ids // 10 ** d) % 10
    rec[:, hdr] = ord("\n")
    rec[:, hdr + 1:hdr + 1 + L] = rows[:, :L]
    rec[:, hdr + 1 + L] = ord("\n")
    with open(path, "ab" if append else "wb") as f:
        rec.tofile(f)
    return n * L
