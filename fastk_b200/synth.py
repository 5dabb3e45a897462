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
