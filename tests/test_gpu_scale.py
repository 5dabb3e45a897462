"""GPU parity AT THE MEASURED SCALE: >= 1 Gbase batches through the C ABI (fkgpu_ingest / fkgpu_finish) against the files
the reference FastK (oracle/_ref, built from the reference's own sources) writes for the same FASTA -- every .hist bin,
max_inst, the .ktab prefix index and every suffix + count byte (fastk_b200/formats.compare_with_fastk_files).
Covers the geometry the bench runs in (22-bit bucket ids, 32+ position bits, ~10^5..10^6 work groups) and a batch of
more than 2^32 positions; the small oracle-checked cases of test_gpu_parity.py never reach either."""
import os
import shutil
import subprocess
import tempfile
import threading

import numpy as np
import pytest

from fastk_b200 import FastKGPU, formats, synth


pytestmark = pytest.mark.gpu


def count_rows_e2e(rows, k, cutoff, nthr=4):
    nreads, L1 = rows.shape
    npos = nreads * L1
    g = FastKGPU(k=k, table_cutoff=cutoff, nthreads=nthr, reserve_bases=npos)
    per = max(1, min(10000, (1_000_000 - 1) // L1))
    boff = (np.arange(per + 1, dtype=np.int64) * L1).astype(np.int32)
    blocks = [(r0, min(nreads, r0 + per)) for r0 in range(0, nreads, per)]
    base = rows.ctypes.data
    errs = []

    def worker(tid):
        try:
            for bi in range(len(blocks) * tid // nthr, len(blocks) * (tid + 1) // nthr):
                r0, r1 = blocks[bi]
                g.ingest_ptr(base + r0 * L1, boff.ctypes.data, r1 - r0, tid=tid)
        except Exception as e:                      # noqa: BLE001
            errs.append(e)

    th = [threading.Thread(target=worker, args=(t,)) for t in range(nthr)]
    for t in th:
        t.start()
    for t in th:
        t.join()
    assert not errs, errs
    res = g.finish(fetch_table=True, copy_table=False)
    return g, res


@pytest.mark.parametrize("genome_mbp,read_len,cov,k,cutoff", [
    (22.0, 15000, 50.0, 40, 1),          # 1.1 Gbases: bucket ids 22 bits wide, as in the bench
    (30.0, 150, 40.0, 21, 4),            # 1.2 Gbases of short reads, k=21 -t4 (config 3's shape)
    (88.0, 15000, 50.0, 40, 2),          # 4.4 Gbases: more than 2^32 positions in one batch
])
def test_gbase_batch_equals_reference_fastk(ref_bin, genome_mbp, read_len, cov, k, cutoff):
    if ref_bin is None:
        pytest.skip("oracle/_ref/FastK not built")
    nreads = int(genome_mbp * 1e6 * cov / read_len)
    rows = synth.workload_rows(int(genome_mbp * 1e6), nreads, read_len, 0.001, 4242)
    d = tempfile.mkdtemp(prefix="fastk_scale_", dir="/dev/shm" if os.path.isdir("/dev/shm") else None)
    g = None
    try:
        fasta = os.path.join(d, "reads.fasta")
        synth.write_rows_fasta(rows, fasta)
        subprocess.check_call([os.path.join(ref_bin, "FastK"), f"-k{k}", f"-t{cutoff}", f"-T{min(32, os.cpu_count() or 1)}", "-M16",
                               "-P" + d, "-N" + os.path.join(d, "ref"), fasta], stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
        os.remove(fasta)
        g, res = count_rows_e2e(rows, k, cutoff)
        assert res.nbases == nreads * read_len and res.nreads == nreads
        assert g.last_path() == 1, "expected the super-mer pipeline"
        # beyond one device round the table arrives as sorted runs (NPARTS semantics): merge them as Merge_Tables would
        table = res.merged_runs() if res.nruns > 1 else res.view_table()
        bad = formats.compare_with_fastk_files(d, "ref", k, cutoff, res.hist, res.max_inst, table)
        assert bad == [], bad
    finally:
        if g is not None:
            g.close()
        shutil.rmtree(d, ignore_errors=True)
