"""Multi-GPU k-mer counting: one process per GPU, canonical k-mers sharded by contiguous key-prefix ranges.

  1. every rank scans ITS reads: histogram of the top PREFIX_BITS key bits        (fkgpu_prefix_hist, CUDA)
  2. all-reduce(sum) of the histogram over NCCL; every rank derives the same contiguous splitters with the
     cumulative-threshold rule of the reference's thread split (MSDsort.c:330-352)  (plumbing, torch.distributed)
  3. every rank scatters its canonical k-mers into prefix order                   (fkgpu_scatter_prefix, CUDA)
     -> the records owned by rank r are ONE contiguous slice of the local buffer
  4. one all-to-all of those slices over NVLink                                   (dist.all_to_all_single, NCCL)
  5. purely local sort / count / histogram / table of the received records        (fkgpu_count_records, CUDA)
  6. all-reduce of the 32768-bin histogram and the scalars; rank order == key order, so the global table is the
     rank-ordered concatenation of the per-rank tables (all-gather of the entry counts gives the offsets).

A canonical k-mer has exactly one owner, so no count is ever merged across ranks.  The pure-torch helpers
(`splitters_from_hist`, `exchange_plan`, `exchange_records`) carry the N>1 logic and are exercised on CPU with the
gloo backend in tests/test_multigpu_gloo.py.
"""
import numpy as np
import torch
import torch.distributed as dist

PREFIX_BITS = 11


def splitters_from_hist(ghist, world):
    """ghist: 1-D int64 tensor/array of global per-prefix counts.  -> list beg[0..world] of prefix cut points:
    rank r owns prefixes [beg[r], beg[r+1]).  Same rule as msd_sort's panel split (MSDsort.c:330-352)."""
    h = np.asarray(ghist.cpu() if torch.is_tensor(ghist) else ghist, dtype=np.int64)
    total = int(h.sum())
    beg = [0]
    n, s = 0, 0
    thr = total // world
    for x in range(len(h)):
        s += int(h[x])
        if s >= thr and n < world - 1:
            n += 1
            beg.append(x + 1)
            thr = (total * (n + 1)) // world
    while len(beg) < world:
        beg.append(len(h))
    beg.append(len(h))
    return beg


def exchange_plan(local_offsets, beg):
    """local_offsets: int64 [nprefix+1] starts of every prefix group in this rank's scattered buffer.
    -> send counts (records) to each rank."""
    lo = np.asarray(local_offsets.cpu() if torch.is_tensor(local_offsets) else local_offsets, dtype=np.int64)
    return [int(lo[beg[r + 1]] - lo[beg[r]]) for r in range(len(beg) - 1)]


def exchange_records(records, send_counts, group=None):
    """records: [n, w] int64 tensor ordered by destination rank.  One all-to-all of the variable-size slices.
    -> (received [m, w] tensor, recv_counts)."""
    world = dist.get_world_size(group)
    dev = records.device
    sc = torch.tensor(send_counts, dtype=torch.int64, device=dev)
    rc = torch.empty(world, dtype=torch.int64, device=dev)
    dist.all_to_all_single(rc, sc, group=group)
    recv_counts = [int(x) for x in rc.cpu()]
    out = torch.empty((sum(recv_counts) + 4, records.shape[1]), dtype=records.dtype, device=dev)   # +4 records of slack
    dist.all_to_all_single(out[:sum(recv_counts)], records[:sum(send_counts)], recv_counts, list(send_counts), group=group)
    return out, recv_counts


class MultiResult:
    pass


class MultiGPUCounter:
    def __init__(self, eng, world, rank, dev):
        self.eng, self.world, self.rank, self.dev = eng, world, rank, dev
        self.w = eng.lib.fkgpu_record_bytes(eng.k) // 8
        self.nb = 1 << PREFIX_BITS
        self.hist = torch.zeros(self.nb, dtype=torch.int64, device=dev)
        self.offs = torch.zeros(self.nb + 1, dtype=torch.int64, device=dev)
        self.send = None
        self.ev = [torch.cuda.Event(enable_timing=True) for _ in range(2)]

    def count_packed(self, d_seq, d_val, npos, fetch_table=False):
        eng = self.eng
        t0 = torch.cuda.Event(enable_timing=True)
        t1 = torch.cuda.Event(enable_timing=True)
        t0.record()
        eng.prefix_hist(d_seq.data_ptr(), d_val.data_ptr(), npos, PREFIX_BITS, self.hist.data_ptr())
        ghist = self.hist.clone()
        dist.all_reduce(ghist)
        beg = splitters_from_hist(ghist, self.world)
        nloc = int(self.hist.sum().item())
        if self.send is None or self.send.shape[0] < nloc + 4:
            self.send = None
            self.send = torch.empty((nloc + nloc // 16 + 4, self.w), dtype=torch.int64, device=self.dev)
        eng.scatter_prefix(d_seq.data_ptr(), d_val.data_ptr(), npos, PREFIX_BITS, self.hist.data_ptr(),
                           self.send.data_ptr(), self.send.shape[0], self.offs.data_ptr())
        send_counts = exchange_plan(self.offs, beg)
        recv, recv_counts = exchange_records(self.send, send_counts)
        torch.cuda.current_stream().synchronize()      # NCCL wrote `recv` on torch's stream; the library runs on its own
        nrecv = sum(recv_counts)
        res = eng.count_records(recv.data_ptr(), nrecv, fetch_table=fetch_table)
        # global reductions: histogram + scalars (sum), table sizes (gather)
        h = torch.from_numpy(res.hist).to(self.dev)
        sc = torch.tensor([res.max_inst, res.nkmers, res.ndistinct, res.ntable], dtype=torch.int64, device=self.dev)
        dist.all_reduce(h)
        dist.all_reduce(sc)
        sizes = torch.zeros(self.world, dtype=torch.int64, device=self.dev)
        mine = torch.tensor([res.ntable], dtype=torch.int64, device=self.dev)
        dist.all_gather_into_tensor(sizes, mine)
        t1.record()
        torch.cuda.synchronize()
        out = MultiResult()
        out.local = res
        out.hist = h.cpu().numpy()
        out.max_inst, out.nkmers, out.ndistinct, out.ntable = [int(x) for x in sc.cpu()]
        out.table_sizes = [int(x) for x in sizes.cpu()]
        out.table_offset = sum(out.table_sizes[:self.rank])
        out.kmer_bytes = res.kmer_bytes
        out.ms_total = t0.elapsed_time(t1)
        out.sent_records = nloc - send_counts[self.rank]
        out.owned_prefixes = (beg[self.rank], beg[self.rank + 1])
        return out
