"""Multi-GPU k-mer counting: one process per GPU.

Super-mer path (k in 18..64, the default; `MultiGPUCounter._count_super`):
  1. every rank scans ITS reads into 8-byte super-mer records [minimizer bucket | # k-mers | GLOBAL position]
     partitioned by bucket                                                          (fkgpu_super_scan, CUDA)
  2. all-reduce(sum) of the 2^11-bin bucket histogram; contiguous bucket ranges per rank by the cumulative-threshold
     rule of the reference's thread split (MSDsort.c:330-352)                        (plumbing, torch.distributed)
  3. all-to-all of the 8-byte records and of their 32-byte left-aligned base strings (fkgpu_super_payload gathers them
     from the rank's own reads) over NVLink: ~3.7 B per k-mer instead of 16         (dist.all_to_all_single, NCCL)
  4. the owner expands + hash-counts its buckets on chip from the received base strings (fkgpu_super_count, CUDA).
     FKGPU_MG=peer instead exchanges ONLY the records and lets the counting kernel gather the bases straight out of the
     source rank's packed reads in peer HBM (CUDA IPC pointers): fine at 2 GPUs, but small random NVLink reads from
     three or more peers collapse (measured: 22 ms -> 490 ms for the kernel at 4 GPUs), hence not the default
     -> histogram / scalars complete per rank (a canonical k-mer lives in exactly one bucket): all-reduce(sum)
  5. only when a sorted table is wanted: the distinct (key | count) entries are partitioned by key prefix, exchanged
     with a second all-to-all and put in key order locally (fkgpu_entries_partition / fkgpu_entries_sort, CUDA);
     rank order == key order, so the global table is the rank-ordered concatenation.

Record path (any k; `MultiGPUCounter._count_records`): canonical k-mers sharded by contiguous key-prefix ranges.

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
import os
import time

import numpy as np
import torch
import torch.distributed as dist

PREFIX_BITS = 11
MAX_SUPER_RANKS = 8          # SUP_MAXRANKS of the bucket kernel: read streams one record can point into


def eng_wants_entries(eng):
    """distinct entries leave the bucket kernel only when a table (-t) or profiles (-p) are wanted"""
    return bool(getattr(eng, "table_cutoff", 0) > 0 or getattr(eng, "profile", False))


def splitters_from_hist(ghist, world):
    """ghist: 1-D int64 tensor/array of global per-prefix counts.  -> list beg[0..world] of prefix cut points:
    rank r owns prefixes [beg[r], beg[r+1]).  Same rule as msd_sort's panel split (MSDsort.c:330-352)."""
    h = np.asarray(ghist.cpu() if torch.is_tensor(ghist) else ghist, dtype=np.int64)
    cs = np.cumsum(h)
    total = int(cs[-1]) if len(cs) else 0
    beg, prev = [0], -1
    for n in range(1, world):                      # the n-th cut falls after the first bin whose running sum reaches n/world
        x = max(int(np.searchsorted(cs, (total * n) // world, side="left")), prev + 1)
        if x >= len(h):
            break
        beg.append(x + 1)
        prev = x
    while len(beg) < world:
        beg.append(len(h))
    beg.append(len(h))
    return beg


def exchange_plan(local_offsets, beg):
    """local_offsets: int64 [nprefix+1] starts of every prefix group in this rank's scattered buffer.
    -> send counts (records) to each rank."""
    lo = np.asarray(local_offsets.cpu() if torch.is_tensor(local_offsets) else local_offsets, dtype=np.int64)
    return [int(lo[beg[r + 1]] - lo[beg[r]]) for r in range(len(beg) - 1)]


def exchange_records(records, send_counts, group=None, recv_counts=None):
    """records: [n, w] int64 tensor ordered by destination rank.  One all-to-all of the variable-size slices.
    -> (received [m, w] tensor, recv_counts).  recv_counts may be passed when a previous exchange with the same plan
    already learnt them."""
    world = dist.get_world_size(group)
    dev = records.device
    if recv_counts is None:
        sc = torch.tensor(send_counts, dtype=torch.int64, device=dev)
        rc = torch.empty(world, dtype=torch.int64, device=dev)
        dist.all_to_all_single(rc, sc, group=group)
        recv_counts = [int(x) for x in rc.cpu()]
    out = torch.empty((sum(recv_counts) + 8, records.shape[1]), dtype=records.dtype, device=dev)   # +8 records of slack
    dist.all_to_all_single(out[:sum(recv_counts)], records[:sum(send_counts)], recv_counts, list(send_counts), group=group)
    return out, recv_counts


class MultiResult:
    pass


class _Timer:
    """host wall-clock laps of the staged pipeline (FKGPU_MG_TIMING=1, rank 0)"""

    def __init__(self, on):
        self.on, self.t, self.laps = bool(on), time.perf_counter(), []

    def lap(self, name):
        if self.on:
            torch.cuda.synchronize()
            t = time.perf_counter()
            self.laps.append((name, 1e3 * (t - self.t)))
            self.t = t

    def done(self):
        if self.on:
            print("[mg] " + ", ".join(f"{n} {v:.2f} ms" for n, v in self.laps), flush=True)


class _DevView:
    """Zero-copy view of library-owned device memory for torch (CUDA array interface)."""

    def __init__(self, ptr, n):
        self.__cuda_array_interface__ = {"shape": (int(n),), "typestr": "<i8", "data": (int(ptr), False), "version": 2}


def device_view(ptr, n, dev):
    """int64 tensor over n 8-byte words at device address ptr (no copy)."""
    if n <= 0 or not ptr:
        return torch.empty(0, dtype=torch.int64, device=dev)
    return torch.as_tensor(_DevView(ptr, n), device=dev)


class MultiGPUCounter:
    def __init__(self, eng, world, rank, dev):
        self.eng, self.world, self.rank, self.dev = eng, world, rank, dev
        self.w = eng.lib.fkgpu_record_bytes(eng.k) // 8
        self.nb = 1 << PREFIX_BITS
        self.hist = torch.zeros(self.nb, dtype=torch.int64, device=dev)
        self.offs = torch.zeros(self.nb + 1, dtype=torch.int64, device=dev)
        self.send = None
        self.ev = [torch.cuda.Event(enable_timing=True) for _ in range(2)]

    # ---- packed reads owned by the library so that peers can map them (CUDA IPC) ------------------------------
    def alloc_reads(self, npos):
        """-> (d_seq_ptr, d_val_ptr) for this rank's packed stream of npos positions.  Collective: every rank calls it;
        the buffers are exported to / imported from every peer and the global position bases are agreed on."""
        eng = self.eng
        self.close_peers()
        seq, val = eng.reads_alloc(npos)
        self.seq_ptr, self.val_ptr, self.npos = seq, val, npos
        sizes = torch.zeros(self.world, dtype=torch.int64, device=self.dev)
        dist.all_gather_into_tensor(sizes, torch.tensor([npos], dtype=torch.int64, device=self.dev))
        sizes = [int(x) for x in sizes.cpu()]
        self.pos_base = [0]
        for x in sizes:
            self.pos_base.append(self.pos_base[-1] + ((x + 63) // 64) * 64)
        self.peer_seq, self._opened = None, []
        self.super_ok = eng.super_supported()
        mode = os.environ.get("FKGPU_MG", "")
        if self.super_ok and self.world <= MAX_SUPER_RANKS and (mode == "peer" or (mode != "payload" and self.world <= 2)):
            # peer variant: every rank maps every other rank's packed reads (CUDA IPC).  Collective, so all ranks agree on
            # whether it worked; if any rank cannot export / map, everyone uses the payload exchange instead.
            ok = 1
            try:
                mine = torch.frombuffer(bytearray(eng.ipc_export(seq)), dtype=torch.uint8).to(self.dev)
            except Exception:
                ok, mine = 0, torch.zeros(64, dtype=torch.uint8, device=self.dev)
            allh = torch.zeros(self.world * 64, dtype=torch.uint8, device=self.dev)
            dist.all_gather_into_tensor(allh, mine)
            allh = allh.cpu().numpy().tobytes()
            flag = torch.tensor([ok], dtype=torch.int64, device=self.dev)
            dist.all_reduce(flag, op=dist.ReduceOp.MIN)
            peers = []
            if int(flag.item()):
                try:
                    for r in range(self.world):
                        peers.append(seq if r == self.rank else eng.ipc_open(allh[64 * r:64 * r + 64]))
                        if r != self.rank:
                            self._opened.append(peers[-1])
                except Exception:
                    ok = 0
            flag = torch.tensor([ok], dtype=torch.int64, device=self.dev)
            dist.all_reduce(flag, op=dist.ReduceOp.MIN)
            if int(flag.item()):
                self.peer_seq = peers
            else:
                self.close_peers()
        return seq, val

    def close_peers(self):
        for p in getattr(self, "_opened", []):
            self.eng.ipc_close(p)
        self._opened = []
        self.peer_seq = None

    def _all_agree(self, ok):
        """collective AND of a rank-local condition: every rank takes the same branch (a rank that went into an
        all-to-all alone would hang the job)"""
        flag = torch.tensor([1 if ok else 0], dtype=torch.int64, device=self.dev)
        dist.all_reduce(flag, op=dist.ReduceOp.MIN)
        return bool(int(flag.item()))

    def count_packed(self, d_seq, d_val, npos, fetch_table=False, copy_table=True):
        """d_seq / d_val: torch tensors or raw device pointers.  The super-mer path needs the buffers of alloc_reads().
        The choice of the pipeline, and the fall back to the record path when the super-mer scan of ANY rank declines
        its input, are agreed collectively."""
        sp = d_seq if isinstance(d_seq, int) else d_seq.data_ptr()
        vp = d_val if isinstance(d_val, int) else d_val.data_ptr()
        self.copy_table = copy_table
        mine = (getattr(self, "super_ok", False) and sp == getattr(self, "seq_ptr", None) and npos == getattr(self, "npos", -1)
                and os.environ.get("FKGPU_MG") != "records")
        if self._all_agree(mine):
            out = self._count_super(sp, vp, npos, fetch_table)
            if out is not None:
                return out
        return self._count_records(sp, vp, npos, fetch_table)

    def _finish(self, out, res, hist, scalars, t0, t1):
        h = torch.from_numpy(hist).to(self.dev)
        sc = torch.tensor(scalars, dtype=torch.int64, device=self.dev)
        dist.all_reduce(h)
        dist.all_reduce(sc)
        sizes = torch.zeros(self.world, dtype=torch.int64, device=self.dev)
        dist.all_gather_into_tensor(sizes, torch.tensor([scalars[3]], dtype=torch.int64, device=self.dev))
        t1.record()
        torch.cuda.synchronize()
        out.local = res
        out.hist = h.cpu().numpy()
        out.max_inst, out.nkmers, out.ndistinct, out.ntable = [int(x) for x in sc.cpu()]
        out.table_sizes = [int(x) for x in sizes.cpu()]
        out.table_offset = sum(out.table_sizes[:self.rank])
        out.kmer_bytes = res.kmer_bytes
        out.ms_total = t0.elapsed_time(t1)
        return out

    def _count_super(self, sp, vp, npos, fetch_table):
        eng, dev = self.eng, self.dev
        t0 = torch.cuda.Event(enable_timing=True)
        t1 = torch.cuda.Event(enable_timing=True)
        t0.record()
        total = self.pos_base[-1]
        tm = _Timer(self.rank == 0 and os.environ.get("FKGPU_MG_TIMING"))
        try:
            sc = eng.super_scan(sp, vp, npos, total, self.pos_base[self.rank])
        except Exception as e:                       # e.g. more super-mers than the staging buffer holds on this rank
            sc, self.last_super_error = None, str(e)
        if not self._all_agree(sc is not None):
            return None                              # every rank falls back to the record path together
        scan_ms = dict(eng.stage_times())
        tm.lap("super_scan")
        nb = 1 << sc["bits"]
        ghist = device_view(sc["hist"], nb, dev).clone()
        dist.all_reduce(ghist)
        beg = splitters_from_hist(ghist, self.world)
        send_counts = exchange_plan(device_view(sc["offsets"], nb + 1, dev), beg)
        recs = device_view(sc["records"], sc["n"], dev).view(-1, 1)
        tm.lap("hist all-reduce + splitters")
        mode = os.environ.get("FKGPU_MG", "")
        peer_mode = self.peer_seq is not None and ((mode == "peer") or (mode != "payload" and self.world <= 2))
        payload = None
        if not peer_mode:
            # the base string of every super-mer (32 bytes, left aligned), in record order: travels beside the records
            payload = torch.empty((sc["n"] + 8, 4), dtype=torch.int64, device=dev)
            eng.super_payload(sp, self.pos_base[self.rank], total, sc["records"], sc["n"], payload.data_ptr())
            tm.lap("payload gather")
        recv, recv_counts = exchange_records(recs, send_counts)
        torch.cuda.current_stream().synchronize()      # NCCL wrote `recv` on torch's stream; the library runs on its own
        tm.lap("record all-to-all")
        nrecv = sum(recv_counts)
        recv_pl, ready = None, None
        if payload is not None:
            # asynchronous: the payload crosses NVLink while the library partitions the records; the counting kernel is
            # ordered behind `ready` on the device, the host never waits for it
            recv_pl = torch.empty((nrecv + 8, 4), dtype=torch.int64, device=dev)
            work = dist.all_to_all_single(recv_pl[:nrecv], payload[:sum(send_counts)], recv_counts, list(send_counts),
                                          async_op=True)
            work.wait()
            ready = torch.cuda.Event()
            ready.record()
        want_entries = eng_wants_entries(eng)
        res, ent_ptr, nent = eng.super_count(recv.data_ptr(), nrecv, total, self.peer_seq if peer_mode else [sp],
                                             self.pos_base if peer_mode else [0, total], want_entries,
                                             d_payload_ptr=None if recv_pl is None else recv_pl.data_ptr(),
                                             ready_event=None if ready is None else ready.cuda_event)
        del recv_pl, payload
        tm.lap("super_count")
        out = MultiResult()
        out.path = "super-mer"
        out.exchange = "peer-gather" if peer_mode else "payload"
        out.sent_records = sc["n"] - send_counts[self.rank]
        out.supermers, out.entries = sc["n"], nent
        out.owned_buckets = (beg[self.rank], beg[self.rank + 1])
        out.stage_ms = dict(eng.stage_times())
        for kname, v in scan_ms.items():
            if v:
                out.stage_ms[kname] = out.stage_ms.get(kname, 0.0) + v
        hist, ntable = res.hist, 0
        if want_entries:
            # second exchange: distinct (key | count) entries by key prefix -> rank order == key order
            del recv
            nb2 = 1 << PREFIX_BITS
            e_hist = torch.zeros(nb2, dtype=torch.int64, device=dev)
            e_offs = torch.zeros(nb2 + 1, dtype=torch.int64, device=dev)
            part = torch.empty((nent + 8, eng.lib.fkgpu_entry_bytes(eng.k) // 8), dtype=torch.int64, device=dev)
            eng.entries_partition(ent_ptr, nent, PREFIX_BITS, part.data_ptr(), e_hist.data_ptr(), e_offs.data_ptr())
            g2 = e_hist.clone()
            dist.all_reduce(g2)
            beg2 = splitters_from_hist(g2, self.world)
            sc2 = exchange_plan(e_offs, beg2)
            tm.lap("entries partition + all-reduce")
            recv2, rc2 = exchange_records(part, sc2)
            torch.cuda.current_stream().synchronize()
            tm.lap("entries all-to-all")
            tres = eng.entries_sort(recv2.data_ptr(), sum(rc2), fetch_table=fetch_table, copy_table=self.copy_table)
            tm.lap("entries_sort")
            for kname, v in eng.stage_times().items():
                if v:
                    out.stage_ms[kname] = out.stage_ms.get(kname, 0.0) + v
            out.owned_prefixes = (beg2[self.rank], beg2[self.rank + 1])
            out.sent_entries = nent - sc2[self.rank]
            ntable = tres.ntable
            tres.hist = hist
            local = tres
        else:
            local = res
        local.max_inst, local.ndistinct = res.max_inst, res.ndistinct
        tm.done()
        return self._finish(out, local, hist, [res.max_inst, sc["nkmers"], res.ndistinct, ntable], t0, t1)

    def _count_records(self, sp, vp, npos, fetch_table=False):
        eng = self.eng
        t0 = torch.cuda.Event(enable_timing=True)
        t1 = torch.cuda.Event(enable_timing=True)
        t0.record()
        eng.prefix_hist(sp, vp, npos, PREFIX_BITS, self.hist.data_ptr())
        ghist = self.hist.clone()
        dist.all_reduce(ghist)
        beg = splitters_from_hist(ghist, self.world)
        nloc = int(self.hist.sum().item())
        if self.send is None or self.send.shape[0] < nloc + 4:
            self.send = None
            self.send = torch.empty((nloc + nloc // 16 + 4, self.w), dtype=torch.int64, device=self.dev)
        eng.scatter_prefix(sp, vp, npos, PREFIX_BITS, self.hist.data_ptr(),
                           self.send.data_ptr(), self.send.shape[0], self.offs.data_ptr())
        send_counts = exchange_plan(self.offs, beg)
        recv, recv_counts = exchange_records(self.send, send_counts)
        torch.cuda.current_stream().synchronize()      # NCCL wrote `recv` on torch's stream; the library runs on its own
        nrecv = sum(recv_counts)
        res = eng.count_records(recv.data_ptr(), nrecv, fetch_table=fetch_table, copy_table=getattr(self, 'copy_table', True))
        out = MultiResult()
        out.path = "records"
        out.stage_ms = dict(eng.stage_times())
        out.sent_records = nloc - send_counts[self.rank]
        out.owned_prefixes = (beg[self.rank], beg[self.rank + 1])
        return self._finish(out, res, res.hist, [res.max_inst, res.nkmers, res.ndistinct, res.ntable], t0, t1)
