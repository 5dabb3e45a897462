"""ctypes binding of include/fastk_gpu.h.  Fails loudly when the CUDA library is missing."""
import ctypes as C
import os
import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("FKGPU_LIB") or os.path.join(HERE, "lib", "libfastk_gpu.so")

HIST_BINS = 32768
NSTAGES = 14
STAGES = ["pack", "scan_hist", "scan_scatter", "l1_hist_records", "refine", "sortcount", "compact", "profile",
          "super_scan", "super_partition", "bucket_count", "entry_partition", "super_refine", "spill"]
PACK_PAD = 16


class FkgpuError(RuntimeError):
    pass


class _Config(C.Structure):
    _fields_ = [("kmer", C.c_int32), ("do_table", C.c_int32), ("do_profile", C.c_int32),
                ("bc_prefix", C.c_int32), ("device", C.c_int32), ("nthreads", C.c_int32),
                ("reserve_bases", C.c_int64), ("mem_limit", C.c_int64)]


class _Result(C.Structure):
    _fields_ = [("kmer", C.c_int32), ("kmer_bytes", C.c_int32),
                ("nbases", C.c_int64), ("nreads", C.c_int64), ("nkmers", C.c_int64), ("ndistinct", C.c_int64),
                ("hist", C.POINTER(C.c_int64)), ("max_inst", C.c_int64),
                ("ntable", C.c_int64), ("table", C.POINTER(C.c_uint8)), ("table_dev", C.c_void_p),
                ("ms_pack", C.c_float), ("ms_count", C.c_float), ("ms_total", C.c_float),
                ("nruns", C.c_int32), ("run_ntable", C.POINTER(C.c_int64)), ("run_table", C.POINTER(C.POINTER(C.c_uint8)))]


_lib = None


def load_library(path=None):
    """dlopen libfastk_gpu.so and declare every prototype of include/fastk_gpu.h."""
    global _lib
    if _lib is not None and path is None:
        return _lib
    p = path or LIB_PATH
    if not os.path.exists(p):
        raise FkgpuError(f"{p} is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
                         "(there is no CPU fallback)")
    lib = C.CDLL(p)
    vp, i32, i64 = C.c_void_p, C.c_int32, C.c_int64
    lib.fkgpu_create.argtypes = [C.POINTER(_Config), C.POINTER(vp)]
    lib.fkgpu_create.restype = C.c_int
    lib.fkgpu_destroy.argtypes = [vp]
    lib.fkgpu_destroy.restype = None
    lib.fkgpu_reset.argtypes = [vp]
    lib.fkgpu_reset.restype = C.c_int
    lib.fkgpu_last_error.argtypes = []
    lib.fkgpu_last_error.restype = C.c_char_p
    lib.fkgpu_device_count.argtypes = []
    lib.fkgpu_device_count.restype = C.c_int
    lib.fkgpu_ingest.argtypes = [vp, C.c_int, C.c_char_p, C.POINTER(i32), i32, i32]
    lib.fkgpu_ingest.restype = C.c_int
    lib.fkgpu_finish.argtypes = [vp, C.c_int, C.POINTER(_Result)]
    lib.fkgpu_finish.restype = C.c_int
    lib.fkgpu_profiles.argtypes = [vp, C.POINTER(i64), C.POINTER(C.POINTER(i64)), C.POINTER(C.POINTER(C.c_uint16))]
    lib.fkgpu_profiles.restype = C.c_int
    lib.fkgpu_profiles_packed.argtypes = [vp, vp, vp, i64, C.POINTER(i64), C.POINTER(i32), i64, C.POINTER(i64),
                                          C.POINTER(C.POINTER(i64)), C.POINTER(C.POINTER(C.c_uint16))]
    lib.fkgpu_profiles_packed.restype = C.c_int
    lib.fkgpu_load_profile_table.argtypes = [vp, C.POINTER(C.c_uint8), i64]
    lib.fkgpu_load_profile_table.restype = C.c_int
    lib.fkgpu_merge_tables.argtypes = [vp, C.POINTER(C.POINTER(C.c_uint8)), C.POINTER(i64), C.c_int, C.c_int, C.POINTER(_Result)]
    lib.fkgpu_merge_tables.restype = C.c_int
    lib.fkgpu_read_counts.argtypes = [vp, C.POINTER(i64)]
    lib.fkgpu_read_counts.restype = C.c_int
    lib.fkgpu_packed_words.argtypes = [i64, C.POINTER(i64), C.POINTER(i64)]
    lib.fkgpu_packed_words.restype = None
    lib.fkgpu_pack_ascii_dev.argtypes = [vp, vp, i64, vp, vp]
    lib.fkgpu_pack_ascii_dev.restype = C.c_int
    lib.fkgpu_count_packed.argtypes = [vp, vp, vp, i64, C.c_int, C.POINTER(_Result)]
    lib.fkgpu_count_packed.restype = C.c_int
    lib.fkgpu_record_bytes.argtypes = [C.c_int]
    lib.fkgpu_record_bytes.restype = C.c_int
    lib.fkgpu_prefix_hist.argtypes = [vp, vp, vp, i64, C.c_int, vp]
    lib.fkgpu_prefix_hist.restype = C.c_int
    lib.fkgpu_scatter_prefix.argtypes = [vp, vp, vp, i64, C.c_int, vp, vp, i64, vp]
    lib.fkgpu_scatter_prefix.restype = C.c_int
    lib.fkgpu_count_records.argtypes = [vp, vp, i64, C.c_int, C.POINTER(_Result)]
    lib.fkgpu_count_records.restype = C.c_int
    lib.fkgpu_launch_count.argtypes = [vp]
    lib.fkgpu_launch_count.restype = i64
    lib.fkgpu_last_stats.argtypes = [vp, C.POINTER(i64)]
    lib.fkgpu_last_stats.restype = C.c_int
    lib.fkgpu_last_path.argtypes = [vp]
    lib.fkgpu_last_path.restype = C.c_int
    lib.fkgpu_stage_times.argtypes = [vp, C.POINTER(C.c_float), C.POINTER(C.c_double)]
    lib.fkgpu_stage_times.restype = C.c_int
    u8p, pp = C.POINTER(C.c_uint8), C.POINTER(vp)
    lib.fkgpu_super_supported.argtypes = [C.c_int]
    lib.fkgpu_super_supported.restype = C.c_int
    lib.fkgpu_entry_bytes.argtypes = [C.c_int]
    lib.fkgpu_entry_bytes.restype = C.c_int
    lib.fkgpu_super_bucket_bits.argtypes = [C.c_int, i64]
    lib.fkgpu_super_bucket_bits.restype = C.c_int
    lib.fkgpu_reads_alloc.argtypes = [vp, i64, pp, pp]
    lib.fkgpu_reads_alloc.restype = C.c_int
    lib.fkgpu_ipc_export.argtypes = [vp, vp, u8p]
    lib.fkgpu_ipc_export.restype = C.c_int
    lib.fkgpu_ipc_open.argtypes = [vp, u8p, pp]
    lib.fkgpu_ipc_open.restype = C.c_int
    lib.fkgpu_ipc_close.argtypes = [vp, vp]
    lib.fkgpu_ipc_close.restype = C.c_int
    lib.fkgpu_super_scan.argtypes = [vp, vp, vp, i64, i64, i64, pp, C.POINTER(i64), C.POINTER(i64), pp, pp, C.POINTER(i32)]
    lib.fkgpu_super_scan.restype = C.c_int
    lib.fkgpu_super_payload.argtypes = [vp, vp, i64, i64, vp, i64, vp]
    lib.fkgpu_super_payload.restype = C.c_int
    lib.fkgpu_super_count.argtypes = [vp, vp, i64, i64, i32, pp, C.POINTER(i64), vp, vp, C.c_int, C.POINTER(_Result), pp,
                                      C.POINTER(i64)]
    lib.fkgpu_super_count.restype = C.c_int
    lib.fkgpu_entries_partition.argtypes = [vp, vp, i64, C.c_int, vp, vp, vp]
    lib.fkgpu_entries_partition.restype = C.c_int
    lib.fkgpu_entries_sort.argtypes = [vp, vp, i64, C.c_int, C.POINTER(_Result)]
    lib.fkgpu_entries_sort.restype = C.c_int
    lib.fkgpu_comm_id.argtypes = [u8p]
    lib.fkgpu_comm_id.restype = C.c_int
    lib.fkgpu_comm_init.argtypes = [vp, C.c_int, C.c_int, u8p]
    lib.fkgpu_comm_init.restype = C.c_int
    lib.fkgpu_comm_info.argtypes = [vp, C.POINTER(i64), C.POINTER(i64)]
    lib.fkgpu_comm_info.restype = C.c_int
    lib.fkgpu_count_packed_multi.argtypes = [vp, vp, vp, i64, C.c_int, C.POINTER(_Result)]
    lib.fkgpu_count_packed_multi.restype = C.c_int
    if path is None:
        _lib = lib
    return lib


EXPORTS = ["fkgpu_create", "fkgpu_destroy", "fkgpu_reset", "fkgpu_last_error", "fkgpu_device_count",
           "fkgpu_ingest", "fkgpu_finish", "fkgpu_profiles", "fkgpu_profiles_packed", "fkgpu_load_profile_table", "fkgpu_merge_tables", "fkgpu_read_counts", "fkgpu_packed_words", "fkgpu_pack_ascii_dev",
           "fkgpu_count_packed", "fkgpu_record_bytes", "fkgpu_prefix_hist", "fkgpu_scatter_prefix",
           "fkgpu_count_records", "fkgpu_launch_count", "fkgpu_last_path", "fkgpu_last_stats", "fkgpu_stage_times",
           "fkgpu_super_supported", "fkgpu_entry_bytes", "fkgpu_super_bucket_bits", "fkgpu_reads_alloc", "fkgpu_ipc_export", "fkgpu_ipc_open",
           "fkgpu_ipc_close", "fkgpu_super_scan", "fkgpu_super_payload", "fkgpu_super_count", "fkgpu_entries_partition", "fkgpu_entries_sort",
           "fkgpu_comm_id", "fkgpu_comm_init", "fkgpu_comm_info", "fkgpu_count_packed_multi"]


class FkResult:
    """Host-side copy of an fkgpu_result (numpy arrays own their data)."""

    def __init__(self, r, copy_table=True):
        self.kmer = r.kmer
        self.kmer_bytes = r.kmer_bytes
        self.nbases, self.nreads = r.nbases, r.nreads
        self.nkmers, self.ndistinct = r.nkmers, r.ndistinct
        self.max_inst = r.max_inst
        self.hist = np.ctypeslib.as_array(r.hist, shape=(HIST_BINS,)).copy()
        self.ntable = r.ntable
        self.table_dev = r.table_dev
        tw = r.kmer_bytes + 2
        self.nruns = max(1, int(r.nruns))
        self.run_ntable = [int(r.run_ntable[i]) for i in range(r.nruns)] if r.nruns > 0 and bool(r.run_ntable) else [int(r.ntable)]
        self._run_ptrs = [r.run_table[i] for i in range(r.nruns)] if r.nruns > 1 and bool(r.run_table) else None
        if copy_table and r.ntable > 0 and bool(r.table):
            self.table = np.ctypeslib.as_array(r.table, shape=(r.ntable * tw,)).copy().reshape(r.ntable, tw)
        elif copy_table and r.ntable > 0 and self._run_ptrs is not None and all(bool(x) for x in self._run_ptrs):
            self.table = self.merged_runs()
        else:
            self.table = None
        self._table_ptr = r.table if (r.ntable > 0 and bool(r.table)) else None
        self.ms_pack, self.ms_count, self.ms_total = r.ms_pack, r.ms_count, r.ms_total

    def view_runs(self):
        """list of (n_i, kmer_bytes + 2) uint8 views, one per sorted run (no copy; same lifetime as view_table)"""
        tw = self.kmer_bytes + 2
        if self._run_ptrs is None:
            return [self.view_table()]
        return [np.ctypeslib.as_array(p, shape=(n * tw,)).reshape(n, tw) if n else np.zeros((0, tw), np.uint8)
                for p, n in zip(self._run_ptrs, self.run_ntable)]

    def merged_runs(self):
        """the runs of a multi-round count merged into one key-ordered table (what Merge_Tables does with the part files,
        table.c:382-394): the runs hold disjoint keys, so the merge is a sort of their concatenation by the key bytes"""
        tw = self.kmer_bytes + 2
        allr = np.concatenate(self.view_runs())
        pad = np.zeros((len(allr), 16), dtype=np.uint8)
        pad[:, :self.kmer_bytes] = allr[:, :self.kmer_bytes]
        w = pad.view(">u8")                                    # key order = order of the two big-endian words
        order = np.lexsort((w[:, 1], w[:, 0]))
        return np.ascontiguousarray(allr[order])

    def view_table(self):
        """(ntable, kmer_bytes + 2) uint8 view of the context's pinned host copy -- no copy; valid until the next
        finish / count / reset / close on the context that produced it."""
        tw = self.kmer_bytes + 2
        if self._table_ptr is None:
            return np.zeros((0, tw), dtype=np.uint8)
        return np.ctypeslib.as_array(self._table_ptr, shape=(self.ntable * tw,)).reshape(self.ntable, tw)


class FastKGPU:
    """One context = one GPU.  Mirrors the stages the reference's driver sequences (FastK.c:498-540):
    ingest() <-> Distribute_Block, finish() <-> Sorting, profiles() <-> Merge_Profiles."""

    def __init__(self, k=40, table_cutoff=0, profile=False, bc_prefix=0, device=0, nthreads=1, reserve_bases=0, mem_limit=0):
        self.lib = load_library()
        cfg = _Config(k, table_cutoff, 1 if profile else 0, bc_prefix, device, nthreads, reserve_bases, mem_limit)
        h = C.c_void_p()
        rc = self.lib.fkgpu_create(C.byref(cfg), C.byref(h))
        if rc != 0:
            raise FkgpuError(f"fkgpu_create -> {rc}: {self.lib.fkgpu_last_error().decode()}")
        self.h = h
        self.k = k
        self.table_cutoff, self.profile = table_cutoff, bool(profile)

    def _chk(self, rc, what):
        if rc != 0:
            raise FkgpuError(f"{what} -> {rc}: {self.lib.fkgpu_last_error().decode()}")

    def close(self):
        if getattr(self, "h", None):
            self.lib.fkgpu_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def reset(self):
        self._chk(self.lib.fkgpu_reset(self.h), "fkgpu_reset")

    def ingest(self, bases: bytes, boff, tid=0, rem=0):
        """bases/boff exactly as in a DATA_BLOCK (FastK.h:87-98)."""
        boff = np.ascontiguousarray(boff, dtype=np.int32)
        n = len(boff) - 1
        self._chk(self.lib.fkgpu_ingest(self.h, tid, bases, boff.ctypes.data_as(C.POINTER(C.c_int32)), n, rem),
                  "fkgpu_ingest")

    def ingest_ptr(self, bases_ptr, boff_ptr, nreads, tid=0, rem=0):
        self._chk(self.lib.fkgpu_ingest(self.h, tid, C.cast(bases_ptr, C.c_char_p),
                                        C.cast(boff_ptr, C.POINTER(C.c_int32)), nreads, rem), "fkgpu_ingest")

    def finish(self, fetch_table=True, copy_table=True):
        r = _Result()
        self._chk(self.lib.fkgpu_finish(self.h, 1 if fetch_table else 0, C.byref(r)), "fkgpu_finish")
        return FkResult(r, copy_table)

    def packed_words(self, npos):
        a, b = C.c_int64(), C.c_int64()
        self.lib.fkgpu_packed_words(npos, C.byref(a), C.byref(b))
        return a.value, b.value

    def pack_ascii_dev(self, d_ascii_ptr, npos, d_seq_ptr, d_val_ptr):
        self._chk(self.lib.fkgpu_pack_ascii_dev(self.h, d_ascii_ptr, npos, d_seq_ptr, d_val_ptr), "fkgpu_pack_ascii_dev")

    def count_packed(self, d_seq_ptr, d_val_ptr, npos, fetch_table=False, copy_table=True):
        r = _Result()
        self._chk(self.lib.fkgpu_count_packed(self.h, d_seq_ptr, d_val_ptr, npos, 1 if fetch_table else 0, C.byref(r)),
                  "fkgpu_count_packed")
        return FkResult(r, copy_table)

    def prefix_hist(self, d_seq_ptr, d_val_ptr, npos, bits, d_hist_ptr):
        self._chk(self.lib.fkgpu_prefix_hist(self.h, d_seq_ptr, d_val_ptr, npos, bits, d_hist_ptr), "fkgpu_prefix_hist")

    def scatter_prefix(self, d_seq_ptr, d_val_ptr, npos, bits, d_hist_ptr, d_rec_ptr, cap, d_off_ptr):
        self._chk(self.lib.fkgpu_scatter_prefix(self.h, d_seq_ptr, d_val_ptr, npos, bits, d_hist_ptr, d_rec_ptr, cap,
                                                d_off_ptr), "fkgpu_scatter_prefix")

    def count_records(self, d_rec_ptr, n, fetch_table=False, copy_table=True):
        r = _Result()
        self._chk(self.lib.fkgpu_count_records(self.h, d_rec_ptr, n, 1 if fetch_table else 0, C.byref(r)),
                  "fkgpu_count_records")
        return FkResult(r, copy_table)

    # ---- multi-GPU count inside the library (NCCL communicator; collective calls) ----------------------------------
    def comm_id(self):
        b = (C.c_uint8 * 128)()
        self._chk(self.lib.fkgpu_comm_id(b), "fkgpu_comm_id")
        return bytes(b)

    def comm_init(self, nranks, rank, comm_id: bytes):
        b = (C.c_uint8 * 128).from_buffer_copy(comm_id)
        self._chk(self.lib.fkgpu_comm_init(self.h, nranks, rank, b), "fkgpu_comm_init")
        self.nranks, self.rank = nranks, rank

    def comm_info(self):
        v = (C.c_int64 * 6)()
        ts = (C.c_int64 * max(1, getattr(self, "nranks", 1)))()
        self._chk(self.lib.fkgpu_comm_info(self.h, v, ts), "fkgpu_comm_info")
        return dict(nranks=int(v[0]), rank=int(v[1]), ntable=int(v[2]), table_offset=int(v[3]), sent_records=int(v[4]),
                    sent_entries=int(v[5]), table_sizes=[int(x) for x in ts])

    def count_packed_multi(self, d_seq_ptr, d_val_ptr, npos, fetch_table=False, copy_table=True):
        r = _Result()
        self._chk(self.lib.fkgpu_count_packed_multi(self.h, d_seq_ptr, d_val_ptr, npos, 1 if fetch_table else 0, C.byref(r)),
                  "fkgpu_count_packed_multi")
        return FkResult(r, copy_table)

    # ---- multi-GPU stages of the super-mer path (include/fastk_gpu.h) ------------------------------------------
    def super_supported(self):
        return bool(self.lib.fkgpu_super_supported(self.k))

    def reads_alloc(self, npos):
        a, b = C.c_void_p(), C.c_void_p()
        self._chk(self.lib.fkgpu_reads_alloc(self.h, npos, C.byref(a), C.byref(b)), "fkgpu_reads_alloc")
        return a.value, b.value

    def ipc_export(self, d_ptr):
        h = (C.c_uint8 * 64)()
        self._chk(self.lib.fkgpu_ipc_export(self.h, d_ptr, h), "fkgpu_ipc_export")
        return bytes(h)

    def ipc_open(self, handle: bytes):
        h = (C.c_uint8 * 64).from_buffer_copy(handle)
        out = C.c_void_p()
        self._chk(self.lib.fkgpu_ipc_open(self.h, h, C.byref(out)), "fkgpu_ipc_open")
        return out.value

    def ipc_close(self, d_ptr):
        self._chk(self.lib.fkgpu_ipc_close(self.h, d_ptr), "fkgpu_ipc_close")

    def super_scan(self, d_seq_ptr, d_val_ptr, npos, npos_total, pos_offset):
        """-> dict(records=device ptr, n, nkmers, hist=device ptr, offsets=device ptr, bits)"""
        rec, hist, offs = C.c_void_p(), C.c_void_p(), C.c_void_p()
        n, nk, bits = C.c_int64(), C.c_int64(), C.c_int32()
        self._chk(self.lib.fkgpu_super_scan(self.h, d_seq_ptr, d_val_ptr, npos, npos_total, pos_offset, C.byref(rec),
                                            C.byref(n), C.byref(nk), C.byref(hist), C.byref(offs), C.byref(bits)),
                  "fkgpu_super_scan")
        return dict(records=rec.value or 0, n=n.value, nkmers=nk.value, hist=hist.value, offsets=offs.value, bits=bits.value)

    def super_payload(self, d_seq_ptr, pos_offset, npos_total, d_rec_ptr, n, d_payload_ptr):
        self._chk(self.lib.fkgpu_super_payload(self.h, d_seq_ptr, pos_offset, npos_total, d_rec_ptr, n, d_payload_ptr),
                  "fkgpu_super_payload")

    def super_count(self, d_rec_ptr, n, npos_total, seq_ptrs, pos_base, want_entries, d_payload_ptr=None, ready_event=None):
        """-> (FkResult with the histogram of this rank's buckets, entries device ptr, # entries)"""
        nr = len(seq_ptrs)
        sp = (C.c_void_p * max(nr, 1))(*seq_ptrs)
        pb = (C.c_int64 * (nr + 1))(*pos_base)
        r = _Result()
        ent, ne = C.c_void_p(), C.c_int64()
        self._chk(self.lib.fkgpu_super_count(self.h, d_rec_ptr, n, npos_total, nr, sp, pb, d_payload_ptr, ready_event,
                                             1 if want_entries else 0, C.byref(r), C.byref(ent), C.byref(ne)),
                  "fkgpu_super_count")
        return FkResult(r, False), (ent.value or 0), ne.value

    def entries_partition(self, d_ent_ptr, n, bits, d_out_ptr, d_hist_ptr, d_off_ptr):
        self._chk(self.lib.fkgpu_entries_partition(self.h, d_ent_ptr, n, bits, d_out_ptr, d_hist_ptr, d_off_ptr),
                  "fkgpu_entries_partition")

    def entries_sort(self, d_ent_ptr, n, fetch_table=False, copy_table=True):
        r = _Result()
        self._chk(self.lib.fkgpu_entries_sort(self.h, d_ent_ptr, n, 1 if fetch_table else 0, C.byref(r)), "fkgpu_entries_sort")
        return FkResult(r, copy_table)

    def last_stats(self):
        v = (C.c_int64 * 8)()
        self._chk(self.lib.fkgpu_last_stats(self.h, v), "fkgpu_last_stats")
        return dict(path=int(v[0]), supermers=int(v[1]), entries=int(v[2]), groups=int(v[3]), rounds=int(v[4]), split_classes=int(v[5]),
                    spilled_kmers=int(v[6]), supermers_expanded=int(v[7]))

    def last_path(self):
        return int(self.lib.fkgpu_last_path(self.h))

    def launch_count(self):
        return int(self.lib.fkgpu_launch_count(self.h))

    def stage_times(self):
        ms = (C.c_float * NSTAGES)()
        by = (C.c_double * NSTAGES)()
        self._chk(self.lib.fkgpu_stage_times(self.h, ms, by), "fkgpu_stage_times")
        return {STAGES[i]: float(ms[i]) for i in range(NSTAGES)}

    @staticmethod
    def _prof_arrays(n, off, prof, copy):
        offs = np.ctypeslib.as_array(off, shape=(n.value + 1,))
        tot = int(offs[-1])
        p = np.ctypeslib.as_array(prof, shape=(max(tot, 1),))[:tot]
        return (offs.copy(), p.copy()) if copy else (offs, p)

    def load_profile_table(self, table):
        """-p:<table>: table = (n, kmer_bytes + 2) uint8 records in key order; finish() then counts nothing and profiles() are
        relative to this table"""
        t = np.ascontiguousarray(table, dtype=np.uint8)
        self._chk(self.lib.fkgpu_load_profile_table(self.h, t.ctypes.data_as(C.POINTER(C.c_uint8)), t.shape[0]),
                  "fkgpu_load_profile_table")

    def merge_tables(self, tables, fetch_table=True):
        """GPU Fastmerge: tables = list of (n_i, kmer_bytes + 2) uint8 arrays in key order -> FkResult of the merged table"""
        ts = [np.ascontiguousarray(t, dtype=np.uint8) for t in tables]
        ptrs = (C.POINTER(C.c_uint8) * len(ts))(*[t.ctypes.data_as(C.POINTER(C.c_uint8)) for t in ts])
        ns = (C.c_int64 * len(ts))(*[t.shape[0] for t in ts])
        r = _Result()
        self._chk(self.lib.fkgpu_merge_tables(self.h, ptrs, ns, len(ts), 1 if fetch_table else 0, C.byref(r)), "fkgpu_merge_tables")
        return FkResult(r, True)

    def profiles(self, copy=True):
        """-> (off int64 [nreads+1], prof uint16): prof[off[r]:off[r+1]] = counts of read r (tid-major read order).
        copy=False returns views of the context's pinned memory (valid until the next call on the context)."""
        n = C.c_int64()
        off = C.POINTER(C.c_int64)()
        prof = C.POINTER(C.c_uint16)()
        self._chk(self.lib.fkgpu_profiles(self.h, C.byref(n), C.byref(off), C.byref(prof)), "fkgpu_profiles")
        return self._prof_arrays(n, off, prof, copy)

    def profiles_packed(self, d_seq_ptr, d_val_ptr, npos, read_start, read_len, copy=True):
        """profiles over a packed stream counted with count_packed (do_profile contexts)"""
        rs = np.ascontiguousarray(read_start, dtype=np.int64)
        rl = np.ascontiguousarray(read_len, dtype=np.int32)
        n = C.c_int64()
        off = C.POINTER(C.c_int64)()
        prof = C.POINTER(C.c_uint16)()
        self._chk(self.lib.fkgpu_profiles_packed(self.h, d_seq_ptr, d_val_ptr, npos, rs.ctypes.data_as(C.POINTER(C.c_int64)),
                                                 rl.ctypes.data_as(C.POINTER(C.c_int32)), len(rs), C.byref(n), C.byref(off),
                                                 C.byref(prof)), "fkgpu_profiles_packed")
        return self._prof_arrays(n, off, prof, copy)
