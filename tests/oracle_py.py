"""TEST INFRASTRUCTURE: ctypes binding of oracle/libfastk_oracle.so (the CPU restatement).
Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg may import this."""
import ctypes as C
import numpy as np


class _Tab(C.Structure):
    _fields_ = [("k", C.c_int), ("kbytes", C.c_int), ("n", C.c_int64),
                ("keys", C.POINTER(C.c_uint8)), ("cnt", C.POINTER(C.c_int64))]


class Oracle:
    def __init__(self, lib):
        self.lib = lib
        L = lib
        L.fko_count.argtypes = [C.c_char_p, C.POINTER(C.c_int64), C.c_int64, C.c_int, C.c_int]
        L.fko_count.restype = C.POINTER(_Tab)
        L.fko_free_table.argtypes = [C.POINTER(_Tab)]
        L.fko_histogram.argtypes = [C.POINTER(_Tab), C.POINTER(C.c_int64), C.POINTER(C.c_int64)]
        L.fko_table_entries.argtypes = [C.POINTER(_Tab), C.c_int, C.POINTER(C.c_uint8)]
        L.fko_table_entries.restype = C.c_int64
        L.fko_profile.argtypes = [C.POINTER(_Tab), C.c_char_p, C.c_int64, C.c_int, C.POINTER(C.c_uint16)]
        L.fko_profile.restype = C.c_int64
        L.fko_encode_profile.argtypes = [C.POINTER(C.c_uint16), C.c_int64, C.POINTER(C.c_uint8)]
        L.fko_encode_profile.restype = C.c_int64
        L.fko_decode_profile.argtypes = [C.POINTER(C.c_uint8), C.c_int64, C.POINTER(C.c_uint16), C.c_int64]
        L.fko_decode_profile.restype = C.c_int64
        L.fko_idx_bytes.argtypes = [C.c_int64, C.c_int]
        L.fko_idx_bytes.restype = C.c_int
        L.fko_run_files.argtypes = [C.c_int, C.POINTER(C.c_char_p), C.c_char_p, C.c_char_p,
                                    C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int]
        L.fko_run_files.restype = C.c_int

    def count(self, reads, k, bc_prefix=0, cutoff=1, profiles=False):
        """reads: list of bytes.  -> dict(hist, max_inst, table (n x (kbytes+2) uint8), nkmers, ndistinct[, profiles])"""
        from fastk_b200.synth import to_block
        bases, boff = to_block(reads)
        boff = np.ascontiguousarray(boff, dtype=np.int64)
        t = self.lib.fko_count(bases, boff.ctypes.data_as(C.POINTER(C.c_int64)), len(reads), k, bc_prefix)
        if not t:
            raise RuntimeError("fko_count failed")
        try:
            hist = np.zeros(32768, dtype=np.int64)
            mi = C.c_int64()
            self.lib.fko_histogram(t, hist.ctypes.data_as(C.POINTER(C.c_int64)), C.byref(mi))
            kb = t.contents.kbytes
            n = self.lib.fko_table_entries(t, cutoff, None)
            tab = np.zeros((max(n, 1), kb + 2), dtype=np.uint8)
            self.lib.fko_table_entries(t, cutoff, tab.ctypes.data_as(C.POINTER(C.c_uint8)))
            tab = tab[:n]
            nd = t.contents.n
            cnts = np.ctypeslib.as_array(t.contents.cnt, shape=(max(nd, 1),))[:nd]
            out = dict(hist=hist, max_inst=mi.value, table=tab, ndistinct=int(nd), nkmers=int(cnts.sum()))
            if profiles:
                pro = []
                for r in reads:
                    buf = np.zeros(max(len(r), 1), dtype=np.uint16)
                    pl = self.lib.fko_profile(t, r, len(r), bc_prefix, buf.ctypes.data_as(C.POINTER(C.c_uint16)))
                    pro.append(buf[:pl].copy())
                out["profiles"] = pro
            return out
        finally:
            self.lib.fko_free_table(t)

    def encode_profile(self, prof):
        prof = np.ascontiguousarray(prof, dtype=np.uint16)
        out = np.zeros(2 * len(prof) + 4, dtype=np.uint8)
        n = self.lib.fko_encode_profile(prof.ctypes.data_as(C.POINTER(C.c_uint16)), len(prof),
                                        out.ctypes.data_as(C.POINTER(C.c_uint8)))
        return out[:n].tobytes()

    def decode_profile(self, code, cap=1 << 24):
        a = np.frombuffer(code, dtype=np.uint8).copy() if len(code) else np.zeros(1, dtype=np.uint8)
        out = np.zeros(cap, dtype=np.uint16)
        n = self.lib.fko_decode_profile(a.ctypes.data_as(C.POINTER(C.c_uint8)), len(code),
                                        out.ctypes.data_as(C.POINTER(C.c_uint16)), cap)
        return out[:n].copy()

    def run_files(self, files, outdir, root, k, table=0, profile=False, bc=0, compress=False, nparts=4):
        arr = (C.c_char_p * len(files))(*[f.encode() for f in files])
        rc = self.lib.fko_run_files(len(files), arr, outdir.encode(), root.encode(), k, table, int(profile), bc,
                                    int(compress), nparts)
        if rc != 0:
            raise RuntimeError("fko_run_files failed")


def load(path):
    return Oracle(C.CDLL(path))
