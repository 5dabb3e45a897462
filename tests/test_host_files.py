"""CPU: the product's host-side writers (fastk_b200/host/fk_files.c) fed with oracle results must reproduce
the reference's files (golden vectors): .hist bytes, .ktab stub + payload, decoded .prof."""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

import util

ROOT = util.ROOT


@pytest.fixture(scope="module")
def fkfiles(tmp_path_factory):
    d = tmp_path_factory.mktemp("fkfiles")
    so = os.path.join(str(d), "libfkfiles.so")
    subprocess.check_call(["gcc", "-O2", "-shared", "-fPIC", "-o", so, os.path.join(ROOT, "fastk_b200", "host", "fk_files.c")])
    L = C.CDLL(so)
    L.fk_write_hist.argtypes = [C.c_char_p, C.c_char_p, C.c_int, C.POINTER(C.c_int64), C.c_int64]
    L.fk_write_ktab.argtypes = [C.c_char_p, C.c_char_p, C.c_int, C.c_int, C.c_int, C.POINTER(C.c_uint8), C.c_int64]
    L.fk_write_prof.argtypes = [C.c_char_p, C.c_char_p, C.c_int, C.c_int, C.POINTER(C.c_int64), C.POINTER(C.c_int64),
                                C.POINTER(C.c_uint16)]
    L.fk_idx_bytes.argtypes = [C.c_int64, C.c_int]
    return L


@pytest.mark.parametrize("name", util.golden_cases())
def test_writers_reproduce_reference_files(oracle_lib, fkfiles, name, tmp_path):
    g = util.golden(name)
    reads = util.read_seq_file(g["src"])
    r = oracle_lib.count(reads, g["k"], cutoff=g["t"], profiles=True)
    d = str(tmp_path).encode()
    hist = np.ascontiguousarray(r["hist"], dtype=np.int64)
    assert fkfiles.fk_write_hist(d, b"w", g["k"], hist.ctypes.data_as(C.POINTER(C.c_int64)), r["max_inst"]) == 0
    h = util.read_hist_file(os.path.join(str(tmp_path), "w.hist"))
    assert np.array_equal(h["hist"][1:], g["hist"][1:]) and [h["k"], h["low"], h["high"], h["ilow"], h["max_inst"]] == g["hist_header"]
    tab = np.ascontiguousarray(r["table"], dtype=np.uint8)
    assert fkfiles.fk_write_ktab(d, b"w", g["k"], g["t"], g["T"], tab.ctypes.data_as(C.POINTER(C.c_uint8)), len(tab)) == 0
    kt = util.read_ktab_files(str(tmp_path), "w")
    assert kt["stub"] == g["ktab_stub"] and kt["payload"] == g["ktab_payload"]
    util.check_parts_on_first_byte_boundaries(kt)
    off = np.zeros(len(reads) + 1, dtype=np.int64)
    for i, p in enumerate(r["profiles"]):
        off[i + 1] = off[i] + len(p)
    prof = np.concatenate(r["profiles"]).astype(np.uint16) if len(reads) else np.zeros(1, np.uint16)
    nparts = 3
    rbeg = np.array([len(reads) * t // nparts for t in range(nparts + 1)], dtype=np.int64)
    assert fkfiles.fk_write_prof(d, b"w", g["k"], nparts, rbeg.ctypes.data_as(C.POINTER(C.c_int64)),
                                 off.ctypes.data_as(C.POINTER(C.c_int64)), prof.ctypes.data_as(C.POINTER(C.c_uint16))) == 0
    p2, o2, np_ = util.decode_prof_files(str(tmp_path), "w", oracle_lib)
    assert np_ == nparts and np.array_equal(o2, g["prof_off"]) and np.array_equal(p2, g["prof"])


def test_idx_bytes_rule(fkfiles):
    # count.c:1620-1626
    assert fkfiles.fk_idx_bytes(0x40000 - 1, 40) == 1
    assert fkfiles.fk_idx_bytes(0x40000, 40) == 2
    assert fkfiles.fk_idx_bytes(0x4000000, 40) == 2
    assert fkfiles.fk_idx_bytes(0x4000001, 40) == 3
    assert fkfiles.fk_idx_bytes(0x4000001, 11) == 2
    assert fkfiles.fk_idx_bytes(0x4000001, 7) == 1
