"""CPU: the C-ABI library loads, exports every symbol include/fastk_gpu.h declares, and fails loudly
(no CPU fallback) when there is no CUDA device."""
import ctypes
import os
import re

import pytest

import fastk_b200
from fastk_b200 import lib as fklib

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    h = open(os.path.join(ROOT, "include", "fastk_gpu.h")).read()
    h = re.sub(r"/\*.*?\*/", "", h, flags=re.S)
    return sorted(set(re.findall(r"\b(fkgpu_[a-z_0-9]+)\s*\(", h)))


def test_library_exports_every_declared_symbol():
    assert os.path.exists(fklib.LIB_PATH), "build first: python -c 'import __graft_entry__ as g; g.build()'"
    dll = ctypes.CDLL(fklib.LIB_PATH)
    syms = declared_symbols()
    assert len(syms) >= 17
    for s in syms:
        assert hasattr(dll, s), f"{s} declared in include/fastk_gpu.h but not exported"
    assert sorted(fklib.EXPORTS) == syms, "fastk_b200/lib.py binds a different symbol set than the header declares"


def test_binding_prototypes_load():
    lib = fastk_b200.load_library()
    assert lib.fkgpu_record_bytes(21) == 8 and lib.fkgpu_record_bytes(40) == 16 and lib.fkgpu_record_bytes(63) == 16
    a, b = ctypes.c_int64(), ctypes.c_int64()
    lib.fkgpu_packed_words(1000, ctypes.byref(a), ctypes.byref(b))
    assert b.value == 32 + fklib.PACK_PAD and a.value == 64 + fklib.PACK_PAD


def test_no_cpu_fallback_and_argument_errors():
    lib = fastk_b200.load_library()
    if lib.fkgpu_device_count() == 0:
        with pytest.raises(fastk_b200.FkgpuError) as e:
            fastk_b200.FastKGPU(k=40)
        assert "no CUDA device" in str(e.value)
    else:
        with pytest.raises(fastk_b200.FkgpuError):
            fastk_b200.FastKGPU(k=0)
        with pytest.raises(fastk_b200.FkgpuError) as e:
            fastk_b200.FastKGPU(k=65)
        assert "not supported" in str(e.value)


def test_host_program_refuses_without_input():
    exe = os.path.join(ROOT, "fastk_b200", "bin", "FastK")
    if not os.path.exists(exe):
        pytest.skip("host program not built")
    import subprocess
    r = subprocess.run([exe], capture_output=True, text=True)
    assert r.returncode == 1 and "Usage: FastK" in r.stderr


def test_block_feeder_loads_and_rejects_bad_arguments():
    """The producer-thread helper of bench.py's e2e arm (fastk_b200/host/fk_block_feeder.c) links against the library,
    exports fk_feed_blocks and refuses a NULL context without touching the device."""
    so = os.path.join(ROOT, "fastk_b200", "lib", "libfk_feeder.so")
    if not os.path.exists(so):
        pytest.skip("host helpers not built")
    dll = ctypes.CDLL(so)
    dll.fk_feed_blocks.restype = ctypes.c_int
    dll.fk_feed_blocks.argtypes = [ctypes.c_void_p, ctypes.c_int, ctypes.c_void_p, ctypes.c_int64, ctypes.c_int32, ctypes.c_int32, ctypes.c_int]
    buf = ctypes.create_string_buffer(64)
    assert dll.fk_feed_blocks(None, 4, buf, 1, 16, 1, 0) == -3          # FKGPU_E_ARG
    assert dll.fk_feed_blocks(None, 0, buf, 1, 16, 1, 0) == -3
