"""fastk_b200 -- B200 (sm_100a) k-mer counting hot path behind FastK's interfaces.

The product is the C-ABI shared library ``fastk_b200/lib/libfastk_gpu.so`` (sources in
``fastk_b200/csrc``, header ``include/fastk_gpu.h``) plus the C host program in ``fastk_b200/host``.
This Python package is only the thin ctypes binding used by the tests, ``bench.py`` and the
multi-GPU driver; it never computes anything itself and has no CPU fallback.
"""
from .lib import FastKGPU, FkResult, load_library, FkgpuError, LIB_PATH  # noqa: F401
