import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


def _have_gpu():
    try:
        import fastk_b200
        return fastk_b200.load_library().fkgpu_device_count() > 0
    except Exception:
        return False


def pytest_collection_modifyitems(config, items):
    """gpu-marked tests are skipped (not failed) on a box without a CUDA device or without the built library"""
    if any(it.get_closest_marker("gpu") for it in items) and not _have_gpu():
        skip = pytest.mark.skip(reason="no CUDA device / libfastk_gpu.so not built")
        for it in items:
            if it.get_closest_marker("gpu"):
                it.add_marker(skip)


@pytest.fixture(scope="session")
def oracle_lib():
    """TEST INFRASTRUCTURE: builds (if needed) and loads the CPU oracle."""
    so = os.path.join(ROOT, "oracle", "libfastk_oracle.so")
    if not os.path.exists(so):
        subprocess.check_call(["make", "-C", os.path.join(ROOT, "oracle"), "libfastk_oracle.so", "fastk_oracle"])
    import oracle_py
    return oracle_py.load(so)


@pytest.fixture(scope="session")
def ref_bin():
    """Directory of the reference binaries built by oracle/Makefile (None if unavailable)."""
    d = os.path.join(ROOT, "oracle", "_ref")
    if not os.path.exists(os.path.join(d, "FastK")) and os.path.exists("/root/reference/FastK.c"):
        subprocess.call(["make", "-C", os.path.join(ROOT, "oracle"), "ref"])
    return d if os.path.exists(os.path.join(d, "FastK")) else None
