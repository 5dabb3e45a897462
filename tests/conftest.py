import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


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
