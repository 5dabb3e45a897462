"""GPU (>= 2 devices): the multi-GPU path (super-mer / prefix exchange over NCCL -> local count) against the CPU oracle on
the union of the ranks' reads: identical histogram, scalars and rank-ordered table.  The world is every GPU of the box up
to 8 (the driver's 8-GPU tier runs world 8: the 24-bit bucket geometry of the payload exchange)."""
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu

HERE = os.path.dirname(os.path.abspath(__file__))


def _ngpu():
    import fastk_b200
    return fastk_b200.load_library().fkgpu_device_count()


@pytest.mark.parametrize("k,path,env", [(40, "super-mer", {}), (21, "super-mer", {"FKGPU_MG": "payload"}), (40, "super-mer", {"FKGPU_MG": "peer"}),
                                        (40, "records", {"FKGPU_MG": "records"}), (63, "super-mer", {}), (63, "records", {"FKGPU_MG": "records"})])
def test_multi_gpu_equals_oracle(k, path, env):
    n = _ngpu()
    if n < 2:
        pytest.skip("needs >= 2 GPUs (run with gpurun --gpus 2)")
    world = 8 if n >= 8 else (4 if n >= 4 else 2)
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}",
           "--master-addr", "127.0.0.1", "--master-port", "29541", os.path.join(HERE, "mgpu_worker.py"), str(k), path]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=600, env=dict(os.environ, **env))
    assert r.returncode == 0 and "MGPU_OK" in r.stdout, r.stdout[-2000:] + r.stderr[-3000:]


@pytest.mark.parametrize("k,path,env", [(40, "super-mer", {"FKGPU_MG": "payload"}), (50, "super-mer", {"FKGPU_MG": "peer"}),
                                        (21, "super-mer", {"FKGPU_MG": "payload"}), (40, "records", {"FKGPU_MG": "records"}),
                                        (63, "super-mer", {"FKGPU_MG": "payload"}), (57, "super-mer", {"FKGPU_MG": "peer"})])
def test_multi_gpu_stages_world1(k, path, env):
    _run_world1(k, path, env, "py")


@pytest.mark.parametrize("k", [40, 21, 63])
def test_multi_gpu_in_library_equals_oracle(k):
    """the exchange inside the C library (fkgpu_comm_init + fkgpu_count_packed_multi, NCCL bound at run time): every GPU
    of the box up to 8; the union of the ranks' reads against the oracle"""
    n = _ngpu()
    if n < 2:
        pytest.skip("needs >= 2 GPUs (run with gpurun --gpus 2)")
    world = 8 if n >= 8 else (4 if n >= 4 else 2)
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}",
           "--master-addr", "127.0.0.1", "--master-port", "29545", os.path.join(HERE, "mgpu_worker.py"), str(k), "super-mer", "c"]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0 and "MGPU_OK" in r.stdout, r.stdout[-2000:] + r.stderr[-3000:]


@pytest.mark.parametrize("k", [40, 21, 63])
def test_multi_gpu_in_library_world1(k):
    """the same collective code path with a communicator of ONE rank: runs on a single-GPU box"""
    _run_world1(k, "super-mer", {}, "c")


def _run_world1(k, path, env, impl):
    """The same staged pipeline (scan -> exchange -> count -> entry exchange -> sort) with a world of ONE rank: runs on a
    single-GPU box, so every stage entry point of the multi-GPU path is parity-checked there too."""
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=1",
           "--master-addr", "127.0.0.1", "--master-port", "29543", os.path.join(HERE, "mgpu_worker.py"), str(k), path, impl]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=600, env=dict(os.environ, **env))
    assert r.returncode == 0 and "MGPU_OK" in r.stdout, r.stdout[-2000:] + r.stderr[-3000:]
