"""GPU: the two host programs end to end, files on disk compared with the reference's golden outputs:
  fastk_b200/bin/FastK            our C host (own reader / writers) over the C ABI
  integration/_build/FastK_refhost the REFERENCE's own FastK.c + io.c + table.c + libfastk.c linked with fastk_shim.c
Parity tiers: T0 .hist byte-identical; T1 .ktab stub byte-identical, part count = -T, concatenated
payload byte-identical, parts cut on first-byte boundaries; P1 every decoded profile identical."""
import os
import shutil
import subprocess

import numpy as np
import pytest

import util

pytestmark = pytest.mark.gpu

OURS = os.path.join(util.ROOT, "fastk_b200", "bin", "FastK")
REFHOST = os.path.join(util.ROOT, "integration", "_build", "FastK_refhost")


def run_cli(exe, g, d, extra=()):
    cmd = [exe, "-k%d" % g["k"], "-t%d" % g["t"], "-p", "-T%d" % g["T"], "-P" + d, "-N" + os.path.join(d, "out")] + list(extra) + [g["src"]]
    r = subprocess.run(cmd, capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    return r


def check_outputs(d, g, oracle_lib, prof_parts=None):
    h = util.read_hist_file(os.path.join(d, "out.hist"))
    assert [h["k"], h["low"], h["high"], h["ilow"], h["max_inst"]] == g["hist_header"]
    assert np.array_equal(h["hist"][1:], g["hist"][1:])
    kt = util.read_ktab_files(d, "out")
    assert kt["stub"] == g["ktab_stub"], "ktab stub differs from the reference"
    assert kt["nparts"] == g["T"] and kt["cutoff"] == g["t"]
    assert kt["payload"] == g["ktab_payload"], "concatenated table payload differs from the reference"
    util.check_parts_on_first_byte_boundaries(kt)
    prof, off, nparts = util.decode_prof_files(d, "out", oracle_lib)
    assert np.array_equal(off, g["prof_off"]) and np.array_equal(prof, g["prof"])
    if prof_parts is not None:
        assert nparts == prof_parts


@pytest.mark.parametrize("name", util.golden_cases())
def test_our_fastk_cli(oracle_lib, name, tmp_path):
    assert os.path.exists(OURS), "host program not built"
    g = util.golden(name)
    run_cli(OURS, g, str(tmp_path), ["-v"])
    check_outputs(str(tmp_path), g, oracle_lib)


def test_our_fastk_cli_gz_and_default_names(oracle_lib, tmp_path):
    g = util.golden("c1_k40")
    src = os.path.join(str(tmp_path), "reads.fa")
    shutil.copy(g["src"], src)
    subprocess.check_call(["gzip", src])
    r = subprocess.run([OURS, "-k40", "-t1", "-p", "-T4", src + ".gz"], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    for f in ("reads.hist", "reads.ktab", "reads.prof", ".reads.ktab.4", ".reads.pidx.1"):
        assert os.path.exists(os.path.join(str(tmp_path), f)), f
    os.rename(os.path.join(str(tmp_path), "reads.hist"), os.path.join(str(tmp_path), "out.hist"))
    h = util.read_hist_file(os.path.join(str(tmp_path), "out.hist"))
    assert np.array_equal(h["hist"][1:], g["hist"][1:])


@pytest.mark.parametrize("name", util.golden_cases())
def test_reference_host_with_gpu_shim(oracle_lib, name, tmp_path):
    """The drop-in proof: the reference's unmodified driver, input module and table merger running on our library."""
    if not os.path.exists(REFHOST):
        pytest.skip("integration/_build/FastK_refhost not built (needs /root/reference at build time)")
    g = util.golden(name)
    run_cli(REFHOST, g, str(tmp_path))
    check_outputs(str(tmp_path), g, oracle_lib, prof_parts=g.get("prof_parts"))


def test_cli_error_paths(tmp_path):
    r = subprocess.run([OURS, "-k40", os.path.join(str(tmp_path), "missing")], capture_output=True, text=True)
    assert r.returncode == 1 and "Cannot find" in r.stderr
    r = subprocess.run([OURS, "-k99", util.golden("c1_k40")["src"]], capture_output=True, text=True)
    assert r.returncode == 1 and "not supported" in r.stderr


@pytest.mark.parametrize("exe_name", ["ours", "refhost"])
def test_cli_multi_round_with_small_memory(oracle_lib, exe_name, tmp_path):
    """-M (ours) / FASTK_GPU_MEM_GB (reference host + shim) far below the one-round working set: the count takes several
    rounds; our writer merges the runs, the reference's own Merge_Tables merges them as NPARTS part files (table.c:382-394).
    Files must still equal the reference's golden outputs byte for byte."""
    g = dict(util.golden("c1_k40"))
    d = str(tmp_path)
    if exe_name == "ours":
        assert os.path.exists(OURS)
        cmd, env = [OURS, "-k40", "-t1", "-T4", "-M0.006", "-v", "-P" + d, "-N" + os.path.join(d, "out"), g["src"]], dict(os.environ)
    else:
        if not os.path.exists(REFHOST):
            pytest.skip("integration/_build/FastK_refhost not built")
        cmd, env = [REFHOST, "-k40", "-t1", "-T4", "-v", "-P" + d, "-N" + os.path.join(d, "out"), g["src"]], dict(os.environ, FASTK_GPU_MEM_GB="0.006")
    r = subprocess.run(cmd, capture_output=True, text=True, env=dict(env, FKGPU_VERBOSE="1"))
    assert r.returncode == 0, r.stderr
    import re
    m = re.search(r"multi-round count: (\d+) rounds", r.stderr)
    assert m and int(m.group(1)) > 1, r.stderr
    h = util.read_hist_file(os.path.join(d, "out.hist"))
    assert np.array_equal(h["hist"][1:], g["hist"][1:])
    kt = util.read_ktab_files(d, "out")
    assert kt["stub"] == g["ktab_stub"] and kt["payload"] == g["ktab_payload"]
    util.check_parts_on_first_byte_boundaries(kt)


@pytest.mark.parametrize("exe_name", ["ours", "refhost"])
def test_cli_relative_profiles(oracle_lib, exe_name, tmp_path):
    """FastK -p:<table>: build the table with the program under test (-t2), then profile other reads against it: only
    .prof is written (no .hist, no .ktab: FastK.c:328-337) and every decoded profile equals the reference's own."""
    g = util.golden_relative()
    d = str(tmp_path)
    exe = OURS if exe_name == "ours" else REFHOST
    if not os.path.exists(exe):
        pytest.skip(exe + " not built")
    r = subprocess.run([exe, "-k%d" % g["k"], "-t%d" % g["table_cutoff"], "-T4", "-P" + d, "-N" + os.path.join(d, "tab"), g["table_src"]],
                       capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    r = subprocess.run([exe, "-k%d" % g["k"], "-p:" + os.path.join(d, "tab"), "-T%d" % g["T"], "-P" + d, "-N" + os.path.join(d, "out"), g["src"]],
                       capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    assert not os.path.exists(os.path.join(d, "out.hist")) and not os.path.exists(os.path.join(d, "out.ktab"))
    prof, off, nparts = util.decode_prof_files(d, "out", oracle_lib)
    assert np.array_equal(off, g["prof_off"]) and np.array_equal(prof, g["prof"])
    r = subprocess.run([exe, "-k21", "-p:" + os.path.join(d, "tab"), "-P" + d, "-N" + os.path.join(d, "bad"), g["src"]], capture_output=True, text=True)
    assert r.returncode == 1 and "k-mer size" in r.stderr


def test_cli_multi_gpu_host_program(tmp_path):
    """FASTK_GPUS=<n>: our host program drives one context per GPU (NCCL communicator inside the library), every GPU returns
    its key range of the table and the writer lays the rank-ordered ranges down as one .ktab: files equal the golden ones."""
    import fastk_b200
    n = fastk_b200.load_library().fkgpu_device_count()
    if n < 2:
        pytest.skip("needs >= 2 GPUs (run with gpurun --gpus 2)")
    g = util.golden("c1_k40")
    d = str(tmp_path)
    world = 8 if n >= 8 else (4 if n >= 4 else 2)
    r = subprocess.run([OURS, "-k40", "-t1", "-T%d" % max(g["T"], world), "-v", "-P" + d, "-N" + os.path.join(d, "out"), g["src"]],
                       capture_output=True, text=True, env=dict(os.environ, FASTK_GPUS=str(world)), timeout=600)
    assert r.returncode == 0, r.stderr
    h = util.read_hist_file(os.path.join(d, "out.hist"))
    assert np.array_equal(h["hist"][1:], g["hist"][1:]) and h["max_inst"] == g["hist_header"][4]
    kt = util.read_ktab_files(d, "out")
    assert kt["payload"] == g["ktab_payload"] and kt["stub"][16:] == g["ktab_stub"][16:]
    util.check_parts_on_first_byte_boundaries(kt)
