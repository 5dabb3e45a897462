"""CPU: the reference arm of bench.py (`--impl reference`) prints one JSON line with the keys the driver's contract names.
(The GPU arm needs a device; its line carries the same keys plus roofline / clocks / gpu_launches.)"""
import json
import os
import subprocess
import sys

import util


def test_reference_arm_json_line():
    exe = os.path.join(util.ROOT, "oracle", "_ref", "FastK")
    port = os.path.join(util.ROOT, "oracle", "fastk_oracle")
    assert os.path.exists(exe) or os.path.exists(port), "build first: python -c 'import __graft_entry__ as g; g.build()'"
    r = subprocess.run([sys.executable, os.path.join(util.ROOT, "bench.py"), "--impl", "reference", "--steps", "1",
                        "--warmup", "0", "--genome-mbp", "0.2"],
                       capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [ln for ln in r.stdout.splitlines() if ln.startswith("{")]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["unit"] == "Gbases/s" and d["higher_is_better"] is True
    assert d["metric"].startswith("Gbases/sec counted") and d["value"] > 0 and d["ms_per_step"] > 0
    assert d["n_gpus"] == 1 and d["steps"] == 1 and d["warmup"] == 0 and d["scaling"] == "weak" and d["vs_baseline"] is None
    assert "workload" in d["config"] and d["data"] == "synthetic"
    cb = d["cpu_baseline"]
    assert cb["kind"] in ("reference", "port") and cb["cores"] >= 1 and cb["value"] == d["value"] and cb["sample"]
    assert d["e2e"] == {"value": d["value"], "unit": "Gbases/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}


def test_reference_arm_other_ranks_exit_quietly():
    r = subprocess.run([sys.executable, os.path.join(util.ROOT, "bench.py"), "--impl", "reference", "--gpus", "2"],
                       capture_output=True, text=True, timeout=60, env=dict(os.environ, RANK="1", WORLD_SIZE="2", LOCAL_RANK="1"))
    assert r.returncode == 0 and r.stdout.strip() == ""


def test_parity_checker_against_reference_files(oracle_lib, ref_bin, tmp_path):
    """bench.py's parity check (fastk_b200/formats.compare_with_fastk_files) on CPU: the oracle's count of a workload of
    bench shape must compare clean against the files the reference FastK writes for the same FASTA, and a single flipped
    count byte / histogram bin must be reported."""
    import numpy as np
    import pytest
    from fastk_b200 import formats, synth
    if ref_bin is None:
        pytest.skip("oracle/_ref not built")
    rows = synth.workload_rows(60_000, 400, 3000, 0.002, 99)
    fasta = os.path.join(str(tmp_path), "reads.fasta")
    synth.write_rows_fasta(rows, fasta)
    subprocess.check_call([os.path.join(ref_bin, "FastK"), "-k40", "-t1", "-T4", "-P" + str(tmp_path),
                           "-N" + os.path.join(str(tmp_path), "cpu_out"), fasta], stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
    reads = [bytes(r[:-1]) for r in rows]
    want = oracle_lib.count(reads, 40, cutoff=1)
    assert formats.compare_with_fastk_files(str(tmp_path), "cpu_out", 40, 1, want["hist"], want["max_inst"], want["table"]) == []
    t2 = want["table"].copy()
    t2[len(t2) // 2, -2] ^= 1
    assert formats.compare_with_fastk_files(str(tmp_path), "cpu_out", 40, 1, want["hist"], want["max_inst"], t2)
    h2 = want["hist"].copy()
    h2[3] += 1
    assert formats.compare_with_fastk_files(str(tmp_path), "cpu_out", 40, 1, h2, want["max_inst"], want["table"])
