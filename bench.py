#!/usr/bin/env python
"""bench.py -- Gbases/sec counted on B200, the BASELINE.json metric.

One "step" = one pass of the whole counting hot path (encode + canonicalise -> minimizer super-mers -> bucket
partition -> on-chip count -> key-order sort of the distinct entries -> histogram + table) over one batch of
synthetic reads of the shape BASELINE.json names (`--config`, 1-based index into its `configs`: 2 = HiFi-like k=40 -t1
(default, the configuration the metric is quoted on), 3 = Illumina-like k=21 -t4, 4 = k=40 -t1 -p, 5 = k=63 -t1).

  value   packed reads already resident in HBM when the timed region starts; it ends with the sorted [key][count] runs
          and the histogram resident in PINNED HOST memory
  e2e     the same batch through the reference-facing C ABI with HOST buffers: fkgpu_ingest of DATA_BLOCKs from pinned
          host memory (one ingest thread per host core up to 16, like io.c's ITHREADS), H2D, count, D2H
  parity  outside the timed region, at every N: the GPU result for the reads of the reference arm's FASTA is compared
          byte for byte (every .hist bin, max_inst, the .ktab prefix index and every suffix + count byte) with the files
          the reference FastK (oracle/_ref) writes for that FASTA; a mismatch exits non-zero
  roofline  dominant kernel: algorithmic bytes / CUDA-event time vs the measured HBM copy peak
  cpu_baseline  that same reference run, timed (N=1): the same reads, the same box

`--impl reference` times only the CPU reference arm on the same FASTA.  Under torchrun (N>1) every rank owns its own
reads of the same shape (weak scaling) and minimizer buckets / key ranges are exchanged over NCCL (fastk_b200/multigpu.py).
"""
import argparse
import json
import os
import shutil
import subprocess
import sys
import tempfile
import threading
import time

_OUT = None


def emit(line):
    """The JSON line goes to the process's ORIGINAL stdout; fd 1 itself is pointed at stderr from the start of main()
    (NCCL prints its version banner on stdout, the reference FastK its -v chatter), so stdout carries the line only."""
    print(json.dumps(line), file=_OUT or sys.stdout, flush=True)


def claim_stdout():
    global _OUT
    if _OUT is None:
        sys.stdout.flush()
        _OUT = os.fdopen(os.dup(1), "w")
        os.dup2(2, 1)

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

CONFIGS = {
    2: dict(name="configs[1]", desc="synthetic HiFi-like", read_len=15000, coverage=50.0, genome_mbp=40.0, sub_rate=0.001,
            kmer=40, cutoff=1, profile=False),
    3: dict(name="configs[2]", desc="synthetic Illumina-like", read_len=150, coverage=50.0, genome_mbp=40.0, sub_rate=0.002,
            kmer=21, cutoff=4, profile=False),
    4: dict(name="configs[3]", desc="synthetic HiFi-like", read_len=15000, coverage=50.0, genome_mbp=40.0, sub_rate=0.001,
            kmer=40, cutoff=1, profile=True),
    5: dict(name="configs[4]", desc="synthetic HiFi-like", read_len=15000, coverage=100.0, genome_mbp=20.0, sub_rate=0.001,
            kmer=63, cutoff=1, profile=False),
}


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--config", type=int, default=2, choices=sorted(CONFIGS), help="1-based index into BASELINE.json configs")
    ap.add_argument("-k", "--kmer", type=int, default=None)
    ap.add_argument("--genome-mbp", type=float, default=None, help="random genome size per GPU (Mbp)")
    ap.add_argument("--coverage", type=float, default=None)
    ap.add_argument("--read-len", type=int, default=None)
    ap.add_argument("--sub-rate", type=float, default=None)
    ap.add_argument("--cutoff", type=int, default=None, help="-t<cutoff>")
    ap.add_argument("--ingest-threads", type=int, default=0, help="e2e arm: ingest threads (0 = host cores, at most 16)")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-files", action="store_true", help="skip the FASTA-on-disk -> files-on-disk run of fastk_b200/bin/FastK")
    ap.add_argument("--e2e-feeder", default="c", choices=["c", "py"],
                    help="e2e arm: the producer threads that call fkgpu_ingest per DATA_BLOCK: C pthreads (fastk_b200/host/fk_block_feeder.c, "
                         "what a FastK host has) or Python threads")
    ap.add_argument("--e2e-order", default="rr", choices=["rr", "contig"], help="e2e arm: how DATA_BLOCKs are dealt to the ingest threads")
    ap.add_argument("--no-cpu", action="store_true", help="skip the reference run: no cpu_baseline and NO parity check")
    ap.add_argument("--seed", type=int, default=1234)
    ap.add_argument("--mg-impl", default="c", choices=["c", "py"], help="N>1: the exchange inside the C library (NCCL, fkgpu_count_packed_multi) "
                    "or driven from Python over torch.distributed (fastk_b200/multigpu.py)")
    ap.add_argument("--mem-limit-gb", type=float, default=0.0, help="fkgpu_config.mem_limit (the host's -M): below the one-round "
                    "working set the count takes several rounds")
    ap.add_argument("--device-gen", action="store_true", help="generate the reads on the device chunk by chunk (batches too large "
                    "to stage as ASCII on the host: no e2e arm, no FASTA, parity = invariants only)")
    a = ap.parse_args()
    cfg = CONFIGS[a.config]
    for key in ("kmer", "genome_mbp", "coverage", "read_len", "sub_rate", "cutoff"):
        if getattr(a, key) is None:
            setattr(a, key, cfg[key])
    a.profile = cfg["profile"]
    a.cfg_name, a.cfg_desc = cfg["name"], cfg["desc"]
    a.nreads = max(1, int(a.genome_mbp * 1e6 * a.coverage / a.read_len))
    a.genome_bp = max(int(a.genome_mbp * 1e6), 2 * a.read_len)
    return a


def workload_name(a):
    return (f"{a.cfg_name} scaled to one in-HBM batch per GPU: {a.cfg_desc} {a.read_len} bp reads, "
            f"{a.coverage:g}x of a {a.genome_bp/1e6:g} Mbp random genome ({a.nreads * a.read_len / 1e9:.2f} Gbases/GPU), "
            f"{a.sub_rate*100:g}% subs, FastK -k{a.kmer} -t{a.cutoff}" + (" -p" if a.profile else ""))


def make_rows(a, rank, out=None):
    """the reads of rank `rank`: the ONE generator both arms and the parity check use (fastk_b200/synth.py)"""
    from fastk_b200 import synth
    return synth.workload_rows(a.genome_bp, a.nreads, a.read_len, a.sub_rate, a.seed + 7919 * rank, out=out)


class ClockSampler:
    """SM clock and throttle reasons sampled during the timed region (B200_PROFILING.md).  NVML in-process (one cheap query
    every 50 ms from a thread); a polling nvidia-smi child was measured to stall the driver by tens of milliseconds per step on
    some boxes.  Falls back to nvidia-smi when the NVML binding is missing."""
    Q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.rows, self.p, self.h, self.stopf = [], None, None, False
        self.sm, self.mx, self.reasons = [], None, set()
        try:
            import pynvml
            pynvml.nvmlInit()
            vis = os.environ.get("CUDA_VISIBLE_DEVICES")
            phys = int(vis.split(",")[index]) if vis and vis.split(",")[index].isdigit() else index
            self.nv, self.h = pynvml, pynvml.nvmlDeviceGetHandleByIndex(phys)
            self.mx = float(pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM))
            self.t = threading.Thread(target=self._poll, daemon=True)
            self.t.start()
            return
        except Exception:
            self.h = None
        try:
            self.p = subprocess.Popen(["nvidia-smi", "-i", str(index), "--query-gpu=" + self.Q,
                                       "--format=csv,noheader,nounits", "-lms", "100"],
                                      stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._pump, daemon=True)
            self.t.start()
        except Exception:
            self.p = None

    def _poll(self):
        nv = self.nv
        names = {"hw_slowdown": "nvmlClocksEventReasonHwSlowdown", "hw_thermal_slowdown": "nvmlClocksEventReasonHwThermalSlowdown",
                 "sw_thermal_slowdown": "nvmlClocksEventReasonSwThermalSlowdown", "sw_power_cap": "nvmlClocksEventReasonSwPowerCap"}
        getr = getattr(nv, "nvmlDeviceGetCurrentClocksEventReasons", None) or getattr(nv, "nvmlDeviceGetCurrentClocksThrottleReasons", None)
        while not self.stopf:
            try:
                self.sm.append(float(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM)))
                if getr is not None:
                    r = getr(self.h)
                    for nm, attr in names.items():
                        bit = getattr(nv, attr, None) or getattr(nv, attr.replace("Event", "Throttle"), 0)
                        if bit and (r & bit):
                            self.reasons.add(nm)
            except Exception:
                pass
            time.sleep(0.05)

    def _pump(self):
        for line in self.p.stdout:
            self.rows.append(line.strip())

    def stop(self):
        if self.h is not None:
            self.stopf = True
            self.t.join(timeout=1)
            sm = sorted(self.sm)
            return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": self.mx, "reasons": sorted(self.reasons),
                    "samples": len(sm), "source": "nvml"}
        if self.p is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.p.terminate()
        try:
            self.p.wait(timeout=2)
        except Exception:
            self.p.kill()
        sm, mx, reasons = [], None, set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for row in self.rows:
            f = [x.strip() for x in row.split(",")]
            if len(f) < 6:
                continue
            try:
                sm.append(float(f[0]))
                mx = float(f[1])
            except ValueError:
                continue
            for nm, v in zip(names, f[2:6]):
                if v.lower().startswith("active"):
                    reasons.add(nm)
        sm.sort()
        med = sm[len(sm) // 2] if sm else None
        return {"sm_mhz": med, "sm_max_mhz": mx, "reasons": sorted(reasons), "samples": len(sm), "source": "nvidia-smi"}


# ----------------------------------------------------------------------------------------------------------
# CPU reference arm: oracle/_ref/FastK (the reference's own sources, compiled by oracle/Makefile) on the FASTA of rank 0's reads

def ref_binary():
    p = os.path.join(ROOT, "oracle", "_ref", "FastK")
    return p if os.path.exists(p) else None


def scratch_dir(prefix):
    return tempfile.mkdtemp(prefix=prefix, dir="/dev/shm" if os.path.isdir("/dev/shm") else None)


def run_reference_once(a, fasta, tmpdir, cores):
    """-> (seconds, kind, cores used); leaves <tmpdir>/cpu_out.{hist,ktab} (+ hidden parts) behind"""
    exe = ref_binary()
    out = os.path.join(tmpdir, "cpu_out")
    t0 = time.perf_counter()
    if exe is not None:
        kind = "reference"
        cmd = [exe, f"-k{a.kmer}", f"-t{a.cutoff}", f"-T{cores}", "-M16", f"-P{tmpdir}", f"-N{out}"]
        if a.profile:
            cmd.append("-p")
        subprocess.check_call(cmd + [fasta], stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
    else:
        kind = "port"
        cores = 1
        exe = os.path.join(ROOT, "oracle", "fastk_oracle")
        subprocess.check_call([exe, f"-k{a.kmer}", f"-t{a.cutoff}", f"-N{out}", fasta],
                              stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
    return time.perf_counter() - t0, kind, cores


def files_arm(a, fasta, tmpdir, cores, device, t_ref):
    """SURVEY.md 8(d), reported separately: FASTA on disk -> .hist/.ktab(/.prof) on disk through OUR host program
    (fastk_b200/bin/FastK, same options as the reference run beside it, process start and CUDA initialisation included),
    and the two runs' files compared (tiers T0 / T1).  Never fails the bench: an error is reported in the object."""
    exe = os.path.join(ROOT, "fastk_b200", "bin", "FastK")
    if not os.path.exists(exe):
        return {"error": "fastk_b200/bin/FastK not built"}
    try:
        from fastk_b200 import formats
        out = os.path.join(tmpdir, "gpu_out")
        cmd = [exe, "-v", f"-k{a.kmer}", f"-t{a.cutoff}", f"-T{cores}", f"-P{tmpdir}", f"-N{out}"] + (["-p"] if a.profile else []) + [fasta]
        env = dict(os.environ, FASTK_GPU=str(device))
        t0 = time.perf_counter()
        r = subprocess.run(cmd, stdout=subprocess.DEVNULL, stderr=subprocess.PIPE, text=True, env=env, timeout=900)
        t = time.perf_counter() - t0
        if r.returncode != 0:
            return {"error": "exit %d: %s" % (r.returncode, r.stderr[-300:])}
        phases = [ln.strip() for ln in r.stderr.splitlines() if ln.startswith("Total Resources")]
        bad = formats.compare_fastk_outputs(tmpdir, "cpu_out", "gpu_out", table=a.cutoff > 0)
        nb = a.nreads * a.read_len
        return {"value": nb / t / 1e9, "unit": "Gbases/s", "wall_s": round(t, 3), "reference_wall_s": round(t_ref, 3),
                "what": "FASTA on tmpfs -> .hist/.ktab" + ("/.prof" if a.profile else "") + " on tmpfs: " + " ".join(
                    [os.path.relpath(cmd[0], ROOT)] + cmd[1:4] + [f"-T{cores}"] + (["-p"] if a.profile else [])) +
                        "; one run, process start + CUDA initialisation included; host-bound (parse, pinned allocation, file "
                        "writes): the GPU part of such a run is what `e2e` times",
                "phases": phases[-1] if phases else None,
                "files_equal_reference": not bad, "mismatches": bad}
    except Exception as e:                                   # noqa: BLE001 -- a side measurement must not end the bench
        return {"error": str(e)[:300]}


def sample_text(a, cores, extra=""):
    return (f"{a.nreads} reads x {a.read_len} bp = {a.nreads * a.read_len / 1e9:.3f} Gbases ({a.coverage:g}x of "
            f"{a.genome_bp/1e6:.1f} Mbp): the FASTA of rank 0's reads (the whole workload at N=1), on tmpfs, "
            f"FastK -k{a.kmer} -t{a.cutoff}{' -p' if a.profile else ''} -T{cores} -M16{extra}")


def reference_arm(a, rank):
    if rank != 0:
        return
    from fastk_b200 import synth
    cores = os.cpu_count() or 1
    tmpdir = scratch_dir("fastk_ref_")
    try:
        fasta = os.path.join(tmpdir, "reads.fasta")
        nb = synth.write_rows_fasta(make_rows(a, 0), fasta)
        for _ in range(a.warmup):
            run_reference_once(a, fasta, tmpdir, cores)
        ts, kind, used = [], "reference", cores
        for _ in range(a.steps):
            t, kind, used = run_reference_once(a, fasta, tmpdir, cores)
            ts.append(t)
    finally:
        shutil.rmtree(tmpdir, ignore_errors=True)
    tot = sum(ts)
    val = nb * a.steps / tot / 1e9
    line = {"impl": "reference", "metric": "Gbases/sec counted (k=%d)" % a.kmer, "value": val, "unit": "Gbases/s",
            "n_gpus": a.gpus, "steps": a.steps, "warmup": a.warmup, "ms_per_step": 1e3 * tot / a.steps,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "u64", "data": "synthetic",
            "config": {"workload": workload_name(a)},
            "cpu_baseline": {"value": val, "unit": "Gbases/s", "cores": used, "kind": kind,
                             "sample": sample_text(a, used, " per step; FASTA parse and file writes included")},
            "e2e": {"value": val, "unit": "Gbases/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    emit(line)


# ----------------------------------------------------------------------------------------------------------

def device_generate_packed(torch, dev, eng, a, rank, seq_ptr, val_ptr):
    """same read model as synth.workload_rows, drawn with the device RNG and packed chunk by chunk (64 reads x n at a time:
    a multiple of 64 positions, so every chunk starts on whole packed words)"""
    g = torch.Generator(device=dev)
    g.manual_seed(a.seed + 7919 * rank)
    G, L = a.genome_bp, a.read_len
    genome = torch.randint(0, 4, (G,), dtype=torch.uint8, device=dev, generator=g)
    lut = torch.tensor(list(b"ACGT"), dtype=torch.uint8, device=dev)
    ar = torch.arange(L, device=dev)
    per = max(64, ((48 << 20) // (L + 1)) // 64 * 64)
    pos = 0
    for r0 in range(0, a.nreads, per):
        n = min(per, a.nreads - r0)
        st = torch.randint(0, G - L + 1, (n,), device=dev, generator=g)
        r = genome[(st[:, None] + ar[None, :])]
        if a.sub_rate > 0:
            m = torch.rand((n, L), device=dev, generator=g) < a.sub_rate
            add = torch.randint(1, 4, (n, L), dtype=torch.uint8, device=dev, generator=g)
            r = torch.where(m, (r + add) % 4, r)
        flip = torch.rand((n,), device=dev, generator=g) < 0.5
        r = torch.where(flip[:, None], (3 - r).flip(1), r)
        blk = torch.zeros((n * (L + 1) + 64,), dtype=torch.uint8, device=dev)
        blk[:n * (L + 1)].view(n, L + 1)[:, :L] = lut[r.long()]
        assert pos % 64 == 0
        torch.cuda.synchronize()             # torch's stream wrote blk; the library packs on its own stream
        eng.pack_ascii_dev(blk.data_ptr(), n * (L + 1), seq_ptr + (pos // 16) * 4, val_ptr + (pos // 32) * 4)
        torch.cuda.synchronize()
        pos += n * (L + 1)
        del r, m, add, blk
    return pos


def invariants(res_hist, max_inst, nkmers, ndistinct):
    """size-independent properties of any correct count: sum of the histogram = distinct k-mers; sum of c * hist[c]
    below saturation + max_inst = k-mer instances"""
    import numpy as np
    h = np.asarray(res_hist, dtype=np.int64)
    c = np.arange(len(h), dtype=np.int64)
    inst = int((h[:32767] * c[:32767]).sum()) + int(max_inst)
    bad = []
    if int(h[1:].sum()) != int(ndistinct):
        bad.append(f"sum(hist) {int(h[1:].sum())} != distinct {int(ndistinct)}")
    if inst != int(nkmers):
        bad.append(f"sum(c*hist)+max_inst {inst} != k-mer instances {int(nkmers)}")
    return bad


def main():
    a = parse()
    claim_stdout()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))

    if a.impl == "reference":
        reference_arm(a, rank)
        return

    import numpy as np
    import torch
    import torch.distributed as dist
    from fastk_b200 import FastKGPU, formats, synth

    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)

    def barrier():
        if world > 1:
            dist.barrier(device_ids=[local])
        torch.cuda.synchronize()

    k, L, nreads = a.kmer, a.read_len, a.nreads
    nbases = nreads * L
    npos = nreads * (L + 1)

    # ---- the batch: rank r's reads in pinned host memory (DATA_BLOCK layout), then ASCII on the device -> packed
    host_ascii = None
    if not a.device_gen:
        host_ascii = torch.empty((nreads, L + 1), dtype=torch.uint8, pin_memory=True)
        make_rows(a, rank, out=host_ascii.numpy())
        ascii_dev = host_ascii.to(dev, non_blocking=True)
    else:
        a.no_e2e, a.no_cpu = True, True
    nthr = a.ingest_threads or max(1, min(16, (os.cpu_count() or 1) // world))      # ingest threads = ITHREADS of the reference (its -T, FastK.c:367); the host cores are shared by the ranks
    eng = FastKGPU(k=k, table_cutoff=a.cutoff, profile=a.profile, device=local, nthreads=nthr,
                   reserve_bases=0 if a.device_gen else npos, mem_limit=int(a.mem_limit_gb * (1 << 30)))
    runner = None
    cmulti = world > 1 and a.mg_impl == "c"
    if cmulti:
        # one NCCL communicator inside the library: rank 0 makes the id, torch.distributed only hands it round (plumbing)
        idt = torch.zeros(128, dtype=torch.uint8, device=dev)
        if rank == 0:
            idt.copy_(torch.frombuffer(bytearray(eng.comm_id()), dtype=torch.uint8))
        dist.broadcast(idt, 0)
        eng.comm_init(world, rank, bytes(idt.cpu().numpy().tobytes()))
    if world > 1 and not cmulti:
        # packed reads live in library-owned buffers that every peer maps over CUDA IPC (NVLink gathers in the count kernel)
        from fastk_b200 import multigpu
        runner = multigpu.MultiGPUCounter(eng, world, rank, dev)
        seq_ptr, val_ptr = runner.alloc_reads(npos)
    else:
        sw, vw = eng.packed_words(npos)
        d_seq = torch.zeros(sw, dtype=torch.int32, device=dev)
        d_val = torch.zeros(vw, dtype=torch.int32, device=dev)
        seq_ptr, val_ptr = d_seq.data_ptr(), d_val.data_ptr()
    torch.cuda.synchronize()                 # the library packs on its own stream: the ASCII must have landed
    if not a.device_gen:
        eng.pack_ascii_dev(ascii_dev.data_ptr(), npos, seq_ptr, val_ptr)
        torch.cuda.synchronize()
        del ascii_dev
    else:
        device_generate_packed(torch, dev, eng, a, rank, seq_ptr, val_ptr)
    torch.cuda.empty_cache()

    want_table = a.cutoff > 0
    if cmulti:
        def one_step():
            r = eng.count_packed_multi(seq_ptr, val_ptr, npos, fetch_table=want_table, copy_table=False)
            r.local_ntable, r.ntable = r.ntable, eng.comm_info()["ntable"]         # report the global table size
            r.path, r.exchange = "super-mer", "payload (NCCL inside the library)"
            return r
    elif world > 1:
        def one_step():
            return runner.count_packed(seq_ptr, val_ptr, npos, fetch_table=want_table, copy_table=False)
    elif a.profile:
        rstart = np.arange(nreads, dtype=np.int64) * (L + 1)
        rlen = np.full(nreads, L, dtype=np.int32)

        def one_step():
            r = eng.count_packed(seq_ptr, val_ptr, npos, fetch_table=want_table, copy_table=False)
            r.prof_off, r.prof = eng.profiles_packed(seq_ptr, val_ptr, npos, rstart, rlen, copy=False)
            return r
    else:
        def one_step():
            return eng.count_packed(seq_ptr, val_ptr, npos, fetch_table=want_table, copy_table=False)

    # ---- device-resident arm: packed reads in HBM -> table + histogram in pinned host memory ---------------------
    for _ in range(a.warmup):
        res = one_step()
    l0 = eng.launch_count()
    barrier()
    sampler = ClockSampler(local) if rank == 0 else None
    t0 = time.perf_counter()
    stage_ms = {}
    dev_ms = 0.0
    step_ms = []
    for _ in range(a.steps):
        ts0 = time.perf_counter()
        res = one_step()
        step_ms.append(round(1e3 * (time.perf_counter() - ts0), 2))
        dev_ms += res.ms_total
        for kname, v in (res.stage_ms if runner is not None else eng.stage_times()).items():
            stage_ms[kname] = stage_ms.get(kname, 0.0) + v
    barrier()
    t1 = time.perf_counter()
    clocks = sampler.stop() if sampler else None
    launches = eng.launch_count() - l0
    el = torch.tensor([t1 - t0], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(el, op=dist.ReduceOp.MAX)
    elapsed = float(el.item())
    value = nbases * world * a.steps / elapsed / 1e9
    problems = invariants(res.hist, res.max_inst, res.nkmers, res.ndistinct)
    stats = eng.last_stats() if runner is None else None

    # ---- e2e arm (single GPU): host DATA_BLOCKs -> fkgpu_ingest x nthr threads -> finish -> table in pinned host memory
    e2e, r2 = None, None
    if not a.no_e2e and (world == 1 or cmulti):
        rows_per_block = max(1, min(10000, (1_000_000 - 1) // (L + 1)))
        boff_full = (np.arange(rows_per_block + 1, dtype=np.int64) * (L + 1)).astype(np.int32)
        base_ptr = host_ascii.data_ptr()
        blocks = [(r0, min(nreads, r0 + rows_per_block)) for r0 in range(0, nreads, rows_per_block)]

        def worker(tid):
            # -p: tid-major read order must be file order, so every thread takes a contiguous range of blocks (io.c hands
            # each thread a contiguous file range); otherwise the blocks are dealt round-robin
            if a.profile or a.e2e_order == "contig":
                mine = range(len(blocks) * tid // nthr, len(blocks) * (tid + 1) // nthr)
            else:
                mine = range(tid, len(blocks), nthr)
            for bi in mine:
                r0, r1 = blocks[bi]
                eng.ingest_ptr(base_ptr + r0 * (L + 1), boff_full.ctypes.data, r1 - r0, tid=tid)

        feed = None
        feeder_so = os.path.join(os.path.dirname(os.path.abspath(__file__)), "fastk_b200", "lib", "libfk_feeder.so")
        if a.e2e_feeder == "c" and not os.path.exists(feeder_so):
            print("bench.py: %s is missing (run __graft_entry__.build()); the e2e arm uses Python producer threads" % feeder_so, file=sys.stderr)
        if a.e2e_feeder == "c" and os.path.exists(feeder_so):
            import ctypes as C
            flib = C.CDLL(feeder_so)
            flib.fk_feed_blocks.restype = C.c_int
            flib.fk_feed_blocks.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_int64, C.c_int32, C.c_int32, C.c_int]
            contig = 1 if (a.profile or a.e2e_order == "contig") else 0

            def feed():
                rc = flib.fk_feed_blocks(eng.h, nthr, base_ptr, nreads, L + 1, rows_per_block, contig)
                if rc != 0:
                    raise RuntimeError("fk_feed_blocks: %d: %s" % (rc, eng.lib.fkgpu_last_error().decode()))

        e2e_split = {"ingest_ms": 0.0, "finish_ms": 0.0, "profile_ms": 0.0}

        def e2e_step():
            if world > 1:
                dist.barrier(device_ids=[local])
            ta = time.perf_counter()
            eng.reset()
            if feed is not None:
                feed()
            else:
                th = [threading.Thread(target=worker, args=(t,)) for t in range(nthr)]
                for t in th:
                    t.start()
                for t in th:
                    t.join()
            tb = time.perf_counter()
            r = eng.finish(fetch_table=want_table, copy_table=False)      # collective when a communicator is attached
            tc = time.perf_counter()
            if a.profile:
                r.prof_off, r.prof = eng.profiles(copy=False)
            td = time.perf_counter()
            e2e_split["ingest_ms"] += 1e3 * (tb - ta)
            e2e_split["finish_ms"] += 1e3 * (tc - tb)
            e2e_split["profile_ms"] += 1e3 * (td - tc)
            return r

        for _ in range(max(1, a.warmup - 1)):
            r2 = e2e_step()
        torch.cuda.synchronize()
        for key in e2e_split:
            e2e_split[key] = 0.0
        t0 = time.perf_counter()
        for _ in range(a.steps):
            r2 = e2e_step()
        torch.cuda.synchronize()
        t1 = time.perf_counter()
        if world > 1:
            el2 = torch.tensor([t1 - t0], dtype=torch.float64, device=dev)
            dist.all_reduce(el2, op=dist.ReduceOp.MAX)
            t1 = t0 + float(el2.item())
        d2h = int(r2.ntable * (r2.kmer_bytes + 2) + 32768 * 8)
        if a.profile:
            d2h += int(len(r2.prof) * 2 + len(r2.prof_off) * 8)
        e2e = {"value": nbases * world * a.steps / (t1 - t0) / 1e9, "unit": "Gbases/s",
               "h2d_bytes_per_step": int(npos), "d2h_bytes_per_step": d2h,
               "ms_per_step": 1e3 * (t1 - t0) / a.steps,
               "ingest_ms_per_step": e2e_split["ingest_ms"] / a.steps,
               "finish_ms_per_step": e2e_split["finish_ms"] / a.steps,
               "profile_ms_per_step": e2e_split["profile_ms"] / a.steps,
               "finish_device_ms": r2.ms_total,
               "finish_stage_ms": {kn: round(v, 3) for kn, v in eng.stage_times().items() if v > 0},
               "path": f"fkgpu_ingest ({nthr} {'C' if feed is not None else 'Python'} threads, DATA_BLOCKs in pinned host memory; chunks packed + scanned on the device as "
                       f"they land) -> fkgpu_finish(fetch_table={int(want_table)})" + (" -> fkgpu_profiles" if a.profile else "")}
        if not (r2.nkmers == res.nkmers and r2.ndistinct == res.ndistinct and np.array_equal(r2.hist, res.hist)):
            problems.append("e2e and device-resident arms disagree")

    # ---- parity against the reference FastK on the FASTA of rank 0's reads, and the CPU baseline (the same run) --------
    cpu, parity, files = None, {"checked": False}, None
    if not a.no_cpu and ref_binary() is not None and host_ascii is not None:
        tmpdir = scratch_dir("fastk_cpu_") if rank == 0 else None
        try:
            if rank == 0:
                fasta = os.path.join(tmpdir, "reads.fasta")
                synth.write_rows_fasta(host_ascii.numpy(), fasta)
                cores = os.cpu_count() or 1
                t, kind, used = run_reference_once(a, fasta, tmpdir, cores)
                if world == 1 and not a.no_files:
                    files = files_arm(a, fasta, tmpdir, cores, local, t)
                os.remove(fasta)
                if world == 1:
                    cpu = {"value": nbases / t / 1e9, "unit": "Gbases/s", "cores": used, "kind": kind,
                           "sample": sample_text(a, used, f", one run, {t:.1f} s wall; FASTA parse and file writes included")}
            if world == 1:
                got = r2 if r2 is not None else eng.count_packed(seq_ptr, val_ptr, npos, fetch_table=want_table, copy_table=False)
                table = (got.merged_runs() if got.nruns > 1 else got.view_table()) if want_table else None
                ghist, gmax = got.hist, got.max_inst
                via = "fkgpu_ingest/fkgpu_finish (the e2e arm's last step)" if r2 is not None else "fkgpu_count_packed"
            else:
                # the SAME FASTA at every N: rank 0's reads, dealt out in contiguous slices to the N ranks, through the
                # multi-GPU pipeline; the rank-ordered tables are gathered to rank 0
                full = host_ascii.to(dev) if rank == 0 else torch.empty((nreads, L + 1), dtype=torch.uint8, device=dev)
                dist.broadcast(full, 0)
                r0, r1 = nreads * rank // world, nreads * (rank + 1) // world
                mine = full[r0:r1].contiguous()
                del full
                np2 = (r1 - r0) * (L + 1)
                pad = torch.zeros(64, dtype=torch.uint8, device=dev)
                mine = torch.cat([mine.view(-1), pad])
                if cmulti:
                    sw2, vw2 = eng.packed_words(np2)
                    ds2 = torch.zeros(sw2, dtype=torch.int32, device=dev)
                    dv2 = torch.zeros(vw2, dtype=torch.int32, device=dev)
                    s2, v2 = ds2.data_ptr(), dv2.data_ptr()
                else:
                    s2, v2 = runner.alloc_reads(np2)
                torch.cuda.synchronize()
                eng.pack_ascii_dev(mine.data_ptr(), np2, s2, v2)
                torch.cuda.synchronize()
                if cmulti:
                    got = eng.count_packed_multi(s2, v2, np2, fetch_table=want_table, copy_table=False)
                    info = eng.comm_info()
                    got.table_sizes, got.local, got.path = info["table_sizes"], got, "super-mer"
                    got.local_ntable = got.ntable
                else:
                    got = runner.count_packed(s2, v2, np2, fetch_table=want_table, copy_table=False)
                del mine
                ghist, gmax = got.hist, got.max_inst
                table = None
                if want_table:
                    tw = got.kmer_bytes + 2
                    mx = max(got.table_sizes)
                    loc = torch.zeros((mx, tw), dtype=torch.uint8, device=dev)
                    nloc = got.local_ntable if cmulti else got.local.ntable
                    if nloc:
                        loc[:nloc] = torch.from_numpy(got.local.view_table()[:nloc]).to(dev)
                    parts = [torch.zeros((mx, tw), dtype=torch.uint8, device=dev) for _ in range(world)] if rank == 0 else None
                    dist.gather(loc, parts, dst=0)
                    if rank == 0:
                        table = np.concatenate([parts[r][:got.table_sizes[r]].cpu().numpy() for r in range(world)])
                    del loc, parts
                via = (f"fkgpu_count_packed_multi over {world} ranks (NCCL inside the library)" if cmulti else
                       f"multigpu.count_packed over {world} ranks (exchange: {getattr(got, 'exchange', got.path)})")
            if rank == 0:
                bad = formats.compare_with_fastk_files(tmpdir, "cpu_out", k, a.cutoff, ghist, gmax, table)
                parity = {"checked": True, "ok": not bad, "against": "oracle/_ref/FastK (.hist bins + max_inst, .ktab prefix index, "
                          "every suffix + count byte of the hidden parts)", "via": via,
                          "gbases": round(nbases / 1e9, 3), "table_records": int(0 if table is None else table.shape[0]),
                          "mismatches": bad}
                problems += bad
                if a.profile and r2 is not None:
                    pb = compare_profiles(tmpdir, "cpu_out", r2, nreads)
                    parity["profiles"] = "decoded .prof of every read identical" if not pb else pb
                    problems += pb
        finally:
            if tmpdir:
                shutil.rmtree(tmpdir, ignore_errors=True)

    # ---- roofline of the dominant kernel -----------------------------------------------------------------
    peaks = {}
    pk = os.path.join(ROOT, "MEASURED_PEAKS.json")
    which = "fallback"
    if os.path.exists(pk):
        peaks = json.load(open(pk))
        which = "measured"
    peak = float(peaks.get("hbm_gbs", 6650.0))
    W = 8 if k <= 32 else 16
    N, U = res.nkmers, res.ndistinct
    if stats is not None:
        st = stats
        if world > 1:
            N, U = N // world, U // world            # rank 0's share: its own stage times against per-rank counts
    else:
        # rank 0's share of the job: its own stage times against its own record / entry counts
        st = dict(path=1 if res.path == "super-mer" else 0, supermers=getattr(res, "supermers", 0),
                  entries=getattr(res, "entries", 0), groups=0, rounds=1)
        N, U = N // world, U // world
    ntab = res.ntable // world
    if st["path"] == 1:
        # super-mer path: 8-byte super-mer pointers (bucket|len|position) through the partition (histogram read, scatter
        # read+write, refine 2 reads + write = 6 passes), base gather + (key|count) entries out of the bucket
        # kernel, entries through the weighted key-order sort
        S, E = st["supermers"], st["entries"]
        EB = 16 if k <= 56 else 24
        alg = {"super_scan": nbases * 0.375 + S * 8,
               "super_partition": 6 * S * 8,
               "bucket_count": S * 8 + (N + S * (k - 1)) * 0.25 + E * EB,
               "entry_partition": 3 * E * EB,
               "refine": 3 * E * EB,
               # the weighted sort writes the final table records itself (no staging, no compaction pass)
               "sortcount": E * EB + ntab * (res.kmer_bytes + 2)}
        W = EB
    else:
        alg = {"scan_hist": nbases * 0.375,
               "scan_scatter": nbases * 0.375 + N * W,
               "refine": 3 * N * W,
               "sortcount": N * W + U * (W + 4),
               "compact": U * (W + 4) + ntab * (res.kmer_bytes + 2)}
    stage_ms["super_partition"] = stage_ms.get("super_partition", 0.0) + stage_ms.pop("super_refine", 0.0)   # both levels
    if a.profile and stage_ms.get("profile", 0.0) > 0:
        # -p: the packed reads once more, one 32-byte sector of the hash table per k-mer, one u16 out per k-mer
        alg["profile"] = nbases * 0.375 + N * (32 + 2)
    per_stage = {}
    for s, b in alg.items():
        ms = stage_ms.get(s, 0.0) / a.steps
        per_stage[s] = {"ms": round(ms, 3), "alg_gbytes": round(b / 1e9, 3), "gbs": round(b / 1e9 / (ms / 1e3), 1) if ms > 0 else None}
    dom = max(alg.keys(), key=lambda s: per_stage[s]["ms"])
    traffic = None
    tj = os.path.join(ROOT, "profiles", "traffic.json")
    if os.path.exists(tj):
        try:
            tr = json.load(open(tj))
            if dom in tr and tr.get("kmers") and tr.get("kmer", 40) == k:
                traffic = int(tr[dom] * (N / tr["kmers"]))       # ncu capture of a smaller batch, scaled by k-mers
        except Exception:
            traffic = None
    ach = per_stage[dom]["gbs"] or 0.0
    roofline = {"bound": "hbm", "kernel": dom, "achieved": ach, "peak": peak, "unit": "GB/s", "frac": round(ach / peak, 4),
                "traffic": traffic, "peak_source": which + " (MEASURED_PEAKS.json hbm_gbs)" if which == "measured" else "fallback 6650",
                "pipeline": {"alg_bytes_per_kmer": round(sum(alg.values()) / max(N, 1), 2),
                             "gbs": round(sum(alg.values()) / 1e9 / (dev_ms / a.steps / 1e3), 1) if dev_ms > 0 else None,
                             "frac": round(sum(alg.values()) / 1e9 / (dev_ms / a.steps / 1e3) / peak, 4) if dev_ms > 0 else None},
                "stages": per_stage}

    if world == 1:
        parallelism = "single GPU"
    elif cmulti:
        parallelism = ("1 process/GPU; one NCCL communicator inside the C library: grouped ncclSend/ncclRecv all-to-all of the 8-byte "
                       "super-mer records and of their 32-byte base strings, then of the distinct entries by key prefix")
    elif st["path"] == 1:
        parallelism = ("1 process/GPU; all-to-all of 8-byte super-mer records over NCCL, "
                       + ("bases gathered from peer HBM over NVLink inside the count kernel"
                          if getattr(res, "exchange", "") == "peer-gather" else
                          "their 32-byte base strings in a second all-to-all overlapped with the record partition")
                       + ", then all-to-all of the distinct entries by key prefix")
    else:
        parallelism = "1 process/GPU; prefix-range all-to-all of k-mer records over NCCL"
    if rank == 0:
        line = {"metric": "Gbases/sec counted (k=%d)" % k, "value": value, "unit": "Gbases/s", "n_gpus": world,
                "steps": a.steps, "warmup": a.warmup, "ms_per_step": 1e3 * elapsed / a.steps,
                "device_ms_per_step": dev_ms / a.steps, "step_wall_ms": step_ms,
                "all_stage_ms": {kn: round(v / a.steps, 3) for kn, v in stage_ms.items() if v > 0},
                "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "u64", "data": "synthetic",
                "config": {"workload": workload_name(a), "reads_per_gpu": nreads, "kmers_per_gpu": int(N),
                           "distinct_per_gpu": int(U), "table_records": int(res.ntable),
                           "record_bytes": W, "pipeline": "super-mer" if st["path"] == 1 else "records",
                           "supermer_records": st["supermers"], "rounds": st.get("rounds", 1), "sorted_runs": getattr(res, "nruns", 1),
                           "split_classes": st.get("split_classes", 0), "spilled_kmers": st.get("spilled_kmers", 0), "supermers_expanded": st.get("supermers_expanded", 0),
                           "mem_limit_gb": a.mem_limit_gb, "reads": "generated on the device" if a.device_gen else "host generator",
                           "timed_region": "packed reads resident in HBM -> sorted [key][count] table + histogram in pinned host memory",
                           "l2_policy": "inputs_larger_than_L2 (packed reads %.0f MB, records %.1f GB)"
                           % (npos * 0.375 / 1e6, N * W / 1e9),
                           "parallelism": parallelism},
                "clocks": clocks, "gpu_launches": int(launches), "e2e": e2e, "roofline": roofline, "cpu_baseline": cpu,
                "e2e_files": files,
                "parity_checked": bool(parity.get("checked") and parity.get("ok") and not problems), "parity": parity,
                "invariant_violations": problems}
        emit(line)
    fail = torch.tensor([1 if (rank == 0 and problems) else 0], dtype=torch.int64, device=dev)
    if world > 1:
        dist.all_reduce(fail, op=dist.ReduceOp.MAX)
    if runner is not None:
        dist.barrier(device_ids=[local])
        runner.close_peers()
    eng.close()
    if world > 1:
        dist.destroy_process_group()
    if int(fail.item()):
        if rank == 0:
            print("bench.py: PARITY FAILURE: " + "; ".join(problems), file=sys.stderr)
        sys.exit(1)


def compare_profiles(d, root, r2, nreads):
    """decoded .prof of the reference run vs the profiles of the e2e arm's last step -> list of mismatches"""
    import numpy as np
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import oracle_py                      # the checker: Fetch_Profile restated (libfastk.c:1707-1803)
    import util
    so = os.path.join(ROOT, "oracle", "libfastk_oracle.so")
    orc = oracle_py.load(so)
    prof, off, _ = util.decode_prof_files(d, root, orc)
    bad = []
    if len(off) != nreads + 1 or not np.array_equal(off, r2.prof_off):
        bad.append(".prof read offsets differ")
    elif not np.array_equal(prof, r2.prof):
        bad.append(".prof decoded counts differ")
    return bad


if __name__ == "__main__":
    main()
