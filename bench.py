#!/usr/bin/env python
"""bench.py -- Gbases/sec counted (k=40) on B200, the BASELINE.json metric.

One "step" = one pass of the whole counting hot path (encode+canonicalise -> prefix split -> MSD refine ->
in-smem sort/count -> histogram -> table) over one batch of synthetic HiFi-like reads.

  value   device-resident: packed reads already in HBM when the timed region starts, results left in HBM
  e2e     the same batch through the reference-facing C ABI with HOST buffers: fkgpu_ingest of DATA_BLOCKs from
          pinned host memory (one ingest thread per host core up to 16, like io.c's ITHREADS), H2D, count, D2H of the table + histogram
  roofline  dominant kernel (k_bucket_count on the super-mer path): algorithmic bytes / CUDA-event time vs the measured HBM copy peak
  cpu_baseline  the reference FastK (oracle/_ref, built from the reference's own sources) on a bounded sample

`--impl reference` times only that CPU reference arm.  Under torchrun (N>1) every rank owns 1/N of the reads
and the canonical-prefix ranges are exchanged with one NCCL all-to-all (fastk_b200/multigpu.py).
"""
import argparse
import json
import os
import subprocess
import sys
import tempfile
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("-k", "--kmer", type=int, default=40)
    ap.add_argument("--genome-mbp", type=float, default=40.0, help="random genome size per GPU (Mbp)")
    ap.add_argument("--coverage", type=float, default=50.0)
    ap.add_argument("--read-len", type=int, default=15000)
    ap.add_argument("--sub-rate", type=float, default=0.001)
    ap.add_argument("--cutoff", type=int, default=1, help="-t<cutoff>")
    ap.add_argument("--cpu-sample-gbases", type=float, default=0.45)
    ap.add_argument("--ingest-threads", type=int, default=0, help="e2e arm: ingest threads (0 = host cores, at most 16)")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--seed", type=int, default=1234)
    return ap.parse_args()


# ----------------------------------------------------------------------------------------------------------
# synthetic reads (same model as fastk_b200/synth.py, generated on the device for the big batch)

def gen_reads_ascii(torch, dev, genome_bp, nreads, read_len, sub_rate, seed):
    """-> uint8 tensor [nreads, read_len+1] of ASCII reads, each row 0-terminated (DATA_BLOCK layout)."""
    g = torch.Generator(device=dev)
    g.manual_seed(seed)
    genome = torch.randint(0, 4, (genome_bp,), dtype=torch.uint8, device=dev, generator=g)
    lut = torch.tensor(list(b"ACGT"), dtype=torch.uint8, device=dev)
    out = torch.zeros((nreads, read_len + 1), dtype=torch.uint8, device=dev)
    ar = torch.arange(read_len, device=dev)
    chunk = max(1, (64 << 20) // read_len)
    for r0 in range(0, nreads, chunk):
        r1 = min(nreads, r0 + chunk)
        n = r1 - r0
        st = torch.randint(0, genome_bp - read_len + 1, (n,), device=dev, generator=g)
        r = genome[(st[:, None] + ar[None, :])]
        if sub_rate > 0:
            m = torch.rand((n, read_len), device=dev, generator=g) < sub_rate
            add = torch.randint(1, 4, (n, read_len), dtype=torch.uint8, device=dev, generator=g)
            r = torch.where(m, (r + add) % 4, r)
        flip = torch.rand((n,), device=dev, generator=g) < 0.5
        rc = (3 - r).flip(1)
        r = torch.where(flip[:, None], rc, r)
        out[r0:r1, :read_len] = lut[r.long()]
    return out


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled during the timed region (B200_PROFILING.md)."""
    Q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.rows = []
        self.p = None
        try:
            self.p = subprocess.Popen(["nvidia-smi", "-i", str(index), "--query-gpu=" + self.Q,
                                       "--format=csv,noheader,nounits", "-lms", "100"],
                                      stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._pump, daemon=True)
            self.t.start()
        except Exception:
            self.p = None

    def _pump(self):
        for line in self.p.stdout:
            self.rows.append(line.strip())

    def stop(self):
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
        return {"sm_mhz": med, "sm_max_mhz": mx, "reasons": sorted(reasons), "samples": len(sm)}


# ----------------------------------------------------------------------------------------------------------
# CPU reference arm

def ref_binary():
    p = os.path.join(ROOT, "oracle", "_ref", "FastK")
    return p if os.path.exists(p) else None


def write_sample_fasta(args, path, gbases):
    """Bounded sample of the same workload for the CPU arms (numpy, same read model)."""
    import numpy as np
    from fastk_b200 import synth
    nreads = max(1, int(gbases * 1e9 / args.read_len))
    gsize = max(args.read_len * 2, int(nreads * args.read_len / args.coverage))
    rng = np.random.default_rng(args.seed)
    genome = rng.integers(0, 4, gsize, dtype=np.uint8)
    lut = np.frombuffer(b"ACGT", dtype=np.uint8)
    with open(path, "wb") as f:
        for i in range(nreads):
            s = int(rng.integers(0, gsize - args.read_len + 1))
            r = genome[s:s + args.read_len].copy()
            m = rng.random(args.read_len) < args.sub_rate
            r[m] = (r[m] + rng.integers(1, 4, int(m.sum()))) % 4
            if rng.random() < 0.5:
                r = (3 - r)[::-1]
            f.write(b">r%d\n" % i + lut[r].tobytes() + b"\n")
    return nreads * args.read_len, nreads, gsize


def run_reference_once(args, fasta, tmpdir, cores):
    exe = ref_binary()
    t0 = time.perf_counter()
    if exe is not None:
        kind = "reference"
        cmd = [exe, f"-k{args.kmer}", f"-t{args.cutoff}", f"-T{cores}", "-M16", f"-P{tmpdir}",
               f"-N{os.path.join(tmpdir, 'cpu_out')}", fasta]
        subprocess.check_call(cmd, stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
    else:
        kind = "port"
        cores = 1
        exe = os.path.join(ROOT, "oracle", "fastk_oracle")
        subprocess.check_call([exe, f"-k{args.kmer}", f"-t{args.cutoff}", f"-N{os.path.join(tmpdir, 'cpu_out')}", fasta],
                              stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
    return time.perf_counter() - t0, kind, cores


def workload_name(args, per_gpu_gbases):
    return (f"config[1] scaled to one in-HBM batch per GPU: synthetic HiFi-like {args.read_len} bp reads, "
            f"{args.coverage:g}x of a {args.genome_mbp:g} Mbp random genome ({per_gpu_gbases:.2f} Gbases/GPU), "
            f"{args.sub_rate*100:g}% subs, FastK -k{args.kmer} -t{args.cutoff}")


def reference_arm(args, rank):
    if rank != 0:
        return
    cores = os.cpu_count() or 1
    tmpdir = tempfile.mkdtemp(prefix="fastk_ref_", dir="/dev/shm" if os.path.isdir("/dev/shm") else None)
    fasta = os.path.join(tmpdir, "sample.fasta")
    nb, nreads, gsize = write_sample_fasta(args, fasta, args.cpu_sample_gbases)
    for _ in range(args.warmup):
        run_reference_once(args, fasta, tmpdir, cores)
    ts = []
    kind = "reference"
    for _ in range(args.steps):
        t, kind, used = run_reference_once(args, fasta, tmpdir, cores)
        ts.append(t)
    subprocess.call(["rm", "-rf", tmpdir])
    tot = sum(ts)
    val = nb * args.steps / tot / 1e9
    per_gpu = args.genome_mbp * args.coverage / 1e3
    line = {"impl": "reference", "metric": "Gbases/sec counted (k=%d)" % args.kmer, "value": val, "unit": "Gbases/s",
            "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * tot / args.steps,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "u64", "data": "synthetic",
            "config": {"workload": workload_name(args, per_gpu)},
            "cpu_baseline": {"value": val, "unit": "Gbases/s", "cores": used, "kind": kind,
                             "sample": f"{nreads} reads x {args.read_len} bp = {nb/1e9:.3f} Gbases "
                                       f"({args.coverage:g}x of {gsize/1e6:.1f} Mbp) per step, FASTA on tmpfs, "
                                       f"FastK -k{args.kmer} -t{args.cutoff} -T{used}"},
            "e2e": {"value": val, "unit": "Gbases/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line), flush=True)


# ----------------------------------------------------------------------------------------------------------

def main():
    args = parse()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))

    if args.impl == "reference":
        reference_arm(args, rank)
        return

    import numpy as np
    import torch
    import torch.distributed as dist
    from fastk_b200 import FastKGPU

    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)

    def barrier():
        if world > 1:
            dist.barrier(device_ids=[local])
        torch.cuda.synchronize()

    genome_bp = int(args.genome_mbp * 1e6)
    nreads = int(genome_bp * args.coverage / args.read_len)
    k = args.kmer
    nbases = nreads * args.read_len
    npos = nreads * (args.read_len + 1)

    # ---- build the batch: ASCII on the device -> packed (device-resident arm) and pinned host copy (e2e arm)
    ascii_dev = gen_reads_ascii(torch, dev, genome_bp, nreads, args.read_len, args.sub_rate, args.seed + 7919 * rank)
    nthr = args.ingest_threads or max(1, min(16, os.cpu_count() or 1))      # ingest threads = ITHREADS of the reference (its -T, FastK.c:367)
    eng = FastKGPU(k=k, table_cutoff=args.cutoff, device=local, nthreads=nthr, reserve_bases=npos)
    runner = None
    if world > 1:
        # packed reads live in library-owned buffers that every peer maps over CUDA IPC (NVLink gathers in the count kernel)
        from fastk_b200 import multigpu
        runner = multigpu.MultiGPUCounter(eng, world, rank, dev)
        seq_ptr, val_ptr = runner.alloc_reads(npos)
    else:
        sw, vw = eng.packed_words(npos)
        d_seq = torch.zeros(sw, dtype=torch.int32, device=dev)
        d_val = torch.zeros(vw, dtype=torch.int32, device=dev)
        seq_ptr, val_ptr = d_seq.data_ptr(), d_val.data_ptr()
    eng.pack_ascii_dev(ascii_dev.data_ptr(), npos, seq_ptr, val_ptr)
    torch.cuda.synchronize()
    host_ascii = None
    if not args.no_e2e and world == 1:
        host_ascii = torch.empty((nreads, args.read_len + 1), dtype=torch.uint8, pin_memory=True)
        host_ascii.copy_(ascii_dev)
    del ascii_dev
    torch.cuda.empty_cache()

    if world > 1:
        def one_step():
            return runner.count_packed(seq_ptr, val_ptr, npos)
    else:
        def one_step():
            return eng.count_packed(seq_ptr, val_ptr, npos, fetch_table=False)

    # ---- device-resident arm ---------------------------------------------------------------------------
    for _ in range(args.warmup):
        res = one_step()
    l0 = eng.launch_count()
    barrier()
    sampler = ClockSampler(local) if rank == 0 else None
    t0 = time.perf_counter()
    stage_ms = {}
    dev_ms = 0.0
    for _ in range(args.steps):
        res = one_step()
        dev_ms += res.ms_total
        for kname, v in (res.stage_ms if world > 1 else eng.stage_times()).items():
            stage_ms[kname] = stage_ms.get(kname, 0.0) + v
    barrier()
    t1 = time.perf_counter()
    clocks = sampler.stop() if sampler else None
    launches = eng.launch_count() - l0
    el = torch.tensor([t1 - t0], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(el, op=dist.ReduceOp.MAX)
    elapsed = float(el.item())
    value = nbases * world * args.steps / elapsed / 1e9

    # ---- e2e arm (single GPU): host DATA_BLOCKs -> fkgpu_ingest x8 threads -> finish -> table in pinned host memory
    e2e = None
    if host_ascii is not None:
        rows_per_block = max(1, min(10000, (1_000_000 - 1) // (args.read_len + 1)))
        boff_full = (np.arange(rows_per_block + 1, dtype=np.int64) * (args.read_len + 1)).astype(np.int32)
        base_ptr = host_ascii.data_ptr()
        blocks = [(r0, min(nreads, r0 + rows_per_block)) for r0 in range(0, nreads, rows_per_block)]

        def worker(tid):
            for bi in range(tid, len(blocks), nthr):
                r0, r1 = blocks[bi]
                eng.ingest_ptr(base_ptr + r0 * (args.read_len + 1), boff_full.ctypes.data, r1 - r0, tid=tid)

        e2e_split = {"ingest_ms": 0.0, "finish_ms": 0.0}

        def e2e_step():
            ta = time.perf_counter()
            eng.reset()
            th = [threading.Thread(target=worker, args=(t,)) for t in range(nthr)]
            for t in th:
                t.start()
            for t in th:
                t.join()
            tb = time.perf_counter()
            r = eng.finish(fetch_table=True, copy_table=False)
            tc = time.perf_counter()
            e2e_split["ingest_ms"] += 1e3 * (tb - ta)
            e2e_split["finish_ms"] += 1e3 * (tc - tb)
            return r

        for _ in range(max(1, args.warmup - 1)):
            r2 = e2e_step()
        torch.cuda.synchronize()
        e2e_split["ingest_ms"] = e2e_split["finish_ms"] = 0.0
        t0 = time.perf_counter()
        for _ in range(args.steps):
            r2 = e2e_step()
        torch.cuda.synchronize()
        t1 = time.perf_counter()
        e2e = {"value": nbases * args.steps / (t1 - t0) / 1e9, "unit": "Gbases/s",
               "h2d_bytes_per_step": int(npos), "d2h_bytes_per_step": int(r2.ntable * (r2.kmer_bytes + 2) + 32768 * 8),
               "ms_per_step": 1e3 * (t1 - t0) / args.steps,
               "ingest_ms_per_step": e2e_split["ingest_ms"] / args.steps,
               "finish_ms_per_step": e2e_split["finish_ms"] / args.steps,
               "finish_device_ms": r2.ms_total,
               "finish_stage_ms": {kn: round(v, 3) for kn, v in eng.stage_times().items() if v > 0},
               "path": f"fkgpu_ingest ({nthr} threads, DATA_BLOCKs in pinned host memory; chunks packed + scanned on the device as "
                       "they land) -> fkgpu_finish(fetch_table=1)"}
        assert r2.nkmers == res.nkmers and r2.ndistinct == res.ndistinct, "e2e and device-resident arms disagree"

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
    if world == 1:
        st = eng.last_stats()
    else:
        # rank 0's share of the job: its own stage times against its own record / entry counts
        st = dict(path=1 if res.path == "super-mer" else 0, supermers=getattr(res, "supermers", 0),
                  entries=getattr(res, "entries", 0), groups=0)
        N, U = N // world, U // world
    ntab = res.ntable // world
    if st["path"] == 1:
        # super-mer path: 8-byte super-mer pointers (bucket|len|position) through the partition (histogram read, scatter
        # read+write, refine 2 reads + write = 6 passes), base gather + 16-byte (key|count) entries out of the bucket
        # kernel, entries through the weighted key-order sort
        S, E = st["supermers"], st["entries"]
        alg = {"super_scan": nbases * 0.375 + S * 8,
               "super_partition": 6 * S * 8,
               "bucket_count": S * 8 + (N + S * (k - 1)) * 0.25 + E * 16,
               "entry_partition": 3 * E * 16,
               "refine": 3 * E * 16,
               # the weighted sort writes the final table records itself (no staging, no compaction pass)
               "sortcount": E * 16 + ntab * (res.kmer_bytes + 2)}
        W = 16
    else:
        alg = {"scan_hist": nbases * 0.375,
               "scan_scatter": nbases * 0.375 + N * W,
               "refine": 3 * N * W,
               "sortcount": N * W + U * (W + 4),
               "compact": U * (W + 4) + ntab * (res.kmer_bytes + 2)}
    per_stage = {}
    for s, b in alg.items():
        ms = stage_ms.get(s, 0.0) / args.steps
        per_stage[s] = {"ms": round(ms, 3), "alg_gbytes": round(b / 1e9, 3), "gbs": round(b / 1e9 / (ms / 1e3), 1) if ms > 0 else None}
    dom = max(alg.keys(), key=lambda s: per_stage[s]["ms"])
    traffic = None
    tj = os.path.join(ROOT, "profiles", "traffic.json")
    if os.path.exists(tj):
        try:
            tr = json.load(open(tj))
            if dom in tr and tr.get("kmers"):
                traffic = int(tr[dom] * (N / tr["kmers"]))       # ncu capture of a smaller batch, scaled by k-mers
        except Exception:
            traffic = None
    ach = per_stage[dom]["gbs"] or 0.0
    roofline = {"bound": "hbm", "kernel": dom, "achieved": ach, "peak": peak, "unit": "GB/s", "frac": round(ach / peak, 4),
                "traffic": traffic, "peak_source": which + " (MEASURED_PEAKS.json hbm_gbs)" if which == "measured" else "fallback 6650",
                "pipeline": {"alg_bytes_per_kmer": round(sum(alg.values()) / max(N, 1), 2),
                             "gbs": round(sum(alg.values()) / 1e9 / (dev_ms / args.steps / 1e3), 1) if dev_ms > 0 else None,
                             "frac": round(sum(alg.values()) / 1e9 / (dev_ms / args.steps / 1e3) / peak, 4) if dev_ms > 0 else None},
                "stages": per_stage}

    # ---- CPU baseline (rank 0, N=1 only) ------------------------------------------------------------------
    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu:
        cores = os.cpu_count() or 1
        tmpdir = tempfile.mkdtemp(prefix="fastk_cpu_", dir="/dev/shm" if os.path.isdir("/dev/shm") else None)
        fasta = os.path.join(tmpdir, "sample.fasta")
        nb, nr, gsize = write_sample_fasta(args, fasta, args.cpu_sample_gbases)
        t, kind, used = run_reference_once(args, fasta, tmpdir, cores)
        subprocess.call(["rm", "-rf", tmpdir])
        cpu = {"value": nb / t / 1e9, "unit": "Gbases/s", "cores": used, "kind": kind,
               "sample": f"{nr} reads x {args.read_len} bp = {nb/1e9:.3f} Gbases ({args.coverage:g}x of {gsize/1e6:.1f} Mbp), "
                         f"FASTA on tmpfs, one run of FastK -k{k} -t{args.cutoff} -T{used} -M16, {t:.1f} s wall"}

    if rank == 0:
        per_gpu = nbases / 1e9
        line = {"metric": "Gbases/sec counted (k=%d)" % k, "value": value, "unit": "Gbases/s", "n_gpus": world,
                "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * elapsed / args.steps,
                "device_ms_per_step": dev_ms / args.steps,
                "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "u64", "data": "synthetic",
                "config": {"workload": workload_name(args, per_gpu), "reads_per_gpu": nreads, "kmers_per_gpu": int(N),
                           "distinct_per_gpu": int(U), "table_records": int(res.ntable),
                           "record_bytes": W, "pipeline": "super-mer" if st["path"] == 1 else "records",
                           "supermer_records": st["supermers"], "l2_policy": "inputs_larger_than_L2 (packed reads %.0f MB, records %.1f GB)"
                           % (npos * 0.375 / 1e6, N * W / 1e9),
                           "parallelism": (("1 process/GPU; all-to-all of 8-byte super-mer records over NCCL, "
                                            + ("bases gathered from peer HBM over NVLink inside the count kernel"
                                               if getattr(res, "exchange", "") == "peer-gather" else
                                               "their 32-byte base strings in a second all-to-all overlapped with the record partition")
                                            + ", then all-to-all of the distinct entries by key prefix")
                                           if st["path"] == 1 else "1 process/GPU; prefix-range all-to-all of k-mer records over NCCL")
                           if world > 1 else "single GPU"},
                "clocks": clocks, "gpu_launches": int(launches), "e2e": e2e, "roofline": roofline, "cpu_baseline": cpu}
        print(json.dumps(line), flush=True)
    if runner is not None:
        dist.barrier(device_ids=[local])
        runner.close_peers()
    eng.close()
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
