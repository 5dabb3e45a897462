"""torchrun worker for tests/test_gpu_multi.py: N ranks count disjoint read sets through the multi-GPU pipeline; rank 0
counts the union with the CPU ORACLE (oracle/libfastk_oracle.so, the checker) and the global histogram, the scalars and
the rank-ordered table must agree with it bit for bit."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from fastk_b200 import FastKGPU, synth, multigpu  # noqa: E402
import oracle_py  # noqa: E402


def reads_of(rank, k):
    genome = synth.random_genome(150_000, 77)          # same genome on every rank: k-mers recur across ranks
    return synth.sample_reads(genome, 6000, 250, 0.004, 1000 + rank, n_rate=0.001, len_jitter=100)


def main():
    k = int(sys.argv[1])
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    eng = FastKGPU(k=k, table_cutoff=1, device=local)
    reads = reads_of(rank, k)
    bases, boff = synth.to_block(reads)
    npos = len(bases)
    a = torch.frombuffer(bytearray(bases + b"\0" * 64), dtype=torch.uint8).to(dev)
    impl = sys.argv[3] if len(sys.argv) > 3 else "py"
    mg = None
    if impl == "c":
        # the exchange inside the C library: one NCCL communicator, fkgpu_count_packed_multi is collective
        idt = torch.zeros(128, dtype=torch.uint8, device=dev)
        if rank == 0:
            idt.copy_(torch.frombuffer(bytearray(eng.comm_id()), dtype=torch.uint8))
        dist.broadcast(idt, 0)
        eng.comm_init(world, rank, idt.cpu().numpy().tobytes())
        sw, vw = eng.packed_words(npos)
        ts, tv = torch.zeros(sw, dtype=torch.int32, device=dev), torch.zeros(vw, dtype=torch.int32, device=dev)
        d_seq, d_val = ts.data_ptr(), tv.data_ptr()
        torch.cuda.synchronize()
        eng.pack_ascii_dev(a.data_ptr(), npos, d_seq, d_val)
        torch.cuda.synchronize()
        res = eng.count_packed_multi(d_seq, d_val, npos, fetch_table=True)
        info = eng.comm_info()
        out = multigpu.MultiResult()
        out.path, out.local, out.kmer_bytes = "super-mer", res, res.kmer_bytes
        out.hist, out.max_inst, out.nkmers, out.ndistinct = res.hist, res.max_inst, res.nkmers, res.ndistinct
        out.ntable, out.table_sizes = info["ntable"], info["table_sizes"]
    else:
        mg = multigpu.MultiGPUCounter(eng, world, rank, dev)
        d_seq, d_val = mg.alloc_reads(npos)        # library-owned: the peers map them over CUDA IPC (super-mer path)
        torch.cuda.synchronize()
        eng.pack_ascii_dev(a.data_ptr(), npos, d_seq, d_val)
        torch.cuda.synchronize()
        out = mg.count_packed(d_seq, d_val, npos, fetch_table=True)
    want_path = sys.argv[2] if len(sys.argv) > 2 else None
    assert want_path is None or out.path == want_path, (out.path, want_path)
    tw = out.kmer_bytes + 2
    mx = max(out.table_sizes)
    mine = torch.zeros((mx, tw), dtype=torch.uint8, device=dev)
    if out.local.ntable:
        mine[:out.local.ntable] = torch.from_numpy(out.local.table).to(dev)
    allt = torch.zeros((world, mx, tw), dtype=torch.uint8, device=dev)
    dist.all_gather_into_tensor(allt, mine)
    ok = 1
    if rank == 0:
        table = np.concatenate([allt[r, :out.table_sizes[r]].cpu().numpy() for r in range(world)])
        union = []
        for r in range(world):
            union += reads_of(r, k)
        want = oracle_py.load(os.path.join(ROOT, "oracle", "libfastk_oracle.so")).count(union, k, cutoff=1)
        try:
            assert out.nkmers == want["nkmers"] and out.ndistinct == want["ndistinct"] and out.max_inst == want["max_inst"]
            assert np.array_equal(out.hist[1:], want["hist"][1:]), "global histogram differs from the oracle"
            assert out.ntable == len(want["table"]) and np.array_equal(table, want["table"]), "rank-ordered table differs from the oracle"
            print(f"MGPU_OK world={world} k={k} path={out.path} oracle=libfastk_oracle kmers={out.nkmers} distinct={out.ndistinct} sizes={out.table_sizes}")
        except AssertionError as e:
            ok = 0
            print("MGPU_FAIL", e)
    dist.barrier(device_ids=[local])
    if mg is not None:
        mg.close_peers()
    eng.close()
    dist.destroy_process_group()
    sys.exit(0 if ok else 1)


if __name__ == "__main__":
    main()
