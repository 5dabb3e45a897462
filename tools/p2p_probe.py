"""Diagnostic (torchrun): random 8-byte reads out of every peer's IPC-mapped buffer, one peer at a time and all at once."""
import os, sys, time
os.environ["FKGPU_MG"] = "peer"          # map every peer's buffer (CUDA IPC) whatever the world size
import torch, torch.distributed as dist
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from fastk_b200 import FastKGPU, multigpu

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
eng = FastKGPU(k=40, table_cutoff=0, device=local)
npos = 2_000_000_000                       # 0.5 GB of seq words
mg = multigpu.MultiGPUCounter(eng, world, rank, dev)
seq, val = mg.alloc_reads(npos)
nw = npos // 16 // 2                        # int64 words
views = [multigpu.device_view(p, nw, dev) for p in mg.peer_seq]
views[rank].fill_(rank + 1)
torch.cuda.synchronize(); dist.barrier(device_ids=[local])
g = torch.Generator(device=dev); g.manual_seed(rank)
n = 32_000_000
idx = torch.randint(0, nw, (n,), device=dev, generator=g)

def timed(fn, reps=3):
    fn(); torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(reps): fn()
    torch.cuda.synchronize()
    return (time.perf_counter() - t0) / reps

out = []
for p in range(world):
    dist.barrier(device_ids=[local])
    if rank == 0:
        t = timed(lambda: views[p][idx])
        out.append(f"rank0 <- peer{p}: {n/t/1e9:.2f} G reads/s ({n*32/t/1e9:.0f} GB/s of 32-B sectors)")
    dist.barrier(device_ids=[local])
# everyone reads from every peer at once (idx spread over peers)
dist.barrier(device_ids=[local])
def allpeers():
    for p in range(world):
        views[p][idx[p::world]]
t = timed(allpeers)
out.append(f"all ranks <- all peers at once: {n/t/1e9:.2f} G reads/s per rank")
tt = torch.tensor([t], device=dev); dist.all_reduce(tt, op=dist.ReduceOp.MAX)
if rank == 0:
    for o in out: print(o)
    print(f"max over ranks: {n/float(tt)/1e9:.2f} G reads/s per rank", flush=True)
dist.barrier(device_ids=[local])
mg.close_peers(); eng.close(); dist.destroy_process_group()
