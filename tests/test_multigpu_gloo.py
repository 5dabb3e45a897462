"""CPU, world_size 2, gloo: the N>1 host logic of fastk_b200/multigpu.py -- splitters from the all-reduced prefix
histogram, the exchange plan, the variable-size all-to-all -- on synthetic 64-bit 'records'.  (The CUDA stages
themselves are covered by the -m gpu tests.)"""
import os
import socket

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from fastk_b200 import multigpu

BITS = 6


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    rng = np.random.default_rng(100 + rank)
    n = 5000 + 700 * rank
    # skewed keys, like canonical k-mers (low prefixes over-represented)
    keys = (rng.random(n) ** 2 * (1 << 62)).astype(np.int64) * 2
    pref = (keys.astype(np.uint64) >> np.uint64(64 - BITS)).astype(np.int64)
    order = np.argsort(pref, kind="stable")
    keys, pref = keys[order], pref[order]
    hist = torch.from_numpy(np.bincount(pref, minlength=1 << BITS).astype(np.int64))
    offs = torch.zeros((1 << BITS) + 1, dtype=torch.int64)
    offs[1:] = torch.cumsum(hist, 0)
    ghist = hist.clone()
    dist.all_reduce(ghist)
    beg = multigpu.splitters_from_hist(ghist, world)
    assert beg[0] == 0 and beg[-1] == (1 << BITS) and all(beg[i] <= beg[i + 1] for i in range(world))
    send_counts = multigpu.exchange_plan(offs, beg)
    assert sum(send_counts) == n
    recs = torch.from_numpy(keys.reshape(-1, 1).copy())
    recv, recv_counts = multigpu.exchange_records(recs, send_counts)
    got = recv[:sum(recv_counts), 0].numpy()
    gp = (got.astype(np.uint64) >> np.uint64(64 - BITS)).astype(np.int64)
    assert ((gp >= beg[rank]) & (gp < beg[rank + 1])).all(), "received a record outside the owned prefix range"
    # conservation: the global multiset of keys is preserved, and loads are balanced to within one prefix bin
    tot = torch.tensor([int(got.sum() % (1 << 40)), len(got), int(keys.sum() % (1 << 40)), n], dtype=torch.int64)
    dist.all_reduce(tot)
    owned = torch.tensor([len(got)], dtype=torch.int64)
    sizes = [torch.zeros(1, dtype=torch.int64) for _ in range(world)]
    dist.all_gather(sizes, owned)
    if rank == 0:
        q.put((int(tot[1]), int(tot[3]), int(tot[0] % (1 << 40)), int(tot[2] % (1 << 40)), [int(s) for s in sizes],
               int(ghist.max())))
    dist.destroy_process_group()


def test_prefix_exchange_world2():
    world = 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(120)
        assert p.exitcode == 0
    nrecv, nsent, srecv, ssent, sizes, maxbin = q.get(timeout=10)
    assert nrecv == nsent and srecv == ssent
    assert abs(sizes[0] - sizes[1]) <= 2 * maxbin


def test_splitters_rule_matches_reference_panel_split():
    # Appendix C of SURVEY.md (MSDsort.c:330-352) on a hand-checked vector
    h = np.array([5, 0, 3, 2, 0, 0, 10, 1], dtype=np.int64)      # total 21
    assert multigpu.splitters_from_hist(h, 2) == [0, 4, 8]         # 10 >= 21//2 after bin 3
    assert multigpu.splitters_from_hist(h, 3) == [0, 3, 7, 8]      # thr 7 -> bin 2 (8); thr 14 -> bin 6 (20)
    assert multigpu.splitters_from_hist(np.zeros(4, np.int64), 3) == [0, 1, 2, 4]


def _panel_split_loop(h, world):
    """msd_sort's panel split (MSDsort.c:330-352) written as the reference writes it: one pass over the bins."""
    total = int(sum(h))
    beg, n, s, thr = [0], 0, 0, total // world
    for x in range(len(h)):
        s += int(h[x])
        if s >= thr and n < world - 1:
            n += 1
            beg.append(x + 1)
            thr = (total * (n + 1)) // world
    while len(beg) < world:
        beg.append(len(h))
    beg.append(len(h))
    return beg


def test_splitters_equal_the_one_pass_rule_on_random_histograms():
    rng = np.random.default_rng(7)
    for trial in range(300):
        nb = int(rng.integers(1, 70))
        world = int(rng.integers(1, 9))
        kind = trial % 4
        if kind == 0:
            h = rng.integers(0, 50, nb)
        elif kind == 1:
            h = (rng.random(nb) ** 6 * 1000).astype(np.int64)           # a few heavy bins
        elif kind == 2:
            h = np.zeros(nb, np.int64)
            h[rng.integers(0, nb)] = 1000                                # everything in one bin
        else:
            h = rng.integers(0, 3, nb)
        got = multigpu.splitters_from_hist(h.astype(np.int64), world)
        assert got == _panel_split_loop(h, world), (list(h), world, got)
        assert got[0] == 0 and got[-1] == nb and len(got) == world + 1
        assert all(got[i] <= got[i + 1] for i in range(world))
