"""world_size-2 gloo test of the multi-rank reduction logic on CPU: each rank holds a
1024-aligned contiguous slab of patterns, produces level-1 block partials (here from the
oracle, on the GPU from root_lnl), all-gathers them and applies phylo_reduce_partials (the
product's host-side canonical reduction). The result must equal the single-rank value bit
for bit; Fitch lengths are plain integer sums."""
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, q):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import torch
    import torch.distributed as dist

    import bench
    from helpers import dna_gtr_g4
    from oracle.oracle import Oracle
    from phylocaml_b200 import engine, tree

    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    orc = Oracle()
    model = dna_gtr_g4()
    N, T = 5000, 10
    tr = tree.random_tree(T, 3)
    ops, ra, rb, rt, n_nodes = tree.schedule(tr)
    tips = tree.evolve_tips(tr, model, N, seed=4)
    w = np.random.default_rng(1).integers(1, 5, N).astype(float)
    lo, hi = bench.shard_bounds(N, world, rank)
    res = orc.lk_score_tree(model, np.ascontiguousarray(tips[:, lo:hi]), w[lo:hi], ops, n_nodes, ra, rb, rt)
    parts = orc.reduce_blocks(w[lo:hi] * res["site_lnl"])
    sizes = [None] * world
    dist.all_gather_object(sizes, len(parts))
    bufs = [torch.zeros(s, dtype=torch.float64) for s in sizes]
    dist.all_gather(bufs, torch.from_numpy(parts)) if len(set(sizes)) == 1 else None
    if len(set(sizes)) != 1:
        objs = [None] * world
        dist.all_gather_object(objs, parts.tolist())
        bufs = [torch.tensor(o, dtype=torch.float64) for o in objs]
    allp = torch.cat(bufs).numpy()
    lnl = engine.Engine.reduce_partials(_Lib(), allp)
    # Fitch: exact integer allreduce
    chars = tree.random_fitch_chars(T, N, 4, seed=7)
    length = orc.fitch_score_tree(np.ascontiguousarray(chars[:, lo:hi]), None, ops, n_nodes, ra, rb)["length"]
    t = torch.tensor([length], dtype=torch.int64)
    dist.all_reduce(t)
    if rank == 0:
        whole = orc.lk_score_tree(model, tips, w, ops, n_nodes, ra, rb, rt)["lnl"]
        whole_len = orc.fitch_score_tree(chars, None, ops, n_nodes, ra, rb)["length"]
        q.put((lnl, whole, int(t.item()), whole_len))
    dist.barrier()
    dist.destroy_process_group()


class _Lib:
    """just enough of Engine for reduce_partials (host-only entry point, no GPU needed)."""

    def __init__(self):
        from phylocaml_b200 import engine

        self.lib = engine.load()


def test_two_ranks_reproduce_single_rank_bitwise(built):
    import torch.multiprocessing as mp

    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + (os.getpid() % 2000)
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    lnl, whole, length, whole_len = q.get(timeout=180)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert lnl == whole
    assert length == whole_len
