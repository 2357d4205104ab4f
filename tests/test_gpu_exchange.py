"""Device-side scalar exchange (exchange_kernels.cuh): the ranks' level-1 block partials are
written into peer-mapped mailboxes and folded in rank order on the device -- lnL must be bit-identical to
a single-engine evaluation of the whole alignment. One GPU is enough to exercise it: two engines on
their own non-blocking streams, driven by two host threads (the mailboxes are then plain device pointers;
between processes they travel as cudaIpc handles, which bench.py does under torchrun)."""
import threading

import numpy as np
import pytest

from helpers import dna_gtr_g4, setup_lk
from phylocaml_b200 import engine

pytestmark = pytest.mark.gpu


def _streams(n):
    import torch

    return [torch.cuda.Stream() for _ in range(n)]


def test_exchange_world_1_equals_plain_score(eng):
    model = dna_gtr_g4()
    tr, ops, ra, rb, rt, n_nodes, tips = setup_lk(12, 5000, model, seed=4)
    eng.lk_set_model(model)
    eng.lk_set_tips(tips, capacity=n_nodes)
    want = eng.lk_score_tree(ops, ra, rb, rt)
    box, handle = eng.exchange_alloc()
    assert len(handle) == 64
    eng.exchange_set(0, [box])
    eng.set_option(eng.OPT_DEFER_SCALAR, 1)
    try:
        assert np.isnan(eng.lk_score_tree(ops, ra, rb, rt))
        assert eng.lk_exchange_reduce() == want
        assert eng.exchange_sum_u64(123456789012345) == 123456789012345
    finally:
        eng.set_option(eng.OPT_DEFER_SCALAR, 0)


@pytest.mark.parametrize("world,N", [(2, 7000), (3, 5 * 1024 + 17), (2, 1500)])
def test_exchange_between_engines_is_bit_identical_to_one_engine(eng, world, N):
    model = dna_gtr_g4()
    tr, ops, ra, rb, rt, n_nodes, tips = setup_lk(10, N, model, seed=9)
    w = np.random.default_rng(1).integers(1, 9, N).astype(float)
    eng.lk_set_model(model)
    eng.lk_set_tips(tips, weights=w, capacity=n_nodes)
    want = eng.lk_score_tree(ops, ra, rb, rt)
    # shards at multiples of 1024 patterns; the last rank may be empty-handed for a short alignment
    blocks = (N + 1023) // 1024
    bounds = [min(N, 1024 * (blocks * r // world)) for r in range(world)] + [N]
    streams = _streams(world)
    engines = [engine.Engine(0) for _ in range(world)]
    try:
        boxes = []
        for r, en in enumerate(engines):
            en.set_stream(streams[r].cuda_stream)
            boxes.append(en.exchange_alloc()[0])
        got, sums, errs = [None] * world, [None] * world, []

        def run(r):
            try:
                en = engines[r]
                lo, hi = bounds[r], bounds[r + 1]
                en.exchange_set(r, boxes)
                en.lk_set_model(model)
                en.lk_set_tips(np.ascontiguousarray(tips[:, lo:hi]), weights=w[lo:hi], capacity=n_nodes)
                en.set_option(en.OPT_DEFER_SCALAR, 1)
                for _ in range(3):  # the sequence numbers advance in step on every rank
                    en.lk_score_tree(ops, ra, rb, rt)
                    got[r] = en.lk_exchange_reduce()
                sums[r] = en.exchange_sum_u64(10 ** 12 + r)
            except Exception as ex:  # noqa: BLE001
                errs.append(ex)

        if any(bounds[r + 1] == bounds[r] for r in range(world)):
            pytest.skip("an empty shard cannot load tips")
        th = [threading.Thread(target=run, args=(r,)) for r in range(world)]
        [t.start() for t in th]
        [t.join() for t in th]
        assert not errs, errs
        assert all(g == want for g in got), (got, want)
        assert all(s == world * 10 ** 12 + sum(range(world)) for s in sums)
    finally:
        for en in engines:
            en.close()


def test_exchange_times_out_instead_of_hanging(eng):
    """world = 2 but the peer never calls: the kernel gives up after ~2 s and the call fails."""
    model = dna_gtr_g4()
    tr, ops, ra, rb, rt, n_nodes, tips = setup_lk(8, 2048, model, seed=2)
    other = engine.Engine(0)
    try:
        a, _ = eng.exchange_alloc()
        b, _ = other.exchange_alloc()
        eng.exchange_set(0, [a, b])
        eng.lk_set_model(model)
        eng.lk_set_tips(tips, capacity=n_nodes)
        eng.lk_score_tree(ops, ra, rb, rt)
        with pytest.raises(engine.PhyloError):
            eng.lk_exchange_reduce()
    finally:
        other.close()
