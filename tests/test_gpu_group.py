"""phylo_group (several GPUs behind one handle, one process): results against a single engine.
On a one-GPU box the group lists device 0 several times -- the sharding, the worker threads and
the reduction are the same code; with more GPUs visible the last test spreads over all of them.
lnL must be BIT-identical (block partials folded in shard order), Fitch lengths exact."""
import ctypes

import numpy as np
import pytest

from helpers import aa_model, dna_gtr_g4, rel_err, setup_lk
from phylocaml_b200 import engine, tree

pytestmark = pytest.mark.gpu


def _n_devices():
    import torch

    return torch.cuda.device_count()


@pytest.fixture()
def group3(built):
    g = engine.Group([0, 0, 0])
    yield g
    g.close()


@pytest.mark.parametrize("N", [700, 3000, 10000, 300001])
def test_group_lnl_is_bitwise_the_single_engine_lnl(eng, group3, oracle, N):
    """N = 700: one shard, two left empty; 3000: ragged last shard; 300001: the bandwidth kernels."""
    model = dna_gtr_g4()
    tr, ops, ra, rb, rt, n_nodes, tips = setup_lk(20, min(N, 10000), model, seed=23)
    tips = np.ascontiguousarray(np.tile(tips, (1, N // tips.shape[1] + 1))[:, :N])
    w = np.random.default_rng(5).integers(1, 5, N).astype(float)
    eng.lk_set_model(model)
    eng.lk_set_tips(tips, weights=w, capacity=n_nodes)
    one = eng.lk_score_tree(ops, ra, rb, rt)
    group3.lk_set_model(model)
    group3.lk_set_tips(tips, weights=w, capacity=n_nodes)
    bounds = [group3.shard(i) for i in range(3)]
    assert bounds[0][0] == 0 and bounds[-1][1] == N
    assert all(a[1] == b[0] for a, b in zip(bounds, bounds[1:])) and all(lo % 1024 == 0 for lo, hi in bounds if hi > lo)
    many = group3.lk_score_tree(ops, ra, rb, rt)
    assert many == one
    assert np.array_equal(group3.lk_get_site_lnl(), eng.lk_get_site_lnl())
    node = int(ops[-1]["parent"])
    (ca, sa), (cb, sb) = group3.lk_get_clv(node), eng.lk_get_clv(node)
    assert np.array_equal(ca, cb) and np.array_equal(sa, sb)
    if N <= 3000:
        assert rel_err(many, oracle.lk_score_tree(model, tips, w, ops, n_nodes, ra, rb, rt)["lnl"]) <= 1e-9


def test_group_branch_length_loop(eng, group3):
    model = aa_model(4)
    tr, ops, ra, rb, rt, n_nodes, tips = setup_lk(10, 5000, model, seed=3)
    for h in (eng, group3):
        h.lk_set_model(model)
        h.lk_set_tips(tips, capacity=n_nodes)
        h.lk_score_tree(ops, ra, rb, rt)
    ts = [0.01, 0.1, 0.7]
    a, b = eng.lk_edge_lnl(ra, rb, ts), group3.lk_edge_lnl(ra, rb, ts)
    assert np.abs(a - b).max() <= 1e-12 * np.abs(a).max()
    t1, l1, _ = eng.lk_optimize_branch(ra, rb, t0=rt, tol=1e-10)
    t3, l3, it = group3.lk_optimize_branch(ra, rb, t0=rt, tol=1e-10)
    assert it <= 30 and abs(t1 - t3) <= 1e-7 * max(t1, 1e-3) and rel_err(l3, l1) <= 1e-12


@pytest.mark.parametrize("T,N,dtype,ns", [(16, 5000, np.uint8, 4), (33, 70001, np.uint8, 4), (9, 2500, np.uint16, 11)])
def test_group_fitch_is_exact(eng, group3, oracle, T, N, dtype, ns):
    tr = tree.random_tree(T, 4)
    ops, ra, rb, rt, n_nodes = tree.schedule(tr)
    chars = tree.random_fitch_chars(T, N, ns, seed=8, dtype=dtype)
    w = np.random.default_rng(2).integers(0, 4, N).astype(float)
    for weights in (None, w):
        group3.fitch_set_tips(chars, ns, weights=weights, capacity=n_nodes)
        want = oracle.fitch_score_tree(chars, weights, ops, n_nodes, ra, rb, want_sets=True)
        assert group3.fitch_score_tree(ops, ra, rb) == want["length"]
        node = int(ops[-1]["parent"])
        assert np.array_equal(group3.fitch_get_states(node), want["prelim"][node])
    group3.fitch_uppass(ops, ra, rb)
    fin = oracle.fitch_uppass(T, want["prelim"], ops, ra, rb)
    assert np.array_equal(group3.fitch_get_states(ra, final=True), fin[ra])


def test_group_reports_the_failing_shard(group3):
    model = dna_gtr_g4()
    tips = np.ones((4, 5000), dtype=np.uint8)
    tips[2, 4100] = 0  # lands in the last shard
    group3.lk_set_model(model)
    with pytest.raises(engine.PhyloError) as ei:
        group3.lk_set_tips(tips)
    assert ei.value.code == -4 and "shard" in str(ei.value)
    with pytest.raises(engine.PhyloError):
        group3.lk_score_tree(engine.make_ops([4], [0], [1]), 4, 2, 0.1)  # nothing loaded


def test_group_over_all_visible_gpus(oracle, built):
    """One engine per visible GPU (one device on the single-GPU box): bit-identical to a
    single engine on device 0, and to the 3-way split above by construction."""
    nd = _n_devices()
    model = dna_gtr_g4()
    tr, ops, ra, rb, rt, n_nodes, tips = setup_lk(32, 8192, model, seed=41)
    tips = np.ascontiguousarray(np.tile(tips, (1, 40)))
    g = engine.Group(list(range(nd)))
    e = engine.Engine(0)
    try:
        g.lk_set_model(model)
        g.lk_set_tips(tips, capacity=n_nodes)
        e.lk_set_model(model)
        e.lk_set_tips(tips, capacity=n_nodes)
        assert g.lk_score_tree(ops, ra, rb, rt) == e.lk_score_tree(ops, ra, rb, rt)
        assert g.launch_count >= nd
    finally:
        g.close()
        e.close()


@pytest.mark.parametrize("dtype,T,N", [(np.uint8, 24, 50000), (np.uint32, 9, 20011), (np.uint8, 5, 3000)])
def test_group_compression_equals_one_engine(eng, group3, dtype, T, N):
    """phylo_group_compress_patterns (slabs of sites on every device, tables merged on device 0) returns the same
    patterns (first-occurrence order), weights and site map as one engine on the whole alignment -- and as the
    numpy oracle; N = 3000 is below the sharding threshold (one engine does it)."""
    from oracle.oracle import compress_patterns

    rng = np.random.default_rng(N)
    distinct = rng.integers(1, 16, (T, 700)).astype(dtype)
    masks = np.ascontiguousarray(distinct[:, rng.integers(0, 700, N)])
    for weights in (None, rng.integers(1, 4, N).astype(float)):
        one = eng.compress_patterns(masks, weights)
        many = group3.compress_patterns(masks, weights)
        orc = compress_patterns(masks, weights)
        for a, b, c in zip(one, many, orc):
            assert np.array_equal(a, b) and np.array_equal(b, c)
