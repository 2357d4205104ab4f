"""BASELINE.json's configurations at their FULL sizes, checked through properties that do not
need the oracle to run at that size:

* the alignment is a short base block tiled to the full length, so every per-pattern result
  (site lnL, CLV rows, scale counters, Fitch sets) must repeat with the block's period, bit for
  bit, wherever the pattern sits in the grid;
* the first block is compared with the CPU oracle (lnL <= 1e-9 relative, sets bit-exact);
* totals: lnL == the sum of the site lnL (1e-12), Fitch length == full-size oracle (the C port
  finishes 63 M char-ops in well under a second), block partials of two shards reduce to the
  unsharded value bit for bit (what the N-rank bench relies on).

Each test owns its engine and closes it, so the 130 GB CLV arena of config 3 is returned
before the next test."""
import math

import numpy as np
import pytest

from helpers import aa_model, codon_model, dna_gtr_g4, mask_dtype, rel_err
from phylocaml_b200 import engine, tree

pytestmark = pytest.mark.gpu

LNL_RTOL = 1e-9  # BASELINE.json north_star: log-likelihoods <= 1e-9 relative in fp64
BLOCK = 8192     # base block (patterns) that is tiled to the full size
SAMPLE = 2048    # patterns of the first block compared with the oracle


@pytest.fixture()
def own_eng(built):
    e = engine.Engine(0)
    yield e
    e.close()


def _tiled(base, N):
    reps = (N + base.shape[1] - 1) // base.shape[1]
    return np.ascontiguousarray(np.tile(base, (1, reps))[:, :N])


def _periodic(a, period):
    """a[i] == a[i % period] for every i (leading axis), bit for bit."""
    n = a.shape[0]
    full = (n // period) * period
    body = a[:full].reshape((n // period, period) + a.shape[1:])
    return bool((body == body[0]).all()) and bool((a[full:] == a[:n - full]).all())


def _lk_full(own_eng, oracle, model, T, N, mean_bl=0.1, retain=1, fused=1):
    tr = tree.random_tree(T, 1, mean_bl=mean_bl)
    ops, ra, rb, rt, n_nodes = tree.schedule(tr)
    base = tree.evolve_tips(tr, model, BLOCK, 3, dtype=mask_dtype(model["S"]))
    tips = _tiled(base, N)
    own_eng.set_option(own_eng.OPT_FUSED_TREE, fused)
    own_eng.set_option(own_eng.OPT_RETAIN_CLV, retain)
    own_eng.lk_set_model(model)
    own_eng.lk_set_tips(tips, capacity=n_nodes)
    lnl = own_eng.lk_score_tree(ops, ra, rb, rt)
    site = own_eng.lk_get_site_lnl()
    assert site.shape == (N,) and np.isfinite(site).all()
    assert _periodic(site, BLOCK)
    assert rel_err(lnl, math.fsum(site)) <= 1e-12
    want = oracle.lk_score_tree(model, np.ascontiguousarray(base[:, :SAMPLE]), None, ops, n_nodes, ra, rb, rt,
                                want_clv=bool(retain))
    assert np.abs(site[:SAMPLE] - want["site_lnl"]).max() <= LNL_RTOL * np.abs(want["site_lnl"]).max()
    return tr, ops, ra, rb, rt, n_nodes, base, tips, lnl, site, want


def _check_clv(own_eng, node, want, S, K):
    clv, sc = own_eng.lk_get_clv(node)
    clv = clv.reshape(-1, K * S)
    assert _periodic(clv, BLOCK) and _periodic(sc, BLOCK)
    w = want["clv"][node].reshape(-1, K * S)
    assert np.array_equal(sc[:SAMPLE], want["scale"][node])
    assert np.abs(clv[:SAMPLE] - w).max() <= 1e-11 * np.abs(w).max()


def test_cfg2_fitch_64_taxa_1M_characters_all_kernels(own_eng, oracle):
    """BASELINE config 2 at full size, every whole-tree kernel: length, per-node costs and the
    root-side sets bit-exact against the oracle on all 1 M characters."""
    T, N = 64, 1_000_000
    tr = tree.random_tree(T, 1)
    ops, ra, rb, rt, n_nodes = tree.schedule(tr)
    chars = tree.random_fitch_chars(T, N, 4, seed=5)
    want = oracle.fitch_score_tree(chars, None, ops, n_nodes, ra, rb, want_sets=True)
    for walk in (1, 3, 2, 0):
        own_eng.set_option(own_eng.OPT_FITCH_WALK, walk)
        own_eng.fitch_set_tips(chars, 4, capacity=n_nodes)
        assert own_eng.fitch_score_tree(ops, ra, rb) == want["length"], walk
        costs = own_eng.fitch_get_node_costs()
        for op in ops:
            assert costs[int(op["parent"])] == want["node_cost"][int(op["parent"])], walk
        for node in (int(ops[0]["parent"]), int(ops[len(ops) // 2]["parent"]), int(ops[-1]["parent"])):
            assert np.array_equal(own_eng.fitch_get_states(node), want["prelim"][node]), (walk, node)
    # up-pass at full size: final sets of the two root-edge ends against the oracle's rule
    own_eng.fitch_uppass(ops, ra, rb)
    fin = oracle.fitch_uppass(T, want["prelim"], ops, ra, rb)
    for node in (ra, rb, int(ops[0]["parent"])):
        assert np.array_equal(own_eng.fitch_get_states(node, final=True), fin[node]), node


def test_cfg3_dna_256_taxa_4M_patterns_lnl_only_and_shards(own_eng, oracle):
    """BASELINE config 3 at full size, lnL-only tree-fused kernel (no CLV written): periodic
    site lnL, oracle on the first block, and two 1024-aligned shards whose gathered block
    partials reduce to the unsharded lnL bit for bit."""
    model = dna_gtr_g4()
    T, N = 256, 4_000_000
    tr, ops, ra, rb, rt, n_nodes, base, tips, lnl, site, _ = _lk_full(own_eng, oracle, model, T, N, retain=0)
    whole_parts = own_eng.lk_get_block_partials()
    assert own_eng.reduce_partials(whole_parts) == lnl
    cut = 1024 * 1953  # ~ N/2, 1024-aligned like bench.py's slabs
    parts = []
    for lo, hi in ((0, cut), (cut, N)):
        own_eng.lk_set_tips(np.ascontiguousarray(tips[:, lo:hi]), capacity=n_nodes)
        own_eng.lk_score_tree(ops, ra, rb, rt)
        parts.append(own_eng.lk_get_block_partials())
    assert own_eng.reduce_partials(np.concatenate(parts)) == lnl


def test_cfg3_dna_256_taxa_4M_patterns_retained_clvs(own_eng, oracle):
    """Config 3 with every interior CLV kept in HBM (254 x 512 MB): same lnL as lnL-only mode,
    CLVs and scale counters periodic and equal to the oracle on the first block -- for the first
    node written, a mid-tree node and the last one."""
    model = dna_gtr_g4()
    T, N = 256, 4_000_000
    tr, ops, ra, rb, rt, n_nodes, base, tips, lnl, site, want = _lk_full(own_eng, oracle, model, T, N, retain=1)
    for op in (ops[0], ops[len(ops) // 2], ops[-1]):
        _check_clv(own_eng, int(op["parent"]), want, 4, 4)
    # the branch-length entry points run off the retained CLVs
    got = own_eng.lk_edge_lnl(ra, rb, [rt])[0]
    assert rel_err(got, lnl) <= 1e-12
    own_eng.set_option(own_eng.OPT_RETAIN_CLV, 0)
    assert own_eng.lk_score_tree(ops, ra, rb, rt) == lnl


def test_cfg4_aa_128_taxa_500k_patterns(own_eng, oracle):
    """BASELINE config 4 at full size (20 states, K = 4, fp64 tensor-core kernels)."""
    model = aa_model(4)
    T, N = 128, 500_000
    tr, ops, ra, rb, rt, n_nodes, base, tips, lnl, site, want = _lk_full(own_eng, oracle, model, T, N)
    for op in (ops[0], ops[-1]):
        _check_clv(own_eng, int(op["parent"]), want, 20, 4)


def test_cfg5_codon_64_taxa_200k_patterns_and_branch_loop(own_eng, oracle):
    """BASELINE config 5 at full size (61 states, per-site scaling) with its branch-length
    loop: sum-table evaluations == P-matrix edge_lnl at the same lengths, and the Newton optimum
    is a stationary point with negative curvature."""
    model = codon_model()
    T, N = 64, 200_000
    tr, ops, ra, rb, rt, n_nodes, base, tips, lnl, site, want = _lk_full(own_eng, oracle, model, T, N,
                                                                          mean_bl=0.05)
    _check_clv(own_eng, int(ops[-1]["parent"]), want, 61, 1)
    ts = [0.5 * rt, rt, 2.0 * rt, 0.3]
    direct = own_eng.lk_edge_lnl(ra, rb, ts)
    assert rel_err(direct[1], lnl) <= 1e-12
    own_eng.lk_edge_prepare(ra, rb)
    val, d1, d2 = own_eng.lk_edge_eval(ts)
    for a, b in zip(val, direct):
        assert rel_err(a, b) <= LNL_RTOL
    t_opt, l_opt, iters = own_eng.lk_optimize_branch(ra, rb, t0=rt)
    v, g, h = own_eng.lk_edge_eval([t_opt])
    assert iters <= 30 and l_opt >= max(direct) - 1e-9 * abs(l_opt)
    assert rel_err(v[0], l_opt) <= 1e-12
    assert h[0] < 0 and abs(g[0]) <= 1e-5 * abs(h[0]) * max(t_opt, 1e-3) + 1e-6
