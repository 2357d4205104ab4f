"""3-directional CLVs (phylo_lk_uppass; Node.Make3D, lib/node.ml:363-477): after the pre-order pass every
branch is a root edge. Checks: up CLVs == the oracle running the same updates; lnL joined across EVERY
edge == the root-edge value (pulley principle, reversible models) and == the oracle; the branch-length
optimiser works on an arbitrary interior edge."""
import numpy as np
import pytest

from helpers import aa_model, codon_model, dna_gtr_g4, rel_err, setup_lk
from phylocaml_b200 import tree

pytestmark = pytest.mark.gpu


def _models():
    return [("dna_gtr_g4", dna_gtr_g4(), 14, 1500), ("dna_pinvar", dna_gtr_g4(pinvar=0.15), 10, 900),
            ("aa20", aa_model(4), 9, 300), ("codon61", codon_model(), 7, 120)]


@pytest.mark.parametrize("name,model,T,N", _models(), ids=[m[0] for m in _models()])
def test_every_edge_is_a_root_edge(eng, oracle, name, model, T, N):
    tr, ops, ra, rb, rt, n_nodes, tips = setup_lk(T, N, model, seed=31, mean_bl=0.12)
    w = np.random.default_rng(2).integers(1, 4, N).astype(float)
    up_slot, cap, up_ops, edges = tree.uppass_plan(ops, ra, rb, rt, n_nodes)
    eng.lk_set_model(model)
    eng.lk_set_tips(tips, weights=w, capacity=cap)
    lnl = eng.lk_score_tree(ops, ra, rb, rt)
    eng.lk_uppass(ops, ra, rb, rt, up_slot)
    all_ops = np.concatenate([ops, up_ops])
    want = oracle.lk_score_tree(model, tips, w, all_ops, cap, ra, rb, rt, want_clv=True)
    assert rel_err(lnl, want["lnl"]) <= 1e-9
    assert len(edges) == 2 * T - 4
    for v, u, tv in edges:
        clv, sc = eng.lk_get_clv(u)
        assert np.array_equal(sc, want["scale"][u]), (v, u)
        assert np.abs(clv - want["clv"][u]).max() <= 1e-12 * np.abs(want["clv"][u]).max()
        got = eng.lk_edge_lnl(v, u, [tv])[0]
        assert rel_err(got, lnl) <= 1e-11, (v, u, got, lnl)  # pulley principle
    batch = eng.lk_edge_lnl_batch([e[0] for e in edges], [e[1] for e in edges], [e[2] for e in edges])
    assert all(b == eng.lk_edge_lnl(v, u, [tv])[0] for b, (v, u, tv) in zip(batch[:5], edges[:5]))
    assert np.max(np.abs(batch - lnl)) <= 1e-11 * abs(lnl)
    # oracle re-rooted on one interior edge: same schedule, root = that directional pair
    v, u, tv = next(e for e in edges if e[0] >= T)
    assert rel_err(oracle.lk_score_tree(model, tips, w, all_ops, cap, v, u, tv)["lnl"], lnl) <= 1e-11


def test_branch_optimiser_on_an_interior_edge(eng, oracle):
    from scipy.optimize import minimize_scalar

    model = dna_gtr_g4()
    T, N = 12, 3000
    tr, ops, ra, rb, rt, n_nodes, tips = setup_lk(T, N, model, seed=5, mean_bl=0.2)
    up_slot, cap, up_ops, edges = tree.uppass_plan(ops, ra, rb, rt, n_nodes)
    eng.lk_set_model(model)
    eng.lk_set_tips(tips, capacity=cap)
    eng.lk_score_tree(ops, ra, rb, rt)
    eng.lk_uppass(ops, ra, rb, rt, up_slot)
    for v, u, tv in (edges[3], next(e for e in edges if e[0] < T)):
        t_opt, l_opt, iters = eng.lk_optimize_branch(v, u, t0=tv, t_min=1e-8, t_max=50.0, tol=1e-10)
        res = minimize_scalar(lambda t: -eng.lk_edge_lnl(v, u, [t])[0], bounds=(1e-8, 50.0), method="bounded",
                              options={"xatol": 1e-10})
        assert abs(t_opt - res.x) <= 1e-5 * max(1.0, res.x)
        assert l_opt >= -res.fun - 1e-7 * abs(res.fun)


def test_uppass_argument_errors(eng):
    model = dna_gtr_g4()
    tr, ops, ra, rb, rt, n_nodes, tips = setup_lk(8, 200, model, seed=1)
    up_slot, cap, up_ops, edges = tree.uppass_plan(ops, ra, rb, rt, n_nodes)
    eng.lk_set_model(model)
    eng.lk_set_tips(tips, capacity=cap)
    eng.lk_score_tree(ops, ra, rb, rt)
    bad = up_slot.copy()
    bad[int(ops[0]["left"])] = int(ops[-1]["parent"])  # a slot the schedule itself uses
    with pytest.raises(RuntimeError):
        eng.lk_uppass(ops, ra, rb, rt, bad)
    # a child wanted while its parent is skipped
    child = next(int(o["left"]) for o in ops if int(o["parent"]) not in (ra, rb))
    parent = next(int(o["parent"]) for o in ops if int(o["left"]) == child)
    bad = up_slot.copy()
    bad[parent] = -1
    with pytest.raises(RuntimeError):
        eng.lk_uppass(ops, ra, rb, rt, bad)


@pytest.mark.parametrize("K,N", [(1, 777), (2, 3000), (4, 1024 * 1024 + 5), (8, 2100)])
def test_batched_joins_equal_single_joins_bitwise(eng, oracle, K, N):
    """phylo_lk_edge_lnl_batch for 4 states is one launch over all (edge, 1024-pattern block) work items; every
    edge's lnL must be the single-join value bit for bit (same arithmetic, same canonical fold), for every rate-class
    count the batch kernel is instantiated for, tip and CLV operands, ragged N, and more than one fold level
    (N > 1024 * 1024 patterns: the level-1 partials of one edge no longer fit one block)."""
    from helpers import GTR_CO, GTR_PI
    from phylocaml_b200 import mlmodel
    model = mlmodel.create(("GTR", GTR_CO), 4, pi=GTR_PI, site_var=("gamma", K, 0.5) if K > 1 else None)
    T = 6 if N > 100000 else 11
    tr, ops, ra, rb, rt, n_nodes, tips = setup_lk(T, N, model, seed=13, mean_bl=0.2)
    up_slot, cap, up_ops, edges = tree.uppass_plan(ops, ra, rb, rt, n_nodes)
    w = np.random.default_rng(4).integers(1, 3, N).astype(float)
    eng.lk_set_model(model)
    eng.lk_set_tips(tips, weights=w, capacity=cap)
    lnl = eng.lk_score_tree(ops, ra, rb, rt)
    eng.lk_uppass(ops, ra, rb, rt, up_slot)
    ea, eb = [e[0] for e in edges], [e[1] for e in edges]
    ts = [e[2] * (1.0 + 0.1 * i) for i, e in enumerate(edges)]  # not the tree's own lengths: every edge a different P(t)
    batch = eng.lk_edge_lnl_batch(ea, eb, ts)
    for i, (v, u, _) in enumerate(edges):
        assert batch[i] == eng.lk_edge_lnl(v, u, [ts[i]])[0], (K, i)
    own = eng.lk_edge_lnl_batch(ea, eb, [e[2] for e in edges])
    assert np.max(np.abs(own - lnl)) <= 1e-11 * abs(lnl)


@pytest.mark.parametrize("K,N,T", [(4, 2500, 40), (1, 1025, 17), (8, 700, 12), (2, 90000, 24)])
def test_level_batched_uppass_equals_per_update_launches_bitwise(eng, monkeypatch, K, N, T):
    """4 states: one launch per tree level (prune4_level_kernel) against one launch per update
    (PHYLO_UPPASS_BATCH=0): every directional CLV and scale counter bit for bit, fewer launches."""
    from phylocaml_b200 import mlmodel
    from helpers import GTR_CO, GTR_PI

    sv = ("gamma", K, 0.5) if K > 1 else None
    model = mlmodel.create(("GTR", list(GTR_CO)), 4, pi=list(GTR_PI), site_var=sv)
    tr, ops, ra, rb, rt, n_nodes, tips = setup_lk(T, N, model, seed=K + T, mean_bl=0.4)
    up_slot, cap, up_ops, edges = tree.uppass_plan(ops, ra, rb, rt, n_nodes)
    eng.lk_set_model(model)
    eng.lk_set_tips(tips, capacity=cap)
    lnl = eng.lk_score_tree(ops, ra, rb, rt)
    monkeypatch.setenv("PHYLO_UPPASS_BATCH", "0")
    l0 = eng.launch_count
    eng.lk_uppass(ops, ra, rb, rt, up_slot)
    per_update = eng.launch_count - l0
    want = {u: eng.lk_get_clv(u) for v, u, tv in edges}
    # a fresh down pass invalidates nothing the up pass needs, but rewrite the up slots from scratch anyway
    monkeypatch.setenv("PHYLO_UPPASS_BATCH", "1")
    eng.lk_set_tips(tips, capacity=cap)
    assert eng.lk_score_tree(ops, ra, rb, rt) == lnl
    l0 = eng.launch_count
    eng.lk_uppass(ops, ra, rb, rt, up_slot)
    batched = eng.launch_count - l0
    assert batched < per_update, (batched, per_update)
    for v, u, tv in edges:
        clv, sc = eng.lk_get_clv(u)
        assert np.array_equal(sc, want[u][1]), (v, u)
        assert np.array_equal(clv, want[u][0]), (v, u)
    joins = eng.lk_edge_lnl_batch([e[0] for e in edges], [e[1] for e in edges], [e[2] for e in edges])
    assert np.max(np.abs(joins - lnl)) <= 1e-11 * abs(lnl)
