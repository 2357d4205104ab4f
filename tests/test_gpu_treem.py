"""Tree-fused 20 / 61-state pruning (lk_treem_kernel.cuh) against the oracle and against the
per-node DMMA path: CLVs of every interior node, scale counters (bit-exact), site lnL, lnL; random and
caterpillar trees, ragged pattern counts, ambiguous / missing tips, weights, invariant sites,
per-site rescaling, re-scoring from CLVs already resident (STORED operands)."""
import numpy as np
import pytest

from helpers import aa_model, codon_model, rel_err, setup_lk
from phylocaml_b200 import tree

pytestmark = pytest.mark.gpu
LNL_RTOL = 1e-9
CLV_RTOL = 1e-12


def _models():
    return [("aa20_k4", aa_model(4)), ("aa20_k1", aa_model(1)), ("aa20_k2", aa_model(2)), ("codon61", codon_model())]


def _both_paths(eng, ops, ra, rb, rt, nodes):
    out = {}
    for fused in (1, 0):
        eng.set_option(eng.OPT_FUSED_TREE, fused)
        lnl = eng.lk_score_tree(ops, ra, rb, rt)
        site = eng.lk_get_site_lnl()
        clvs = {int(n): eng.lk_get_clv(int(n)) for n in nodes}
        out[fused] = (lnl, site, clvs)
    eng.set_option(eng.OPT_FUSED_TREE, 1)
    return out


@pytest.mark.parametrize("name,model", _models(), ids=[m[0] for m in _models()])
@pytest.mark.parametrize("T,N,kind", [(9, 1, "random"), (12, 777, "random"), (24, 2049, "random"), (17, 300, "caterpillar")])
def test_treem_equals_oracle_and_per_node(eng, oracle, name, model, T, N, kind):
    tr, ops, ra, rb, rt, n_nodes, tips = setup_lk(T, N, model, seed=T + N, tree_kind=kind, mean_bl=0.15, missing=0.05)
    w = np.random.default_rng(5).integers(1, 7, N).astype(float)
    eng.lk_set_model(model)
    eng.lk_set_tips(tips, weights=w, capacity=n_nodes)
    launches0 = eng.launch_count
    nodes = [int(o["parent"]) for o in ops]
    got = _both_paths(eng, ops, ra, rb, rt, nodes)
    want = oracle.lk_score_tree(model, tips, w, ops, n_nodes, ra, rb, rt, want_clv=True)
    for fused in (1, 0):
        lnl, site, clvs = got[fused]
        assert rel_err(lnl, want["lnl"]) <= LNL_RTOL, (fused, lnl, want["lnl"])
        assert np.abs(site - want["site_lnl"]).max() <= 1e-9 * np.abs(want["site_lnl"]).max()
        for n in nodes:
            clv, sc = clvs[n]
            assert np.array_equal(sc, want["scale"][n]), (fused, n)
            assert np.abs(clv - want["clv"][n]).max() <= CLV_RTOL * np.abs(want["clv"][n]).max(), (fused, n)
    # the fused evaluation is a handful of launches, the per-node one at least one per median
    assert rel_err(got[1][0], got[0][0]) <= 1e-13
    assert eng.launch_count - launches0 >= len(ops)


@pytest.mark.parametrize("name,model", [("aa20_k4", aa_model(4)), ("codon61", codon_model())], ids=["aa20_k4", "codon61"])
def test_treem_rescaling_on_a_deep_tree(eng, oracle, name, model):
    """Long caterpillar with long branches: site maxima fall below 2^-256 several times."""
    T, N = (260, 200) if model["S"] == 20 else (160, 96)
    tr, ops, ra, rb, rt, n_nodes, tips = setup_lk(T, N, model, seed=3, tree_kind="caterpillar", mean_bl=1.2)
    eng.lk_set_model(model)
    eng.lk_set_tips(tips, capacity=n_nodes)
    lnl = eng.lk_score_tree(ops, ra, rb, rt)
    want = oracle.lk_score_tree(model, tips, None, ops, n_nodes, ra, rb, rt, want_clv=True)
    assert max(want["scale"][ra].max(), want["scale"][rb].max()) >= 1, "test must trigger rescaling"
    assert rel_err(lnl, want["lnl"]) <= LNL_RTOL
    for n in (int(ops[len(ops) // 2]["parent"]), int(ops[-1]["parent"])):
        clv, sc = eng.lk_get_clv(n)
        assert np.array_equal(sc, want["scale"][n])
        assert np.abs(clv - want["clv"][n]).max() <= CLV_RTOL * np.abs(want["clv"][n]).max()


def test_treem_pinvar_and_all_ambiguous_tip(eng, oracle):
    from phylocaml_b200 import mlmodel

    R, pi = mlmodel.synthetic_reversible(20, 4)
    model = mlmodel.create(("Const", R), 20, pi=pi, site_var=("theta", 4, 0.7, 0.15))
    tr, ops, ra, rb, rt, n_nodes, tips = setup_lk(14, 500, model, seed=9, mean_bl=0.2)
    tips[3, :] = (1 << 20) - 1          # a taxon that is missing everywhere
    tips[5, ::3] |= 0b1010101           # ambiguity codes
    tips[:, :16] = tips[0, :16][None, :]  # constant sites: the invariant-sites term is exercised
    eng.lk_set_model(model)
    eng.lk_set_tips(tips, capacity=n_nodes)
    lnl = eng.lk_score_tree(ops, ra, rb, rt)
    want = oracle.lk_score_tree(model, tips, None, ops, n_nodes, ra, rb, rt)
    assert rel_err(lnl, want["lnl"]) <= LNL_RTOL
    site = eng.lk_get_site_lnl()
    assert np.abs(site - want["site_lnl"]).max() <= 1e-9 * np.abs(want["site_lnl"]).max()


@pytest.mark.parametrize("name,model", [("aa20_k4", aa_model(4)), ("codon61", codon_model())], ids=["aa20_k4", "codon61"])
def test_treem_rescoring_from_resident_clvs(eng, oracle, name, model):
    """A partial schedule (the path from a changed branch to the root) whose other operands are CLVs left
    in their node slots by the previous call: the STORED operand kind of the fused kernel."""
    T, N = 20, 400
    tr, ops, ra, rb, rt, n_nodes, tips = setup_lk(T, N, model, seed=12, mean_bl=0.1)
    eng.lk_set_model(model)
    eng.lk_set_tips(tips, capacity=n_nodes)
    eng.lk_score_tree(ops, ra, rb, rt)
    ops2 = ops.copy()
    first = 0
    ops2["t_left"][first] *= 3.0
    # ops on the path from ops[first].parent to the root, in schedule order
    dirty = {int(ops2["parent"][first])}
    keep = [first]
    for i in range(first + 1, len(ops2)):
        if int(ops2["left"][i]) in dirty or int(ops2["right"][i]) in dirty:
            dirty.add(int(ops2["parent"][i]))
            keep.append(i)
    part = ops2[keep]
    assert len(part) < len(ops2)
    lnl = eng.lk_score_tree(part, ra, rb, rt)
    want = oracle.lk_score_tree(model, tips, None, ops2, n_nodes, ra, rb, rt)
    assert rel_err(lnl, want["lnl"]) <= LNL_RTOL
