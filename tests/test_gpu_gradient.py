"""Model-parameter derivatives (phylo_lk_param_gradient; the reference's gen_subst/rates/prior_opt_func are
`failwith "todo"`, lib/mlModel.ml:822-829): d lnL / d theta from the retained 3-directional CLVs against
central differences of full re-evaluations with the perturbed model."""
import numpy as np
import pytest

from helpers import GTR_CO, GTR_PI, setup_lk
from phylocaml_b200 import mlmodel, tree

pytestmark = pytest.mark.gpu


def richardson(f, h):
    """central difference with the h^2 term removed: (4 D(h/2) - D(h)) / 3"""
    d1 = (f(h) - f(-h)) / (2 * h)
    d2 = (f(h / 2) - f(-h / 2)) / h
    return (4 * d2 - d1) / 3


def _gtr(co, alpha, pinvar=None, pi=GTR_PI):
    sv = ("gamma", 4, alpha) if pinvar is None else ("theta", 4, alpha, pinvar)
    return mlmodel.create(("GTR", list(co)), 4, pi=list(pi), site_var=sv)


def _prepare(eng, model, T, N, seed):
    tr, ops, ra, rb, rt, n_nodes, tips = setup_lk(T, N, model, seed=seed, mean_bl=0.15)
    up_slot, cap, up_ops, edges = tree.uppass_plan(ops, ra, rb, rt, n_nodes)
    w = np.random.default_rng(seed).integers(1, 5, N).astype(float)
    eng.lk_set_model(model)
    eng.lk_set_tips(tips, weights=w, capacity=cap)
    return ops, ra, rb, rt, up_slot, tips, w, cap


@pytest.mark.parametrize("pinvar", [None, 0.12])
def test_gradient_gtr_exchangeabilities_and_gamma_shape(eng, pinvar):
    co, alpha = np.array(GTR_CO), 0.5
    model = _gtr(co, alpha, pinvar)
    ops, ra, rb, rt, up_slot, tips, w, cap = _prepare(eng, model, 11, 3000, seed=6)
    h = 1e-6
    dQ = np.zeros((6, 4, 4))
    drates = np.zeros((6, 4))
    for p in range(5):
        e = np.zeros(5)
        e[p] = h
        dQ[p] = (_gtr(co + e, alpha, pinvar)["Q"] - _gtr(co - e, alpha, pinvar)["Q"]) / (2 * h)
    drates[5] = (_gtr(co, alpha + h, pinvar)["rates"] - _gtr(co, alpha - h, pinvar)["rates"]) / (2 * h)
    lnl = eng.lk_score_tree(ops, ra, rb, rt)
    eng.lk_uppass(ops, ra, rb, rt, up_slot)
    grad = eng.lk_param_gradient(ops, ra, rb, rt, up_slot, dQ=dQ, drates=drates)

    def f(co_, alpha_):
        eng.lk_set_model(_gtr(co_, alpha_, pinvar))
        eng.lk_set_tips(tips, weights=w, capacity=cap)
        return eng.lk_score_tree(ops, ra, rb, rt)

    for p in range(6):
        e = np.zeros(5)
        if p < 5:
            e[p] = 1.0
            fd = richardson(lambda x: f(co + x * e, alpha), 2e-3)
        else:
            fd = richardson(lambda x: f(co, alpha + x), 2e-3)
        assert abs(grad[p] - fd) <= 2e-6 * max(1.0, abs(fd)), (p, grad[p], fd, lnl)


def test_gradient_of_the_priors(eng):
    """d lnL / d pi along a direction that keeps sum(pi) = 1: Q moves with pi (reversible GTR), and the root
    distribution moves too."""
    co, alpha = np.array(GTR_CO), 0.7
    pi = np.array(GTR_PI)
    d = np.array([1.0, -1.0, 0.5, -0.5])
    model = _gtr(co, alpha, pi=pi)
    ops, ra, rb, rt, up_slot, tips, w, cap = _prepare(eng, model, 9, 2000, seed=8)
    h = 1e-6
    mp, mm = _gtr(co, alpha, pi=pi + h * d), _gtr(co, alpha, pi=pi - h * d)
    dQ = ((mp["Q"] - mm["Q"]) / (2 * h))[None]
    dpi = ((mp["pi"] - mm["pi"]) / (2 * h))[None]
    eng.lk_score_tree(ops, ra, rb, rt)
    eng.lk_uppass(ops, ra, rb, rt, up_slot)
    grad = eng.lk_param_gradient(ops, ra, rb, rt, up_slot, dQ=dQ, dpi=dpi)

    def f(pi_):
        eng.lk_set_model(_gtr(co, alpha, pi=pi_))
        eng.lk_set_tips(tips, weights=w, capacity=cap)
        return eng.lk_score_tree(ops, ra, rb, rt)

    fd = richardson(lambda x: f(pi + x * d), 1e-3)
    assert abs(grad[0] - fd) <= 2e-6 * max(1.0, abs(fd)), (grad[0], fd)


def test_gradient_20_states_generic_kernel(eng):
    """20 states through the run-time-S kernel: derivative along a symmetric exchangeability direction."""
    R, pi = mlmodel.synthetic_reversible(20, 4)
    rng = np.random.default_rng(3)
    dR = rng.random((20, 20))
    dR = dR + dR.T
    np.fill_diagonal(dR, 0.0)

    def mk(eps):
        return mlmodel.create(("Const", R + eps * dR), 20, pi=pi, site_var=("gamma", 2, 0.7))

    model = mk(0.0)
    ops, ra, rb, rt, up_slot, tips, w, cap = _prepare(eng, model, 7, 300, seed=3)
    h = 1e-6
    dQ = ((mk(h)["Q"] - mk(-h)["Q"]) / (2 * h))[None]
    eng.lk_score_tree(ops, ra, rb, rt)
    eng.lk_uppass(ops, ra, rb, rt, up_slot)
    grad = eng.lk_param_gradient(ops, ra, rb, rt, up_slot, dQ=dQ)

    def f(eps):
        eng.lk_set_model(mk(eps))
        eng.lk_set_tips(tips, weights=w, capacity=cap)
        return eng.lk_score_tree(ops, ra, rb, rt)

    fd = richardson(f, 2e-4)
    assert abs(grad[0] - fd) <= 2e-6 * max(1.0, abs(fd)), (grad[0], fd)


@pytest.mark.parametrize("K,pinvar,N", [(4, None, 2999), (4, 0.2, 1025), (1, None, 700), (2, None, 1024), (8, None, 333)])
def test_tensor_core_gradient_equals_the_scalar_kernel(eng, monkeypatch, K, pinvar, N):
    """4 states: param_grad4_mma_kernel (every branch in one launch, DMMA, seven parameters per pass) against the
    scalar per-branch kernel (PHYLO_GRAD_MMA=0) -- nine parameters (two passes) mixing rate-matrix, rate and
    prior directions, weights, ragged pattern counts, tip and interior branches. Only the summation order inside
    a DMMA differs."""
    rng = np.random.default_rng(K * 100 + N)
    sv = ("gamma", K, 0.6) if pinvar is None else ("theta", K, 0.6, pinvar)
    model = mlmodel.create(("GTR", list(GTR_CO)), 4, pi=list(GTR_PI), site_var=sv) if K > 1 else \
        mlmodel.create(("GTR", list(GTR_CO)), 4, pi=list(GTR_PI))
    ops, ra, rb, rt, up_slot, tips, w, cap = _prepare(eng, model, 13, N, seed=K + N)
    n_params = 9
    dQ = rng.standard_normal((n_params, 4, 4))
    dQ -= dQ.sum(axis=2, keepdims=True) * np.eye(4)  # rows of a rate-matrix direction sum to zero
    drates = rng.standard_normal((n_params, K))
    dpi = None
    if pinvar is None:
        dpi = rng.standard_normal((n_params, 4))
        dpi -= dpi.mean(axis=1, keepdims=True)
    eng.lk_score_tree(ops, ra, rb, rt)
    eng.lk_uppass(ops, ra, rb, rt, up_slot)
    monkeypatch.setenv("PHYLO_GRAD_MMA", "0")
    want = eng.lk_param_gradient(ops, ra, rb, rt, up_slot, dQ=dQ, drates=drates, dpi=dpi)
    monkeypatch.setenv("PHYLO_GRAD_MMA", "1")
    launches = eng.launch_count
    got = eng.lk_param_gradient(ops, ra, rb, rt, up_slot, dQ=dQ, drates=drates, dpi=dpi)
    assert eng.launch_count - launches == 4  # two passes: gradient kernel + branch sum each
    assert np.all(np.isfinite(got))
    scale = np.maximum(1.0, np.abs(want))
    assert np.max(np.abs(got - want) / scale) <= 1e-10, (got, want)
    # branches in several launches (the per-branch block results are capped at 256 MB; here: three branches per launch)
    nblocks = (N + 1023) // 1024
    monkeypatch.setenv("PHYLO_GRAD_CHUNK_BYTES", str(3 * 8 * 7 * nblocks))
    l0 = eng.launch_count
    chunked = eng.lk_param_gradient(ops, ra, rb, rt, up_slot, dQ=dQ, drates=drates, dpi=dpi)
    assert eng.launch_count - l0 == 2 * 2 * ((2 * 13 - 3 + 2) // 3)
    assert np.array_equal(chunked, got)
    monkeypatch.delenv("PHYLO_GRAD_CHUNK_BYTES")
    # deterministic: the same call returns the same bits
    assert np.array_equal(got, eng.lk_param_gradient(ops, ra, rb, rt, up_slot, dQ=dQ, drates=drates, dpi=dpi))
