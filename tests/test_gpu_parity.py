"""GPU parity tests: the CUDA engine (through the C ABI) against the CPU oracle on the same
seeded inputs. Tolerances: Fitch sets/lengths and scale counters bit-exact; lnL <= 1e-9
relative (BASELINE.json north_star); CLV entries <= 1e-12 relative; P(t) <= 1e-12 absolute."""
import numpy as np
import pytest

from helpers import aa_model, codon_model, dna_gtr_g4, mask_dtype, rel_err, setup_lk
from phylocaml_b200 import engine, mlmodel, tree

pytestmark = pytest.mark.gpu

LNL_RTOL = 1e-9
CLV_RTOL = 1e-12
PT_ATOL = 1e-12


# ------------------------------------------------------------------------- P(t) ----
@pytest.mark.parametrize("t", [0.1, 0.5, 2.0, 100.0, -1.0, 0.0, 1e-11, 2.2250738585072014e-308])
def test_compose_gtr_matches_oracle(eng, oracle, t):
    m = dna_gtr_g4()
    P = eng.compose(m["U"], m["D"], m["Ui"], t)
    assert np.abs(P - oracle.compose(m["U"], m["D"], m["Ui"], t)).max() <= PT_ATOL


@pytest.mark.parametrize("S,t", [(20, 0.3), (61, 0.05), (61, 1.7)])
def test_compose_large_alphabets(eng, oracle, S, t):
    m = aa_model(1) if S == 20 else codon_model()
    P = eng.compose(m["U"], m["D"], m["Ui"], t)
    assert np.abs(P - oracle.compose(m["U"], m["D"], m["Ui"], t)).max() <= PT_ATOL
    assert np.abs(P.sum(1) - 1).max() < 1e-12


@pytest.mark.parametrize("t", [0.1, 0.37, -1.0, 0.0])
def test_compose_sym_float_quirk(eng, oracle, t):
    # lib/mlmodel.c:280: compose_sym takes `float t`
    m = mlmodel.create(("K2P", 0.4), 4)
    assert m["Ui"] is None
    P = eng.compose(m["U"], m["D"], None, t)
    assert np.abs(P - oracle.compose(m["U"], m["D"], None, t)).max() <= PT_ATOL


# ------------------------------------------------------------------- likelihood ----
def _check_lk(eng, oracle, model, T, N, seed=1, weights=None, tree_kind="random", mean_bl=0.1,
              check_clv=True):
    tr, ops, ra, rb, rt, n_nodes, tips = setup_lk(T, N, model, seed, tree_kind, mean_bl)
    eng.lk_set_model(model)
    eng.lk_set_tips(tips, weights=weights, capacity=n_nodes)
    lnl = eng.lk_score_tree(ops, ra, rb, rt)
    want = oracle.lk_score_tree(model, tips, weights, ops, n_nodes, ra, rb, rt, want_clv=check_clv)
    assert np.isfinite(lnl)
    assert rel_err(lnl, want["lnl"]) <= LNL_RTOL, (lnl, want["lnl"])
    site = eng.lk_get_site_lnl()
    assert np.abs(site - want["site_lnl"]).max() <= 1e-9 * np.abs(want["site_lnl"]).max()
    if check_clv:
        for op in ops[[0, len(ops) // 2, len(ops) - 1]]:
            clv, sc = eng.lk_get_clv(int(op["parent"]))
            w = want["clv"][op["parent"]]
            assert np.array_equal(sc, want["scale"][op["parent"]])
            assert np.abs(clv - w).max() <= CLV_RTOL * np.abs(w).max()
    return lnl, want


def test_lk_dna_cfg1(eng, oracle):
    """BASELINE config 1: DNA GTR+G4, 16 taxa x 10k sites."""
    _check_lk(eng, oracle, dna_gtr_g4(), 16, 10000)


def test_lk_dna_ref_literal_rates(eng, oracle):
    """lib/mlModel.ml:93-99 literal Gamma rates: r0 = 0 => identity P for class 0."""
    m = dna_gtr_g4(rates="ref_literal")
    assert m["rates"][0] == 0.0
    _check_lk(eng, oracle, m, 12, 3000)


@pytest.mark.parametrize("N", [1, 31, 1024, 1025, 4097])
def test_lk_dna_ragged_sizes(eng, oracle, N):
    _check_lk(eng, oracle, dna_gtr_g4(), 8, N, seed=7)


@pytest.mark.parametrize("K", [1, 2, 8])
def test_lk_dna_other_rate_counts(eng, oracle, K):
    sv = ("gamma", K, 0.8) if K > 1 else None
    m = mlmodel.create(("HKY85", 2.0), 4, pi=[0.1, 0.2, 0.3, 0.4], site_var=sv)
    _check_lk(eng, oracle, m, 10, 2500, seed=3)


def test_lk_dna_k3_uses_generic_kernel(eng, oracle):
    m = mlmodel.create(("GTR", [1.0, 2.5, 0.8, 1.2, 3.0]), 4, pi=[0.3, 0.2, 0.25, 0.25],
                       site_var=("gamma", 3, 0.6))
    _check_lk(eng, oracle, m, 9, 1500, seed=5)


def test_lk_dna_weights(eng, oracle):
    rng = np.random.default_rng(11)
    _check_lk(eng, oracle, dna_gtr_g4(), 16, 5000, weights=rng.integers(1, 50, 5000).astype(float))


def test_lk_dna_pinvar(eng, oracle):
    _check_lk(eng, oracle, dna_gtr_g4(pinvar=0.2), 12, 4000)


def test_lk_dna_rescaling_deep_tree(eng, oracle):
    """A 400-taxon caterpillar drives site maxima below 2^-256: scale counters must match
    bit-for-bit and lnL stay finite."""
    lnl, want = _check_lk(eng, oracle, dna_gtr_g4(), 400, 600, tree_kind="caterpillar", mean_bl=0.6)
    assert want["scale"].max() >= 1, "test must actually trigger rescaling"


@pytest.mark.parametrize("fused", [1, 2, 0])
def test_lk_pinvar_on_a_deep_tree_stays_finite(eng, oracle, fused):
    """Invariant-sites class with summed scale counters >= 6 (900-taxon caterpillar): the site value
    is formed in the scaled domain (lnl_pinvar, common.cuh), so variable sites keep a finite lnL
    and the branch-length derivatives stay finite. Oracle = the same statement, itself pinned to
    unscaled 80-bit pruning (tests/test_oracle_cpu.py)."""
    model = dna_gtr_g4(pinvar=0.2)
    tr, ops, ra, rb, rt, n_nodes, tips = setup_lk(900, 96, model, seed=2, tree_kind="caterpillar", mean_bl=0.9)
    tips[:, :8] = tips[0, :8][None, :]
    eng.set_option(eng.OPT_FUSED_TREE, fused)
    try:
        eng.lk_set_model(model)
        eng.lk_set_tips(tips, capacity=n_nodes)
        lnl = eng.lk_score_tree(ops, ra, rb, rt)
        want = oracle.lk_score_tree(model, tips, None, ops, n_nodes, ra, rb, rt, want_clv=True)
        assert (want["scale"][ra] + want["scale"][rb]).max() >= 6
        site = eng.lk_get_site_lnl()
        assert np.all(np.isfinite(site)) and np.isfinite(lnl)
        assert np.abs(site - want["site_lnl"]).max() <= 1e-9 * np.abs(want["site_lnl"]).max()
        assert rel_err(lnl, want["lnl"]) <= LNL_RTOL
        # the sum-table path: same value, finite derivatives
        eng.lk_edge_prepare(ra, rb)
        l0, d1, d2 = eng.lk_edge_eval([rt, 0.5 * rt])
        assert np.all(np.isfinite(l0)) and np.all(np.isfinite(d1)) and np.all(np.isfinite(d2))
        assert rel_err(l0[0], lnl) <= 1e-11
    finally:
        eng.set_option(eng.OPT_FUSED_TREE, 1)


def test_lk_jc69_sym_path(eng, oracle):
    m = mlmodel.create(("JC69",), 4, site_var=("gamma", 4, 1.0))
    _check_lk(eng, oracle, m, 14, 3000)


def test_lk_five_state_dynamic_kernel(eng, oracle):
    """gap as a 5th state (lib/mlModel.ml:108-113): S=5 goes through the run-time-S kernel."""
    m = mlmodel.create(("F81",), 5, pi=[0.2, 0.2, 0.25, 0.25, 0.1], site_var=("gamma", 4, 0.5))
    _check_lk(eng, oracle, m, 10, 2000)


def test_lk_aa_20_state(eng, oracle):
    """BASELINE config 4 shape (reduced): 20 states, K=4."""
    _check_lk(eng, oracle, aa_model(4), 16, 3000)


def test_lk_codon_61_state(eng, oracle):
    """BASELINE config 5 shape (reduced): 61 states, K=1."""
    _check_lk(eng, oracle, codon_model(), 12, 1500, mean_bl=0.05)


def test_lk_median_2_per_node_matches_score_tree(eng, oracle):
    """API-fidelity path: one phylo_lk_median_2 per node == the whole-tree entry point."""
    model = dna_gtr_g4()
    tr, ops, ra, rb, rt, n_nodes, tips = setup_lk(12, 2000, model, seed=9)
    eng.lk_set_model(model)
    eng.lk_set_tips(tips, capacity=n_nodes)
    whole = eng.lk_score_tree(ops, ra, rb, rt)
    eng.lk_set_tips(tips, capacity=n_nodes)
    for op in ops:
        eng.lk_median_2(int(op["parent"]), int(op["left"]), op["t_left"], int(op["right"]), op["t_right"])
    per_node = eng.lk_edge_lnl(ra, rb, [rt])[0]
    assert per_node == whole  # same kernels, same order: bit-identical


def test_lk_pulley_principle(eng):
    """Reversible model: lnL is the same at every root edge (independent property pin)."""
    model = dna_gtr_g4()
    tr = tree.random_tree(10, 21)
    tips = tree.evolve_tips(tr, model, 1500, 22)
    eng.lk_set_model(model)
    vals = []
    for e in tr.edges():
        ops, ra, rb, rt, n_nodes = tree.schedule(tr, root_edge=e)
        eng.lk_set_tips(tips, capacity=n_nodes)
        vals.append(eng.lk_score_tree(ops, ra, rb, rt))
    assert max(vals) - min(vals) <= 1e-9 * abs(vals[0])


def test_lk_edge_lnl_many_lengths(eng, oracle):
    model = dna_gtr_g4()
    tr, ops, ra, rb, rt, n_nodes, tips = setup_lk(10, 1200, model, seed=13)
    eng.lk_set_model(model)
    eng.lk_set_tips(tips, capacity=n_nodes)
    eng.lk_score_tree(ops, ra, rb, rt)
    ts = [1e-4, 0.01, 0.1, 0.5, 2.0]
    got = eng.lk_edge_lnl(ra, rb, ts)
    for t, g in zip(ts, got):
        w = oracle.lk_score_tree(model, tips, None, ops, n_nodes, ra, rb, t)["lnl"]
        assert rel_err(g, w) <= LNL_RTOL


def test_lk_block_partials_shard_invariance(eng):
    """Splitting the patterns in 1024-aligned shards and reducing the gathered level-1
    partials reproduces the single-engine lnL bit-for-bit (what N ranks do)."""
    model = dna_gtr_g4()
    tr, ops, ra, rb, rt, n_nodes, tips = setup_lk(8, 5000, model, seed=17)
    eng.lk_set_model(model)
    eng.lk_set_tips(tips, capacity=n_nodes)
    whole = eng.lk_score_tree(ops, ra, rb, rt)
    assert eng.reduce_partials(eng.lk_get_block_partials()) == whole
    parts = []
    for lo, hi in [(0, 2048), (2048, 3072), (3072, 5000)]:
        eng.lk_set_tips(np.ascontiguousarray(tips[:, lo:hi]), capacity=n_nodes)
        eng.lk_score_tree(ops, ra, rb, rt)
        parts.append(eng.lk_get_block_partials())
    assert eng.reduce_partials(np.concatenate(parts)) == whole


def test_lk_rejects_empty_mask(eng):
    model = dna_gtr_g4()
    tips = np.ones((4, 100), dtype=np.uint8)
    tips[2, 17] = 0
    eng.lk_set_model(model)
    with pytest.raises(engine.PhyloError) as ei:
        eng.lk_set_tips(tips)
    assert ei.value.code == -4


def test_lk_rejects_non_postorder(eng):
    model = dna_gtr_g4()
    tr, ops, ra, rb, rt, n_nodes, tips = setup_lk(8, 64, model, seed=2)
    eng.lk_set_model(model)
    eng.lk_set_tips(tips, capacity=n_nodes)
    with pytest.raises(engine.PhyloError):
        eng.lk_score_tree(ops[::-1].copy(), ra, rb, rt)


def test_lk_wide_masks_are_converted(eng, oracle):
    """uint32 tip masks for a 4-state model (bv32-style input) give the same lnL."""
    model = dna_gtr_g4()
    tr, ops, ra, rb, rt, n_nodes, tips = setup_lk(8, 700, model, seed=4)
    eng.lk_set_model(model)
    eng.lk_set_tips(tips, capacity=n_nodes)
    a = eng.lk_score_tree(ops, ra, rb, rt)
    eng.lk_set_tips(tips.astype(np.uint32) | 16, capacity=n_nodes)  # stray gap bit is ignored
    assert eng.lk_score_tree(ops, ra, rb, rt) == a


# ------------------------------------------------- branch-length loop (sum table) ----
def _edge_models():
    return [
        ("gtr_g4", dna_gtr_g4(), 12, 3000),
        ("gtr_g4_pinvar", dna_gtr_g4(pinvar=0.2), 10, 2500),
        ("jc69_sym", mlmodel.create(("JC69",), 4, site_var=("gamma", 4, 1.0)), 9, 2000),
        ("aa20", aa_model(), 8, 700),
        ("codon61", codon_model(), 6, 300),
    ]


@pytest.mark.parametrize("name,model,T,N", _edge_models(), ids=[m[0] for m in _edge_models()])
def test_lk_edge_sumtable_values_and_derivatives(eng, oracle, name, model, T, N):
    """phylo_lk_edge_prepare/_eval: lnL(t) from the sum table == the P-based root join
    (phylo_lk_edge_lnl, <= 1e-12) == the oracle (<= 1e-9); first and second derivatives match
    central differences of the P-based value."""
    tr, ops, ra, rb, rt, n_nodes, tips = setup_lk(T, N, model, seed=21, mean_bl=0.2)
    w = np.random.default_rng(3).integers(1, 5, N).astype(float)
    eng.lk_set_model(model)
    eng.lk_set_tips(tips, weights=w, capacity=n_nodes)
    eng.lk_score_tree(ops, ra, rb, rt)
    eng.lk_edge_prepare(ra, rb)
    ts = np.array([1e-6, 0.003, 0.05, 0.2, rt, 0.9, 3.0, 1e-12])
    lnl, d1, d2 = eng.lk_edge_eval(ts)
    direct = eng.lk_edge_lnl(ra, rb, ts)
    assert np.max(np.abs(lnl - direct) / np.abs(direct)) <= 1e-12
    for t, v in zip(ts[[2, 4]], lnl[[2, 4]]):
        assert rel_err(v, oracle.lk_score_tree(model, tips, w, ops, n_nodes, ra, rb, float(t))["lnl"]) <= LNL_RTOL
    # derivatives: 5-point central differences of the sum-table lnL itself (smooth in t)
    if model["Ui"] is not None:  # the symmetric path rounds t to float (lib/mlmodel.c:280): not differentiable
        for i in (2, 3, 5):
            t, h = ts[i], ts[i] * 1e-3
            f = eng.lk_edge_eval(np.array([t - 2 * h, t - h, t, t + h, t + 2 * h]))[0]
            fd1 = (f[0] - 8 * f[1] + 8 * f[3] - f[4]) / (12 * h)
            fd2 = (-f[0] + 16 * f[1] - 30 * f[2] + 16 * f[3] - f[4]) / (12 * h * h)
            assert abs(d1[i] - fd1) <= 1e-6 * max(1.0, abs(fd1)), (name, t, d1[i], fd1)
            assert abs(d2[i] - fd2) <= 1e-4 * max(1.0, abs(fd2)), (name, t, d2[i], fd2)


def test_lk_optimize_branch_matches_scalar_minimiser(eng, oracle):
    """Safeguarded Newton on the device sum table finds the same optimum as a bounded scalar
    minimiser on the P-based edge likelihood; also on a pendant edge (one side is a tip)."""
    from scipy.optimize import minimize_scalar

    model = dna_gtr_g4()
    tr, ops, ra, rb, rt, n_nodes, tips = setup_lk(14, 4000, model, seed=8, mean_bl=0.15)
    eng.lk_set_model(model)
    eng.lk_set_tips(tips, capacity=n_nodes)
    eng.lk_score_tree(ops, ra, rb, rt)
    for a, b in ((ra, rb), (int(ops[-1]["left"]), int(ops[-1]["right"]))):
        t_opt, l_opt, iters = eng.lk_optimize_branch(a, b, t0=0.5, t_min=1e-8, t_max=50.0, tol=1e-10)
        res = minimize_scalar(lambda t: -eng.lk_edge_lnl(a, b, [t])[0], bounds=(1e-8, 50.0), method="bounded",
                              options={"xatol": 1e-10})
        assert iters <= 30
        assert l_opt >= -res.fun - 1e-9 * abs(res.fun)
        assert abs(t_opt - res.x) <= 1e-5 * max(res.x, 1e-3), (t_opt, res.x)
        # first-order optimality from the engine's own derivative
        _, d1, d2 = eng.lk_edge_eval([t_opt])
        assert abs(d1[0]) <= 1e-5 * abs(d2[0]) * max(t_opt, 1e-3) + 1e-6


# ------------------------------------------------- general-TCM median (CostMatrix) ----
@pytest.mark.parametrize("S,metric", [(4, False), (4, True), (5, False), (6, True)])
def test_tcm_median_equals_costmatrix_restatement(eng, S, metric):
    """phylo_tcm_median_2 == the literal restatement of CostMatrix.find_median_general / _metric
    (lib/costMatrix.ml:68-124) for every character: median set and cost; weighted too."""
    from oracle.oracle import tcm_median_table

    rng = np.random.default_rng(S * 7 + metric)
    M = rng.integers(0, 9, size=(S, S)).astype(np.int32)
    M = (M + M.T) // 2
    np.fill_diagonal(M, 0)
    cost, med = tcm_median_table(M, metric)
    N = 5003
    chars = rng.integers(1, 1 << S, size=(2, N)).astype(np.uint8)
    w = rng.integers(0, 5, N).astype(float)
    for weights in (None, w):
        eng.fitch_set_tips(chars, S, weights=weights, capacity=4)
        eng.tcm_set_matrix(M, metric)
        got = eng.tcm_median_2(2, 0, 1)
        per_char = cost[chars[0], chars[1]]
        assert got == int((per_char * (1 if weights is None else w)).sum())
        assert np.array_equal(eng.fitch_get_states(2), med[chars[0], chars[1]].astype(np.uint8))
        assert eng.tcm_median_2(-1, 0, 1) == got  # cost only


def test_tcm_unit_costs_are_the_fitch_rule_on_a_tree(eng, oracle):
    """With the 0/1 matrix the TCM down-pass is the Fitch down-pass (test/costMatrixTest.ml:110-125):
    same sets at every node, same length."""
    ops, ra, rb, n_nodes, chars = _fitch_setup(20, 4097, 4, np.uint8, seed=12)
    eng.fitch_set_tips(chars, 4, capacity=n_nodes)
    want = oracle.fitch_score_tree(chars, None, ops, n_nodes, ra, rb, want_sets=True)
    for metric in (False, True):
        eng.tcm_set_matrix(1 - np.eye(4, dtype=np.int32), metric)
        assert eng.tcm_score_tree(ops, ra, rb) == want["length"]
        for op in ops:
            p = int(op["parent"])
            assert np.array_equal(eng.fitch_get_states(p), want["prelim"][p])


def test_tcm_tree_length_with_transversion_costs(eng):
    """Transition/transversion matrix (1/2) on a tree: node by node against the table oracle."""
    from oracle.oracle import tcm_median_table

    M = np.array([[0, 2, 1, 2], [2, 0, 2, 1], [1, 2, 0, 2], [2, 1, 2, 0]], dtype=np.int32)
    cost, med = tcm_median_table(M, False)
    ops, ra, rb, n_nodes, chars = _fitch_setup(14, 3000, 4, np.uint8, seed=21)
    eng.fitch_set_tips(chars, 4, capacity=n_nodes)
    eng.tcm_set_matrix(M)
    sets = {t: chars[t] for t in range(14)}
    total = 0
    for op in ops:
        p, l, r = int(op["parent"]), int(op["left"]), int(op["right"])
        total += int(cost[sets[l], sets[r]].sum())
        sets[p] = med[sets[l], sets[r]].astype(np.uint8)
    total += int(cost[sets[ra], sets[rb]].sum())
    assert eng.tcm_score_tree(ops, ra, rb) == total
    assert np.array_equal(eng.fitch_get_states(int(ops[-1]["parent"])), sets[int(ops[-1]["parent"])])


# ------------------------------------------------------ site-pattern compression ----
@pytest.mark.parametrize("dtype,T,N", [(np.uint8, 12, 50000), (np.uint8, 7, 1), (np.uint16, 5, 3001),
                                       (np.uint32, 33, 4097), (np.uint64, 3, 70000), (np.uint8, 256, 20000)])
def test_compress_patterns_equals_oracle(eng, dtype, T, N):
    """phylo_compress_patterns == the numpy oracle exactly: patterns in order of first occurrence,
    summed weights, site -> pattern map; element widths 1/2/4/8 bytes, record padding."""
    from oracle.oracle import compress_patterns

    rng = np.random.default_rng(N + T)
    base = rng.integers(1, 16, size=(T, max(1, N // 7))).astype(dtype)  # ~7 copies of every column
    masks = np.ascontiguousarray(base[:, rng.integers(0, base.shape[1], N)])
    w = rng.integers(1, 6, N).astype(float)
    for weights in (None, w):
        pats, wt, s2p = eng.compress_patterns(masks, weights)
        opats, owt, os2p = compress_patterns(masks, weights)
        assert pats.shape == opats.shape
        assert np.array_equal(pats, opats) and np.array_equal(wt, owt) and np.array_equal(s2p, os2p)


def test_compress_patterns_extremes(eng):
    from oracle.oracle import compress_patterns

    same = np.full((9, 5000), 4, dtype=np.uint8)
    pats, wt, s2p = eng.compress_patterns(same)
    assert pats.shape == (9, 1) and wt[0] == 5000 and not s2p.any()
    distinct = np.arange(1, 40001, dtype=np.uint64).reshape(1, -1).repeat(3, axis=0)
    pats, wt, s2p = eng.compress_patterns(distinct)
    assert pats.shape == (3, 40000) and np.array_equal(pats, distinct) and np.all(wt == 1)
    assert np.array_equal(s2p, np.arange(40000))


def test_compressed_alignment_scores_like_the_raw_one(eng, oracle):
    """lnL and Fitch length of (patterns, weights) == those of the raw alignment."""
    model = dna_gtr_g4()
    tr, ops, ra, rb, rt, n_nodes, tips = setup_lk(10, 30000, model, seed=9, mean_bl=0.05)
    pats, wt, _ = eng.compress_patterns(tips)
    assert pats.shape[1] < tips.shape[1] // 2
    eng.lk_set_model(model)
    eng.lk_set_tips(tips, capacity=n_nodes)
    raw = eng.lk_score_tree(ops, ra, rb, rt)
    eng.lk_set_tips(pats, weights=wt, capacity=n_nodes)
    assert rel_err(eng.lk_score_tree(ops, ra, rb, rt), raw) <= LNL_RTOL
    chars = np.where(tips == 15, 15, tips).astype(np.uint8)
    eng.fitch_set_tips(chars, 4, capacity=n_nodes)
    raw_len = eng.fitch_score_tree(ops, ra, rb)
    eng.fitch_set_tips(pats, 4, weights=wt, capacity=n_nodes)
    assert eng.fitch_score_tree(ops, ra, rb) == raw_len


# ------------------------------------------------------------------------ Fitch ----
def _fitch_setup(T, N, n_states, dtype, seed=1):
    tr = tree.random_tree(T, seed)
    ops, ra, rb, rt, n_nodes = tree.schedule(tr)
    chars = tree.random_fitch_chars(T, N, n_states, seed + 5, dtype=dtype)
    return ops, ra, rb, n_nodes, chars


@pytest.fixture(params=[1, 3, 2, 0], ids=["auto", "tile", "regwalk", "l2walk"])
def fitch_walk(eng, request):
    """All whole-tree Fitch kernels: automatic choice, on-chip tiles, register walk, L2 walk."""
    eng.set_option(eng.OPT_FITCH_WALK, request.param)
    yield request.param
    eng.set_option(eng.OPT_FITCH_WALK, 1)


@pytest.mark.parametrize("T,N", [(16, 10000), (64, 100003), (5, 31), (3, 1), (64, 32), (2, 77)])
def test_fitch_tree_length_and_sets_bit_exact(eng, oracle, fitch_walk, T, N):
    ops, ra, rb, n_nodes, chars = _fitch_setup(T, N, 4, np.uint8)
    eng.fitch_set_tips(chars, 4, capacity=n_nodes)
    length = eng.fitch_score_tree(ops, ra, rb)
    want = oracle.fitch_score_tree(chars, None, ops, n_nodes, ra, rb, want_sets=True)
    assert length == want["length"]
    costs = eng.fitch_get_node_costs()
    for op in ops:
        p = int(op["parent"])
        assert costs[p] == want["node_cost"][p]
        assert np.array_equal(eng.fitch_get_states(p), want["prelim"][p])
    for t in range(T):
        assert np.array_equal(eng.fitch_get_states(t), chars[t])


@pytest.mark.parametrize("dtype,n_states", [(np.uint8, 5), (np.uint8, 6), (np.uint8, 8), (np.uint16, 11),
                                            (np.uint32, 22), (np.uint32, 32), (np.uint64, 40),
                                            (np.uint64, 64)])
def test_fitch_all_widths(eng, oracle, fitch_walk, dtype, n_states):
    """W in {8,16,32,64} (lib/bitvector/bv.h:29-55) and plane counts up to 64."""
    ops, ra, rb, n_nodes, chars = _fitch_setup(12, 3001, n_states, dtype, seed=3)
    eng.fitch_set_tips(chars, n_states, capacity=n_nodes)
    length = eng.fitch_score_tree(ops, ra, rb)
    want = oracle.fitch_score_tree(chars, None, ops, n_nodes, ra, rb, want_sets=True)
    assert length == want["length"]
    p = int(ops[-1]["parent"])
    assert np.array_equal(eng.fitch_get_states(p), want["prelim"][p])


def test_fitch_median_2_and_distance_per_node(eng, oracle):
    """NonAdditive.median_2 / bv_fitch per node, then bv_distance at the root edge."""
    ops, ra, rb, n_nodes, chars = _fitch_setup(20, 5000, 4, np.uint8, seed=8)
    eng.fitch_set_tips(chars, 4, capacity=n_nodes)
    sets = {t: chars[t] for t in range(20)}
    total = 0
    for op in ops:
        p, l, r = int(op["parent"]), int(op["left"]), int(op["right"])
        cost = eng.fitch_median_2(p, l, r)
        c, k = oracle.fitch_median2(sets[l], sets[r])
        sets[p] = c
        assert cost == k
        assert np.array_equal(eng.fitch_get_states(p), c)
        total += cost
    d = eng.fitch_distance(ra, rb)
    assert d == oracle.fitch_distance(sets[ra], sets[rb])
    assert total + d == eng.fitch_score_tree(ops, ra, rb)


def test_fitch_weighted(eng, oracle, fitch_walk):
    ops, ra, rb, n_nodes, chars = _fitch_setup(16, 7000, 4, np.uint8, seed=4)
    w = np.random.default_rng(2).integers(0, 9, 7000).astype(float)
    eng.fitch_set_tips(chars, 4, weights=w, capacity=n_nodes)
    assert eng.fitch_score_tree(ops, ra, rb) == oracle.fitch_score_tree(chars, w, ops, n_nodes, ra, rb)["length"]
    with pytest.raises(engine.PhyloError):
        eng.fitch_set_tips(chars, 4, weights=w + 0.5, capacity=n_nodes)


def test_fitch_caterpillar_and_partial_schedule(eng, oracle, fitch_walk):
    """A 300-taxon caterpillar (stack-free plan, long dependency chain) and a partial
    re-evaluation whose other operands are sets already resident from the first call."""
    T, N = 300, 2500
    tr = tree.caterpillar_tree(T)
    ops, ra, rb, rt, n_nodes = tree.schedule(tr)
    chars = tree.random_fitch_chars(T, N, 4, 11, dtype=np.uint8)
    eng.fitch_set_tips(chars, 4, capacity=n_nodes)
    want = oracle.fitch_score_tree(chars, None, ops, n_nodes, ra, rb, want_sets=True)
    assert eng.fitch_score_tree(ops, ra, rb) == want["length"]
    # re-run only the last third of the schedule: earlier parents are read back from HBM
    cut = 2 * len(ops) // 3
    tail_cost = sum(int(want["node_cost"][int(op["parent"])]) for op in ops[cut:])
    d = oracle.fitch_distance(want["prelim"][ra], want["prelim"][rb])
    assert eng.fitch_score_tree(ops[cut:].copy(), ra, rb) == tail_cost + d
    for op in ops[-4:]:
        p = int(op["parent"])
        assert np.array_equal(eng.fitch_get_states(p), want["prelim"][p])


def test_fitch_tile_program_cache(eng, oracle):
    """The tile kernel keeps the compiled program of the last schedule. Same schedule again (hit), another
    topology (miss), the first one again, new characters of another length (buffers move), weights switched
    on and off (another kernel variant and publication protocol): always the oracle's numbers."""
    eng.set_option(eng.OPT_FITCH_WALK, 3)
    try:
        T = 40
        scheds = []
        for seed in (1, 2):
            tr = tree.random_tree(T, seed)
            scheds.append(tree.schedule(tr))
        for N, cseed in ((5000, 3), (70001, 4)):
            chars = tree.random_fitch_chars(T, N, 4, cseed, dtype=np.uint8)
            n_nodes = scheds[0][4]
            eng.fitch_set_tips(chars, 4, capacity=n_nodes)
            wants = [oracle.fitch_score_tree(chars, None, s[0], s[4], s[1], s[2], want_sets=True) for s in scheds]
            for which in (0, 0, 1, 0, 1, 1):
                ops, ra, rb, rt, nn = scheds[which]
                assert eng.fitch_score_tree(ops, ra, rb) == wants[which]["length"], (N, which)
                costs = eng.fitch_get_node_costs()
                for op in ops:
                    assert costs[int(op["parent"])] == wants[which]["node_cost"][int(op["parent"])]
                p = int(ops[-1]["parent"])
                assert np.array_equal(eng.fitch_get_states(p), wants[which]["prelim"][p])
            w = np.random.default_rng(N).integers(0, 5, N).astype(float)
            ops, ra, rb, rt, nn = scheds[0]
            eng.fitch_set_tips(chars, 4, weights=w, capacity=n_nodes)
            for _ in range(2):
                assert eng.fitch_score_tree(ops, ra, rb) == oracle.fitch_score_tree(chars, w, ops, nn, ra, rb)["length"]
            eng.fitch_set_tips(chars, 4, capacity=n_nodes)
            for _ in range(2):
                assert eng.fitch_score_tree(ops, ra, rb) == wants[0]["length"]
    finally:
        eng.set_option(eng.OPT_FITCH_WALK, 1)


def test_fitch_length_only_mode(eng, oracle):
    """PHYLO_OPT_RETAIN_CLV = 0 for Fitch: the length alone (tree evaluated from its centre edge, no interior
    set written); the parents' slots are invalid afterwards and a retaining call brings them back."""
    for T, N, seed in ((64, 100003, 1), (16, 4099, 2), (3, 77, 3), (40, 1, 4)):
        ops, ra, rb, n_nodes, chars = _fitch_setup(T, N, 4, np.uint8, seed=seed)
        want = oracle.fitch_score_tree(chars, None, ops, n_nodes, ra, rb, want_sets=True)
        eng.fitch_set_tips(chars, 4, capacity=n_nodes)
        eng.set_option(eng.OPT_RETAIN_CLV, 0)
        try:
            for _ in range(2):
                assert eng.fitch_score_tree(ops, ra, rb) == want["length"]
            if len(ops):
                with pytest.raises(engine.PhyloError):
                    eng.fitch_get_states(int(ops[-1]["parent"]))
        finally:
            eng.set_option(eng.OPT_RETAIN_CLV, 1)
        assert eng.fitch_score_tree(ops, ra, rb) == want["length"]
        for op in ops[-3:]:
            p = int(op["parent"])
            assert np.array_equal(eng.fitch_get_states(p), want["prelim"][p])
    # weighted
    ops, ra, rb, n_nodes, chars = _fitch_setup(20, 9000, 4, np.uint8, seed=9)
    w = np.random.default_rng(5).integers(0, 7, 9000).astype(float)
    eng.fitch_set_tips(chars, 4, weights=w, capacity=n_nodes)
    eng.set_option(eng.OPT_RETAIN_CLV, 0)
    try:
        assert eng.fitch_score_tree(ops, ra, rb) == oracle.fitch_score_tree(chars, w, ops, n_nodes, ra, rb)["length"]
    finally:
        eng.set_option(eng.OPT_RETAIN_CLV, 1)


def test_fitch_uppass_final_sets(eng, oracle, fitch_walk):
    ops, ra, rb, n_nodes, chars = _fitch_setup(24, 4099, 4, np.uint8, seed=6)
    eng.fitch_set_tips(chars, 4, capacity=n_nodes)
    eng.fitch_score_tree(ops, ra, rb)
    eng.fitch_uppass(ops, ra, rb)
    want = oracle.fitch_score_tree(chars, None, ops, n_nodes, ra, rb, want_sets=True)
    fin = oracle.fitch_uppass(24, want["prelim"], ops, ra, rb)
    for v in range(n_nodes):
        assert np.array_equal(eng.fitch_get_states(v, final=True), fin[v]), v


def test_fitch_rejects_empty_character(eng):
    chars = np.ones((4, 50), dtype=np.uint8)
    chars[1, 3] = 0
    with pytest.raises(engine.PhyloError) as ei:
        eng.fitch_set_tips(chars, 4)
    assert ei.value.code == -4


def test_bv_set_algebra(eng, oracle):
    """bv_union/inter/popcount/saturation/poly_saturation/compare (lib/bitvector/bv.c:59-144)."""
    rng = np.random.default_rng(5)
    N = 3333
    chars = rng.integers(1, 64, size=(3, N)).astype(np.uint8)
    chars[2] = chars[0]
    eng.fitch_set_tips(chars, 6, capacity=8)
    a, b = chars[0], chars[1]
    eng.bv_union(4, 0, 1)
    eng.bv_inter(5, 0, 1)
    assert np.array_equal(eng.fitch_get_states(4), a | b)
    assert np.array_equal(eng.fitch_get_states(5), a & b)
    pop = np.array([bin(int(x)).count("1") for x in a])
    assert eng.bv_popcount(0) == int(pop.sum())
    assert eng.bv_saturation(0, 0b101) == int(((a & 0b101) != 0).sum())
    for n in range(0, 8):
        assert eng.bv_poly_saturation(0, n) == int((pop == n).sum())
    assert eng.bv_compare(0, 2) == 0
    first = int(np.nonzero(a != b)[0][0])
    assert eng.bv_compare(0, 1) == (1 if a[first] > b[first] else -1)
    assert eng.bv_compare(1, 0) == -eng.bv_compare(0, 1)
    # Bitvector.of_array on an interior slot, then the union laws of test/bitvectorTest.ml:44-57
    eng.fitch_set_states(6, b)
    eng.bv_union(7, 6, 6)
    assert np.array_equal(eng.fitch_get_states(7), b)


# ------------------------------------------------------------ tree-fused kernel ----
def _score(eng, model, tips, ops, ra, rb, rt, n_nodes, fused, retain, weights=None):
    eng.set_option(eng.OPT_FUSED_TREE, fused)
    eng.set_option(eng.OPT_RETAIN_CLV, retain)
    try:
        eng.lk_set_model(model)
        eng.lk_set_tips(tips, weights=weights, capacity=n_nodes)
        before = eng.profile_get().get("tree_fused", (0, 0))[1]
        lnl = eng.lk_score_tree(ops, ra, rb, rt)
        used_fused = eng.profile_get().get("tree_fused", (0, 0))[1] > before
        return lnl, used_fused
    finally:
        eng.set_option(eng.OPT_FUSED_TREE, 1)
        eng.set_option(eng.OPT_RETAIN_CLV, 1)


@pytest.mark.parametrize("T,N,kind", [(16, 10000, "random"), (64, 3000, "random"), (200, 700, "caterpillar"),
                                      (3, 50, "random"), (2, 33, "random"), (40, 1, "random"),
                                      (600, 257, "random"), (24, 90000, "random")])
def test_fused_tree_equals_per_node_path_bitwise(eng, oracle, T, N, kind):
    """The single-launch tree-fused kernel and the one-kernel-per-node path run the same
    arithmetic in the same order: lnL, site lnL, CLVs and scale counters are bit-identical;
    both are within 1e-9 of the oracle."""
    model = dna_gtr_g4()
    tr, ops, ra, rb, rt, n_nodes, tips = setup_lk(T, N, model, seed=31, tree_kind=kind, mean_bl=0.3)
    eng.profile(True)
    try:
        a, fa = _score(eng, model, tips, ops, ra, rb, rt, n_nodes, fused=1, retain=1)
        site_a = eng.lk_get_site_lnl()
        clv_a = [eng.lk_get_clv(int(op["parent"])) for op in ops[-3:]]
        b, fb = _score(eng, model, tips, ops, ra, rb, rt, n_nodes, fused=0, retain=1)
        site_b = eng.lk_get_site_lnl()
        clv_b = [eng.lk_get_clv(int(op["parent"])) for op in ops[-3:]]
        c, fc = _score(eng, model, tips, ops, ra, rb, rt, n_nodes, fused=1, retain=0)
        site_c = eng.lk_get_site_lnl()
        # fused=2: the tile kernel (thread = pattern x rate class) instead of the warp-autonomous one
        d, fd = _score(eng, model, tips, ops, ra, rb, rt, n_nodes, fused=2, retain=1)
        site_d = eng.lk_get_site_lnl()
        clv_d = [eng.lk_get_clv(int(op["parent"])) for op in ops[-3:]]
        f, ff = _score(eng, model, tips, ops, ra, rb, rt, n_nodes, fused=2, retain=0)
    finally:
        eng.profile(False)
    assert fa and fc and fd and ff and not fb
    assert a == b == c == d == f
    assert np.array_equal(site_a, site_b) and np.array_equal(site_a, site_c) and np.array_equal(site_a, site_d)
    for (ca, sa), (cb, sb), (cd, sd) in zip(clv_a, clv_b, clv_d):
        assert np.array_equal(ca, cb) and np.array_equal(sa, sb)
        assert np.array_equal(ca, cd) and np.array_equal(sa, sd)
    want = oracle.lk_score_tree(model, tips, None, ops, n_nodes, ra, rb, rt)["lnl"]
    assert rel_err(a, want) <= LNL_RTOL


@pytest.mark.parametrize("tune", ["2,2,1", "1,2,0", "1,1,1", "3,1,0"])
@pytest.mark.parametrize("K", [4, 2, 1])
def test_fused_tree_kernel_variants_bitwise(eng, tune, K, monkeypatch):
    """The warp-autonomous kernel in its geometry variants (shared-memory stack levels, R = 1 or
    2 patterns per thread, group order; PHYLO_TREEW_TUNE) == the per-node path, bit for bit:
    lnL, site lnL, every CLV and scale counter. N is not a multiple of 64 (odd group count)."""
    sv = ("gamma", K, 0.6) if K > 1 else None
    m = mlmodel.create(("GTR", [1.0, 2.5, 0.8, 1.2, 3.0]), 4, pi=[0.3, 0.2, 0.25, 0.25], site_var=sv)
    tr, ops, ra, rb, rt, n_nodes, tips = setup_lk(40, 7000 + 33, m, seed=44, mean_bl=0.4)
    monkeypatch.setenv("PHYLO_TREEW_TUNE", tune)
    eng.profile(True)
    try:
        a, fa = _score(eng, m, tips, ops, ra, rb, rt, n_nodes, fused=1, retain=1)
        site_a = eng.lk_get_site_lnl()
        clv_a = [eng.lk_get_clv(int(op["parent"])) for op in ops]
        c, fc = _score(eng, m, tips, ops, ra, rb, rt, n_nodes, fused=1, retain=0)
        monkeypatch.delenv("PHYLO_TREEW_TUNE")
        b, fb = _score(eng, m, tips, ops, ra, rb, rt, n_nodes, fused=0, retain=1)
        site_b = eng.lk_get_site_lnl()
        clv_b = [eng.lk_get_clv(int(op["parent"])) for op in ops]
    finally:
        eng.profile(False)
    assert fa and fc and not fb and a == b == c
    assert np.array_equal(site_a, site_b)
    for (ca, sa), (cb, sb) in zip(clv_a, clv_b):
        assert np.array_equal(ca, cb) and np.array_equal(sa, sb)


@pytest.mark.parametrize("tune", [None, "1,2,0", "1,1,1"])
@pytest.mark.parametrize("K,codes,missing", [(4, 0.3, 0.05), (4, 0.0, 0.0), (4, 0.0, 0.6), (4, 0.002, 0.01),
                                             (2, 0.2, 0.1), (1, 0.2, 0.1)])
def test_fused_tree_tip_tables_any_state_set_bitwise(eng, oracle, K, codes, missing, tune, monkeypatch):
    """A TIP operand of the warp-autonomous kernel is a lookup in the branch's tip table (single states and the
    missing cell) or, for other IUPAC-style state sets, the general expression read from the same table: every mix
    of the three -- none missing, mostly missing, a few partial sets among single states (one such lane sends its
    warp down the arithmetic path), many partial sets -- gives the per-node path's lnL, site lnL, CLVs and scale
    counters bit for bit, and the oracle's lnL."""
    sv = ("gamma", K, 0.6) if K > 1 else None
    m = mlmodel.create(("GTR", [1.0, 2.5, 0.8, 1.2, 3.0]), 4, pi=[0.3, 0.2, 0.25, 0.25], site_var=sv)
    T, N = 48, 9000 + 37
    tr, ops, ra, rb, rt, n_nodes, tips = setup_lk(T, N, m, seed=52, mean_bl=0.25, missing=missing)
    rng = np.random.default_rng(9)
    amb = rng.random((T, N)) < codes
    tips = np.where(amb, tips | rng.integers(1, 16, size=(T, N), dtype=np.uint8), tips).astype(np.uint8)
    if tune:
        monkeypatch.setenv("PHYLO_TREEW_TUNE", tune)
    eng.profile(True)
    try:
        a, fa = _score(eng, m, tips, ops, ra, rb, rt, n_nodes, fused=1, retain=1)
        site_a = eng.lk_get_site_lnl()
        clv_a = [eng.lk_get_clv(int(op["parent"])) for op in ops]
        c, fc = _score(eng, m, tips, ops, ra, rb, rt, n_nodes, fused=1, retain=0)
        if tune:
            monkeypatch.delenv("PHYLO_TREEW_TUNE")
        b, fb = _score(eng, m, tips, ops, ra, rb, rt, n_nodes, fused=0, retain=1)
        site_b = eng.lk_get_site_lnl()
        clv_b = [eng.lk_get_clv(int(op["parent"])) for op in ops]
    finally:
        eng.profile(False)
    assert fa and fc and not fb and a == b == c
    assert np.array_equal(site_a, site_b)
    for (ca, sa), (cb, sb) in zip(clv_a, clv_b):
        assert np.array_equal(ca, cb) and np.array_equal(sa, sb)
    assert rel_err(a, oracle.lk_score_tree(m, tips, None, ops, n_nodes, ra, rb, rt)["lnl"]) <= LNL_RTOL


@pytest.mark.parametrize("K", [1, 2, 8])
def test_fused_tree_other_rate_counts(eng, oracle, K):
    sv = ("gamma", K, 0.8) if K > 1 else None
    m = mlmodel.create(("HKY85", 2.0), 4, pi=[0.1, 0.2, 0.3, 0.4], site_var=sv)
    tr, ops, ra, rb, rt, n_nodes, tips = setup_lk(20, 2100, m, seed=5)
    eng.profile(True)
    try:
        a, fa = _score(eng, m, tips, ops, ra, rb, rt, n_nodes, fused=1, retain=1)
        clv_a = [eng.lk_get_clv(int(op["parent"])) for op in ops]
        b, fb = _score(eng, m, tips, ops, ra, rb, rt, n_nodes, fused=0, retain=1)
        clv_b = [eng.lk_get_clv(int(op["parent"])) for op in ops]
        c, fc = _score(eng, m, tips, ops, ra, rb, rt, n_nodes, fused=2, retain=1)
    finally:
        eng.profile(False)
    assert fa and fc and not fb and a == b == c
    for (ca, sa), (cb, sb) in zip(clv_a, clv_b):
        assert np.array_equal(ca, cb) and np.array_equal(sa, sb)
    assert rel_err(a, oracle.lk_score_tree(m, tips, None, ops, n_nodes, ra, rb, rt)["lnl"]) <= LNL_RTOL


def test_fused_tree_no_retain_leaves_no_clvs(eng):
    model = dna_gtr_g4()
    tr, ops, ra, rb, rt, n_nodes, tips = setup_lk(10, 500, model, seed=2)
    _score(eng, model, tips, ops, ra, rb, rt, n_nodes, fused=1, retain=0)
    with pytest.raises(engine.PhyloError):
        eng.lk_get_clv(int(ops[0]["parent"]))


def test_fused_tree_incremental_rescoring_from_stored_clvs(eng, oracle):
    """After a full evaluation, change one branch and re-evaluate only the path from that
    branch to the root edge: the other operands are CLVs already resident in HBM (OPK_STORED).
    Result == full re-evaluation of the modified tree."""
    model = dna_gtr_g4()
    tr, ops, ra, rb, rt, n_nodes, tips = setup_lk(24, 3000, model, seed=12)
    eng.lk_set_model(model)
    eng.lk_set_tips(tips, capacity=n_nodes)
    eng.lk_score_tree(ops, ra, rb, rt)
    ops2 = ops.copy()
    i0 = 4
    ops2[i0]["t_left"] *= 3.0
    # path from op i0's parent up to its root end
    parent_of = {int(o["left"]): j for j, o in enumerate(ops2)}
    parent_of.update({int(o["right"]): j for j, o in enumerate(ops2)})
    path, j = [i0], i0
    while int(ops2[j]["parent"]) in parent_of:
        j = parent_of[int(ops2[j]["parent"])]
        path.append(j)
    sub = ops2[sorted(path)]
    inc = eng.lk_score_tree(sub, ra, rb, rt)
    full = oracle.lk_score_tree(model, tips, None, ops2, n_nodes, ra, rb, rt)["lnl"]
    assert rel_err(inc, full) <= LNL_RTOL
    eng.lk_set_tips(tips, capacity=n_nodes)
    assert eng.lk_score_tree(ops2, ra, rb, rt) == inc


@pytest.mark.parametrize("N", [700, 300000, 600001])
def test_lk_score_alignment_pipelined_equals_two_calls(eng, N):
    """phylo_lk_score_alignment (slab-pipelined upload + fused scoring) == set_tips + score_tree,
    bit for bit, for sizes below and above the slab threshold, with weights."""
    model = dna_gtr_g4()
    tr = tree.random_tree(12, 3)
    ops, ra, rb, rt, n_nodes = tree.schedule(tr)
    base = tree.evolve_tips(tr, model, 5000, seed=4)
    tips = np.ascontiguousarray(np.tile(base, (1, N // 5000 + 1))[:, :N])
    w = np.random.default_rng(1).integers(1, 4, N).astype(float)
    eng.lk_set_model(model)
    eng.lk_set_tips(tips, weights=w, capacity=n_nodes)
    two = eng.lk_score_tree(ops, ra, rb, rt)
    clv_two = eng.lk_get_clv(int(ops[-1]["parent"]))
    eng.lk_set_tips(np.ascontiguousarray(tips[:, ::-1]), weights=w, capacity=n_nodes)  # clobber device state
    one = eng.lk_score_alignment(tips, ops, ra, rb, rt, weights=w, capacity=n_nodes)
    clv_one = eng.lk_get_clv(int(ops[-1]["parent"]))
    assert one == two
    assert np.array_equal(clv_one[0], clv_two[0]) and np.array_equal(clv_one[1], clv_two[1])
    bad = tips.copy()
    bad[3, N // 2] = 0
    with pytest.raises(engine.PhyloError) as ei:
        eng.lk_score_alignment(bad, ops, ra, rb, rt, capacity=n_nodes)
    assert ei.value.code == -4


@pytest.mark.parametrize("S,N", [(20, 2100), (20, 160000 + 13), (61, 152000 + 5)])
def test_lk_score_alignment_large_alphabets_slabs_equal_two_calls(eng, S, N):
    """20 / 61 states through phylo_lk_score_alignment: the tree-fused DMMA kernel launched slab by slab (whole rounds
    of its CTA chunks: 37 blocks, 74, the rest) while the later slabs are still being uploaded == set_tips + score_tree,
    bit for bit (lnL, site lnL, the last CLV and its scale counters); a bad cell in the last slab is still reported."""
    from helpers import aa_model, codon_model
    model = aa_model(4) if S == 20 else codon_model()
    T = 6
    tr = tree.random_tree(T, 3)
    ops, ra, rb, rt, n_nodes = tree.schedule(tr)
    base = tree.evolve_tips(tr, model, 3000, seed=4, dtype=mask_dtype(S))
    tips = np.ascontiguousarray(np.tile(base, (1, N // 3000 + 1))[:, :N])
    w = np.random.default_rng(1).integers(1, 4, N).astype(float)
    eng.lk_set_model(model)
    eng.lk_set_tips(tips, weights=w, capacity=n_nodes)
    two = eng.lk_score_tree(ops, ra, rb, rt)
    site_two = eng.lk_get_site_lnl()
    clv_two = eng.lk_get_clv(int(ops[-1]["parent"]))
    eng.lk_set_tips(np.ascontiguousarray(tips[:, ::-1]), weights=w, capacity=n_nodes)  # clobber device state
    one = eng.lk_score_alignment(tips, ops, ra, rb, rt, weights=w, capacity=n_nodes)
    assert one == two
    assert np.array_equal(eng.lk_get_site_lnl(), site_two)
    clv_one = eng.lk_get_clv(int(ops[-1]["parent"]))
    assert np.array_equal(clv_one[0], clv_two[0]) and np.array_equal(clv_one[1], clv_two[1])
    bad = tips.copy()
    bad[3, N - 5] = 0
    with pytest.raises(engine.PhyloError) as ei:
        eng.lk_score_alignment(bad, ops, ra, rb, rt, capacity=n_nodes)
    assert ei.value.code == -4


# ------------------------------------------------------------ compact upload formats ----
@pytest.mark.parametrize("N", [1, 2, 33, 1024, 1025, 70001])
@pytest.mark.parametrize("pinvar", [None, 0.15])
def test_lk_packed_nibble_input_equals_byte_input(eng, oracle, N, pinvar):
    """mask_bytes = 0: two 4-bit masks per byte. Same lnL / site lnL / CLVs bit for bit as the one-byte
    form under the tree-fused and the per-node paths (the latter rebuilds its byte rows lazily), through
    set_tips and through the pipelined score_alignment."""
    model = dna_gtr_g4(pinvar=pinvar)
    tr, ops, ra, rb, rt, n_nodes, tips = setup_lk(9, N, model, seed=11)
    packed = engine.pack_nibbles(tips)
    eng.lk_set_model(model)
    try:
        for fused in (1, 0):
            eng.set_option(eng.OPT_FUSED_TREE, fused)
            eng.lk_set_tips(tips, capacity=n_nodes)
            want = eng.lk_score_tree(ops, ra, rb, rt)
            site = eng.lk_get_site_lnl()
            clv, sc = eng.lk_get_clv(int(ops[-1]["parent"]))
            eng.lk_set_tips(packed, capacity=n_nodes, packed_n=N)
            assert eng.lk_score_tree(ops, ra, rb, rt) == want
            assert np.array_equal(eng.lk_get_site_lnl(), site)
            clv2, sc2 = eng.lk_get_clv(int(ops[-1]["parent"]))
            assert np.array_equal(clv, clv2) and np.array_equal(sc, sc2)
            assert eng.lk_edge_lnl(0, int(ops[-1]["parent"]), [0.1])[0] == eng.lk_edge_lnl(0, int(ops[-1]["parent"]), [0.1])[0]
            assert eng.lk_score_alignment(packed, ops, ra, rb, rt, capacity=n_nodes, packed_n=N) == want
        ref = oracle.lk_score_tree(model, tips, None, ops, n_nodes, ra, rb, rt)["lnl"]
        assert rel_err(want, ref) <= LNL_RTOL
    finally:
        eng.set_option(eng.OPT_FUSED_TREE, 1)


def test_lk_packed_input_rejects_empty_nibbles_and_other_alphabets(eng):
    model = dna_gtr_g4()
    tr, ops, ra, rb, rt, n_nodes, tips = setup_lk(6, 301, model, seed=2)
    eng.lk_set_model(model)
    bad = tips.copy()
    bad[3, 300] = 0
    with pytest.raises(engine.PhyloError):
        eng.lk_set_tips(engine.pack_nibbles(bad), capacity=n_nodes, packed_n=301)
    eng.lk_set_model(aa_model(1))
    with pytest.raises(engine.PhyloError):
        eng.lk_set_tips(engine.pack_nibbles(tips), capacity=n_nodes, packed_n=301)
    eng.lk_set_model(model)


@pytest.mark.parametrize("N,ns,dt", [(1, 4, np.uint8), (33, 4, np.uint8), (4096 + 17, 4, np.uint8), (2000, 6, np.uint8),
                                     (777, 20, np.uint32), (100, 61, np.uint64)])
def test_fitch_plane_input_equals_element_input(eng, oracle, N, ns, dt):
    """elt_bytes = 0: the characters arrive bit-sliced (phylo_fitch_pack_planes); lengths, per-node costs and
    sets equal the element-wise upload's and the oracle's."""
    T = 10
    tr = tree.random_tree(T, seed=6)
    ops, ra, rb, rt, n_nodes = tree.schedule(tr)
    chars = tree.random_fitch_chars(T, N, ns, seed=8, ambiguity=0.1, dtype=dt)
    want = oracle.fitch_score_tree(chars, None, ops, n_nodes, ra, rb, want_sets=True)
    eng.fitch_set_tips_planes(engine.fitch_pack_planes(chars, ns), N, ns, capacity=n_nodes)
    assert eng.fitch_score_tree(ops, ra, rb) == want["length"]
    for node in (0, T - 1, int(ops[0]["parent"]), int(ops[-1]["parent"])):
        assert np.array_equal(eng.fitch_get_states(node).astype(np.uint64), want["prelim"][node].astype(np.uint64))
    bad = chars.copy()
    bad[2, N - 1] = 0
    with pytest.raises(engine.PhyloError):
        eng.fitch_set_tips_planes(engine.fitch_pack_planes(bad, ns), N, ns, capacity=n_nodes)
