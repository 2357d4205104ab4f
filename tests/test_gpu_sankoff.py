"""Cost-vector Sankoff parsimony (sankoff_kernels.cuh) against the numpy oracle: lengths and every node's
cost vectors bit-exact; tree-fused == node-by-node; the 0/1 matrix reproduces the Fitch length (the
reference's test property for its set medians, test/costMatrixTest.ml:110-125)."""
import numpy as np
import pytest

from oracle import oracle as O
from phylocaml_b200 import tree

pytestmark = pytest.mark.gpu


def _chars(T, N, S, seed, ambiguous=0.1):
    rng = np.random.default_rng(seed)
    states = rng.integers(0, S, (T, N))
    codes = (np.uint64(1) << states.astype(np.uint64)).astype(np.uint32)
    amb = rng.random((T, N)) < ambiguous
    extra = (np.uint64(1) << rng.integers(0, S, (T, N)).astype(np.uint64)).astype(np.uint32)
    codes[amb] |= extra[amb]
    miss = rng.random((T, N)) < 0.02
    codes[miss] = np.uint32((1 << S) - 1) if S < 32 else np.uint32(0xFFFFFFFF)
    return codes


@pytest.mark.parametrize("S,T,N,kind", [(4, 12, 3000, "random"), (5, 9, 777, "random"), (20, 10, 600, "random"),
                                        (21, 7, 300, "caterpillar"), (32, 6, 257, "random"), (3, 40, 500, "caterpillar")])
def test_sankoff_equals_oracle(eng, S, T, N, kind):
    rng = np.random.default_rng(S * 100 + T)
    M = rng.integers(1, 60, (S, S)).astype(np.int32)  # asymmetric, non-metric
    np.fill_diagonal(M, 0)
    codes = _chars(T, N, S, seed=S + N)
    w = rng.integers(0, 5, N).astype(float)
    tr = tree.random_tree(T, 3) if kind == "random" else tree.caterpillar_tree(T, 0.1)
    ops, ra, rb, rt, n_nodes = tree.schedule(tr)
    want = O.sankoff_score_tree(codes, M, w, ops, n_nodes, ra, rb)
    eng.sankoff_set_matrix(M)
    eng.sankoff_set_tips(codes, S, weights=w, capacity=n_nodes)
    got = {}
    for fused in (1, 0):
        eng.set_option(eng.OPT_FUSED_TREE, fused)
        got[fused] = eng.sankoff_score_tree(ops, ra, rb)
        for op in ops[[0, len(ops) // 2, len(ops) - 1]]:
            p = int(op["parent"])
            assert np.array_equal(eng.sankoff_get_costs(p), want["vec"][p]), (fused, p)
    eng.set_option(eng.OPT_FUSED_TREE, 1)
    assert got[1] == got[0] == want["length"]
    # length only (nothing retained) and the per-node entry point
    eng.set_option(eng.OPT_RETAIN_CLV, 0)
    try:
        assert eng.sankoff_score_tree(ops, ra, rb) == want["length"]
    finally:
        eng.set_option(eng.OPT_RETAIN_CLV, 1)
    op = ops[0]
    sub = eng.sankoff_median_2(int(op["parent"]), int(op["left"]), int(op["right"]))
    assert sub == int((want["vec"][int(op["parent"])].min(axis=1) * w.astype(np.int64)).sum())


def test_sankoff_unit_matrix_is_fitch(eng):
    S, T, N = 4, 24, 5000
    codes = _chars(T, N, S, seed=11).astype(np.uint8)
    tr = tree.random_tree(T, 7)
    ops, ra, rb, rt, n_nodes = tree.schedule(tr)
    M = (1 - np.eye(S)).astype(np.int32)
    eng.sankoff_set_matrix(M)
    eng.sankoff_set_tips(codes, S, capacity=n_nodes)
    eng.fitch_set_tips(codes, S, capacity=n_nodes)
    assert eng.sankoff_score_tree(ops, ra, rb) == eng.fitch_score_tree(ops, ra, rb)


def test_sankoff_errors(eng):
    S, T, N = 4, 6, 50
    codes = _chars(T, N, S, seed=1).astype(np.uint8)
    tr = tree.random_tree(T, 1)
    ops, ra, rb, rt, n_nodes = tree.schedule(tr)
    from phylocaml_b200 import engine

    fresh = engine.Engine(0)
    try:
        with pytest.raises(RuntimeError):
            fresh.sankoff_score_tree(ops, ra, rb)  # nothing loaded
        fresh.sankoff_set_tips(codes, S, capacity=n_nodes)
        with pytest.raises(RuntimeError):
            fresh.sankoff_score_tree(ops, ra, rb)  # no matrix
    finally:
        fresh.close()
    bad = codes.copy()
    bad[2, 3] = 0
    with pytest.raises(RuntimeError):
        eng.sankoff_set_tips(bad, S, capacity=n_nodes)
    with pytest.raises(RuntimeError):
        eng.sankoff_set_matrix(-np.ones((S, S), dtype=np.int32))
