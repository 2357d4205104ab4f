"""The OCaml boundary without OCaml: stubs/phylo_stubs.c compiled against the fake caml headers
(oracle/shim) and driven from C like the plugin bodies in ocaml/ drive it (tests/c/stub_lifetime.c):
1000 successive trees through Likelihood_c.median_2 / NonAdditive_c.median_2 on 2*T-slot engines,
node custom blocks finalized by a toy GC, slots and device buffers reused, engines reference
counted. Plus the slot allocator and the new node-level entry points through the C ABI."""
import os
import subprocess

import numpy as np
import pytest

from helpers import dna_gtr_g4, rel_err, setup_lk
from phylocaml_b200 import engine, tree

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_stub_lifetime_1000_trees_on_a_2T_slot_engine(built, tmp_path):
    exe = tmp_path / "stub_lifetime"
    lib = os.path.join(ROOT, "phylocaml_b200", "lib")
    subprocess.run(["gcc", "-O2", "-Wall", "-Werror", "-I" + os.path.join(ROOT, "oracle", "shim"),
                    "-I" + os.path.join(ROOT, "include"), "-o", str(exe),
                    os.path.join(ROOT, "tests", "c", "stub_lifetime.c"), os.path.join(ROOT, "stubs", "phylo_stubs.c"),
                    os.path.join(ROOT, "oracle", "shim", "caml_shim.c"), "-L" + lib, "-lphyloc_b200",
                    "-Wl,-rpath," + lib, "-lm"], check=True)
    r = subprocess.run([str(exe)], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stdout + r.stderr
    assert r.stdout.startswith("OK"), r.stdout


def test_lk_slot_allocator_reuses_slots_and_buffers(eng, oracle):
    model = dna_gtr_g4()
    tr, ops, ra, rb, rt, n_nodes, tips = setup_lk(6, 700, model, seed=3)
    T = 6
    eng.lk_set_model(model)
    eng.lk_set_tips(tips, capacity=2 * T)
    assert eng.node_stats() == (T, 0, 0)
    want = oracle.lk_score_tree(model, tips, None, ops, n_nodes, ra, rb, rt)["lnl"]
    for rep in range(40):
        slots = {}
        for op in ops:
            s, g = eng.node_alloc()
            slots[int(op["parent"])] = (s, g)
            m = lambda i: i if i < T else slots[i][0]
            eng.lk_median_2(s, m(int(op["left"])), op["t_left"], m(int(op["right"])), op["t_right"])
        m = lambda i: i if i < T else slots[i][0]
        lnl = eng.lk_edge_lnl(m(ra), m(rb), [rt])[0]
        assert rel_err(lnl, want) <= 1e-12
        for s, g in slots.values():
            eng.node_release(s, g)
    cap, used, bufs = eng.node_stats()
    assert (cap, used) == (T, 0) and bufs <= T - 2
    # every slot live: the table grows, nothing is lost
    held = [eng.node_alloc() for _ in range(3 * T)]
    assert len({s for s, _ in held}) == 3 * T and eng.node_stats()[0] >= 3 * T
    with pytest.raises(engine.PhyloError):
        eng.node_release(held[0][0] + 1000, held[0][1])
    eng.node_release(*held[0])
    with pytest.raises(engine.PhyloError):
        eng.node_release(*held[0])  # double release
    # a new alignment shape drops every slot: late releases of the old generation are ignored
    eng.lk_set_tips(tips[:, :500], capacity=2 * T)
    eng.node_release(*held[1])
    assert eng.node_stats() == (T, 0, 0)


def test_lk_median_3_is_the_product_of_three_directed_clvs(eng, oracle):
    model = dna_gtr_g4()
    tr, ops, ra, rb, rt, n_nodes, tips = setup_lk(9, 1500, model, seed=5)
    eng.lk_set_model(model)
    eng.lk_set_tips(tips, capacity=n_nodes + 4)
    eng.lk_score_tree(ops, ra, rb, rt)
    # the root-side node ra (if interior) sees its two children and rb across the root edge
    node = ra if ra >= 9 else rb
    other = rb if node == ra else ra
    op = [o for o in ops if int(o["parent"]) == node][0]
    dst = n_nodes
    eng.lk_median_3(dst, int(op["left"]), op["t_left"], int(op["right"]), op["t_right"], other, rt)
    clv3, sc3 = eng.lk_get_clv(dst)
    clvn, scn = eng.lk_get_clv(node)
    want = oracle.lk_score_tree(model, tips, None, ops, n_nodes, ra, rb, rt, want_clv=True)
    P = np.stack([oracle.compose(model["U"], model["D"], model["Ui"], rt * r) for r in model["rates"]])
    if other < 9:
        Lo = np.repeat(((tips[other][:, None] >> np.arange(4)[None, :]) & 1).astype(float)[:, None, :], 4, axis=1)
        sco = np.zeros(tips.shape[1], dtype=np.int64)
    else:
        Lo, sco = want["clv"][other], want["scale"][other].astype(np.int64)
    y = np.einsum("kij,skj->ski", P, Lo)
    ref = want["clv"][node] * y
    refsc = want["scale"][node].astype(np.int64) + sco
    # compare in log space (the engine may have rescaled once more)
    got_log = np.log(clv3.reshape(len(sc3), -1).max(1)) - 256 * np.log(2.0) * sc3
    ref_log = np.log(ref.reshape(len(refsc), -1).max(1)) - 256 * np.log(2.0) * refsc
    assert np.abs(got_log - ref_log).max() <= 1e-10 * np.abs(ref_log).max()
    # the site likelihood sum_k p_k sum_i pi_i L3[k,i] equals the tree's site lnL
    site = np.log(np.einsum("k,i,ski->s", model["probs"], model["pi"], clv3)) - 256 * np.log(2.0) * sc3
    assert np.abs(site - want["site_lnl"]).max() <= 1e-10 * np.abs(want["site_lnl"]).max()


def test_fitch_median_3_equals_the_uppass_and_eltcount(eng, oracle):
    T, N = 12, 5000
    tr = tree.random_tree(T, seed=4)
    ops, ra, rb, rt, n_nodes = tree.schedule(tr)
    chars = tree.random_fitch_chars(T, N, 4, seed=6, ambiguity=0.2)
    eng.fitch_set_tips(chars, 4, capacity=n_nodes + 8)
    eng.fitch_score_tree(ops, ra, rb)
    eng.fitch_uppass(ops, ra, rb)
    parent_of = {}
    for op in ops:
        parent_of[int(op["left"])] = int(op["parent"])
        parent_of[int(op["right"])] = int(op["parent"])
    # a node two levels below the root edge: its parent's FINAL sets must first become a node value
    cand = [op for op in ops if int(op["parent"]) in parent_of and parent_of[int(op["parent"])] not in (ra, rb)
            and parent_of[int(op["parent"])] in parent_of]
    assert cand
    op = cand[0]
    node, par = int(op["parent"]), parent_of[int(op["parent"])]
    fin_par = eng.fitch_get_states(par, final=True)
    s_par, g1 = eng.node_alloc(fitch=True)
    eng.fitch_set_states(s_par, fin_par)
    s_out, g2 = eng.node_alloc(fitch=True)
    eng.fitch_median_3(s_out, node, s_par, int(op["left"]), int(op["right"]))
    assert np.array_equal(eng.fitch_get_states(s_out), eng.fitch_get_states(node, final=True))
    st = eng.fitch_get_states(s_out)
    for i in (0, 1, 31, 32, N - 1):
        assert eng.bv_eltcount(s_out, i) == bin(int(st[i])).count("1")
    eng.node_release(s_par, g1, fitch=True)
    eng.node_release(s_out, g2, fitch=True)


def test_fitch_reupload_invalidates_interior_sets(eng, oracle):
    """after a same-shape re-upload the interior buffers hold the previous alignment's
    sets; they must not be accepted as resident operands."""
    T, N = 8, 3000
    tr = tree.random_tree(T, seed=2)
    ops, ra, rb, rt, n_nodes = tree.schedule(tr)
    chars = tree.random_fitch_chars(T, N, 4, seed=1)
    eng.fitch_set_tips(chars, 4, capacity=n_nodes)
    eng.fitch_score_tree(ops, ra, rb)
    interior = int(ops[-1]["parent"])
    eng.fitch_get_states(interior)
    eng.fitch_set_tips(tree.random_fitch_chars(T, N, 4, seed=9), 4, capacity=n_nodes)
    with pytest.raises(engine.PhyloError):
        eng.fitch_get_states(interior)
    with pytest.raises(engine.PhyloError):
        eng.fitch_distance(interior, 0)
    with pytest.raises(engine.PhyloError):
        eng.fitch_score_tree(ops[-1:], ra, rb)  # its children are stale too
    assert eng.fitch_score_tree(ops, ra, rb) == oracle.fitch_score_tree(
        tree.random_fitch_chars(T, N, 4, seed=9), None, ops, n_nodes, ra, rb)["length"]


def test_scoring_on_a_non_blocking_stream(built, oracle):
    """phylo_engine_set_stream with a cudaStreamNonBlocking stream (what torch hands out): set-up
    memsets and copies are ordered on that stream, so results do not change."""
    import torch

    model = dna_gtr_g4()
    tr, ops, ra, rb, rt, n_nodes, tips = setup_lk(10, 3000, model, seed=8)
    want = oracle.lk_score_tree(model, tips, None, ops, n_nodes, ra, rb, rt)["lnl"]
    e = engine.Engine(0)
    try:
        s = torch.cuda.Stream()
        e.set_stream(s.cuda_stream)
        e.lk_set_model(model)
        for fused in (1, 0):
            e.set_option(e.OPT_FUSED_TREE, fused)
            for _ in range(3):
                e.lk_set_tips(tips, capacity=n_nodes)
                assert rel_err(e.lk_score_tree(ops, ra, rb, rt), want) <= 1e-12
        chars = tree.random_fitch_chars(10, 5000, 4, seed=3)
        fw = oracle.fitch_score_tree(chars, None, ops, n_nodes, ra, rb)["length"]
        for walk in (1, 3, 2, 0):
            e.set_option(e.OPT_FITCH_WALK, walk)
            e.fitch_set_tips(chars, 4, capacity=n_nodes)
            assert e.fitch_score_tree(ops, ra, rb) == fw
        e.sync()
    finally:
        e.close()
