"""CPU tests of the host side of the product: the C-ABI library loads and exports every
symbol include/phylo_engine.h declares, the host-only entry points (eigen-decomposition,
partial reduction) are right, the engine refuses to run without a GPU, and the harness
helpers mirror the reference's conventions. No compute kernels are launched here."""
import ctypes
import os
import re
import subprocess

import numpy as np
import pytest

from phylocaml_b200 import engine, mlmodel, tree

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLD = os.path.join(ROOT, "tests", "golden")


def _declared_symbols():
    with open(os.path.join(ROOT, "include", "phylo_engine.h")) as f:
        src = re.sub(r"/\*.*?\*/", "", f.read(), flags=re.S)
    return sorted(set(re.findall(r"\b(phylo_[a-z0-9_]+)\s*\(", src)))


def test_library_exports_every_declared_symbol(built):
    lib = engine.load()
    declared = _declared_symbols()
    assert len(declared) >= 40
    for name in declared:
        assert hasattr(lib, name), name
    # and the ctypes table binds exactly the declared surface
    assert sorted(engine.SIGNATURES) == declared


def test_library_has_sm100a_code_only(built):
    out = subprocess.run(["cuobjdump", "--list-elf", engine.LIB_PATH], capture_output=True, text=True).stdout
    archs = set(re.findall(r"sm_(\d+a?)", out))
    assert archs == {"100a"}, archs


def test_product_does_not_link_the_oracle(built):
    out = subprocess.run(["ldd", engine.LIB_PATH], capture_output=True, text=True).stdout
    assert "oracle" not in out and "_ref" not in out
    for dirpath, _, files in os.walk(os.path.join(ROOT, "phylocaml_b200")):
        for fn in files:
            if fn.endswith((".py", ".cu", ".cuh", ".cpp", ".h")):
                with open(os.path.join(dirpath, fn)) as f:
                    txt = f.read()
                assert "import oracle" not in txt and "from oracle" not in txt, fn
                assert "liboracle" not in txt and "phylo_oracle.h" not in txt, fn


def test_engine_create_fails_loudly_without_gpu(built):
    lib = engine.load()
    h = ctypes.c_void_p()
    rc = lib.phylo_engine_create(0, ctypes.byref(h))
    if rc == 0:
        lib.phylo_engine_destroy(h)
        pytest.skip("a GPU is present")
    assert rc == -1 and h.value is None
    assert b"no CPU fallback" in lib.phylo_last_error(None)
    with pytest.raises(engine.PhyloError):
        engine.Engine(0)


def test_group_create_fails_loudly_without_gpu(built):
    lib = engine.load()
    h = ctypes.c_void_p()
    devs = (ctypes.c_int * 2)(0, 0)
    rc = lib.phylo_group_create(devs, 2, ctypes.byref(h))
    if rc == 0:
        lib.phylo_group_destroy(h)
        pytest.skip("a GPU is present")
    assert rc == -1 and h.value is None
    assert b"no CPU fallback" in lib.phylo_group_last_error(None)
    with pytest.raises(engine.PhyloError):
        engine.Group([0])
    assert lib.phylo_group_create(devs, 0, ctypes.byref(h)) == -2  # PHYLO_ERR_ARG


def test_alphabet_tables_follow_the_reference():
    """phylocaml_b200/alphabet.py against what the reference pins: test/alphabetTest.ml:12-18
    (1,2,4,8,16 <-> A,C,G,T,- in Alphabet.dna) and the equates of lib/alphabet.ml:301-326."""
    from phylocaml_b200 import alphabet

    d = alphabet.dna_codes()
    assert [d[c] for c in "ACGT-X"] == [1, 2, 4, 8, 16, 32]
    assert [d[c] for c in "01234"] == [1, 2, 4, 8, 16]
    n = alphabet.nucleotides_codes()
    assert [n[c] for c in "ACGT-"] == [1, 2, 4, 8, 16]
    assert n["R"] == 1 | 4 and n["Y"] == 8 | 2 and n["N"] == 15 and n["X"] == 15 and n["?"] == 31
    assert n["1"] == 8 | 16 and n["P"] == 31 and n["E"] == 4 | 8 | 1 | 16
    t = alphabet.nucleotides_table()
    assert t[ord("a")] == t[ord("A")] == 1 and t[ord("Z")] == 0 and t[ord("?")] == 31
    lk = alphabet.nucleotides_table_likelihood()
    assert lk[ord("-")] == 15 and lk[ord("?")] == 15 and lk[ord("8")] == 1 and lk[ord("Z")] == 0
    aa = alphabet.aminoacids_table_likelihood()
    assert aa[ord("A")] == 1 and aa[ord("V")] == 1 << 19 and aa[ord("X")] == (1 << 20) - 1
    assert np.array_equal(alphabet.translate(t, np.frombuffer(b"ACgt-", dtype=np.uint8)), [1, 2, 4, 8, 16])


def test_bench_gpu_local_cpu_binding_is_best_effort(tmp_path):
    """bench.bind_to_gpu_local_cpus: parses sysfs' cpulist, binds only to a proper subset of the
    allowed CPUs, and returns None (without raising) whenever something is missing."""
    import bench

    assert bench.parse_cpulist("0-3,8,10-11\n") == [0, 1, 2, 3, 8, 10, 11]
    assert bench.parse_cpulist("") == []

    class Prop:
        pci_domain_id, pci_bus_id, pci_device_id = 0, 0x1b, 0

    class Cuda:
        @staticmethod
        def get_device_properties(i):
            return Prop()

    class Torch:
        cuda = Cuda()

    before = os.sched_getaffinity(0)
    assert bench.bind_to_gpu_local_cpus(Torch(), 0, sysfs=str(tmp_path)) is None  # no such device directory
    d = tmp_path / "0000:1b:00.0"
    d.mkdir()
    (d / "local_cpulist").write_text(",".join(str(c) for c in sorted(before)))
    assert bench.bind_to_gpu_local_cpus(Torch(), 0, sysfs=str(tmp_path)) is None  # not a proper subset: nothing to do
    (d / "local_cpulist").write_text("100000-100003")
    assert bench.bind_to_gpu_local_cpus(Torch(), 0, sysfs=str(tmp_path)) is None  # no allowed CPU among them
    assert os.sched_getaffinity(0) == before
    if len(before) >= 2:
        one = sorted(before)[0]
        (d / "local_cpulist").write_text(str(one))
        (d / "numa_node").write_text("1\n")
        try:
            assert bench.bind_to_gpu_local_cpus(Torch(), 0, sysfs=str(tmp_path)) is None  # fewer than min_cpus
            note = bench.bind_to_gpu_local_cpus(Torch(), 0, sysfs=str(tmp_path), min_cpus=1)
            assert note is not None and "NUMA node 1" in note and os.sched_getaffinity(0) == {one}
        finally:
            os.sched_setaffinity(0, before)

    class Broken:
        class cuda:
            @staticmethod
            def get_device_properties(i):
                raise RuntimeError("no CUDA")

    assert bench.bind_to_gpu_local_cpus(Broken(), 0) is None


def test_bench_ranks_hold_slices_of_one_global_alignment(monkeypatch):
    """bench.py: for every rank count the shards are contiguous, 1024-aligned, cover [0, N) and
    concatenate to the alignment a single rank scores; the L2 rule flags the small configurations."""
    import bench

    monkeypatch.setattr(engine, "pinned_empty", lambda shape, dtype: np.empty(shape, dtype))
    monkeypatch.setattr(bench, "BASE_PATTERNS", 3000)
    model = mlmodel.create(("JC69",), 4)
    tr = tree.random_tree(6, 2)
    N = 10_000
    whole = bench.build_tips(tree, tr, model, 6, N, 4, 0)
    assert np.array_equal(whole[:, :3000], whole[:, 3000:6000])
    for world in (1, 2, 3, 8):
        bounds = [bench.shard_bounds(N, world, r) for r in range(world)]
        assert bounds[0][0] == 0 and bounds[-1][1] == N
        assert all(a[1] == b[0] for a, b in zip(bounds, bounds[1:]))
        assert all(lo % 1024 == 0 for lo, hi in bounds if hi > lo)
        parts = [bench.build_tips(tree, tr, model, 6, hi - lo, 4, lo) for lo, hi in bounds if hi > lo]
        assert np.array_equal(np.concatenate(parts, axis=1), whole)
    foot = bench.device_footprint
    assert foot("fitch", 64, 4, 1, 1, None, 1_000_000) == 127 * 0.5 * 1_000_000 < 4 * bench.L2_BYTES
    assert foot("dna", 256, 4, 4, 1, "fused", 4_000_000) > 100e9
    assert foot("dna", 256, 4, 4, 1, "fused-lnl", 500_000) < 4 * bench.L2_BYTES


# ------------------------------------------------------------- eigen-decomposition ----
@pytest.mark.parametrize("case", ["irrev4", "irrev20", "pert4", "pert20", "pert61"])
def test_diagonalize_gtr_non_reversible_generators(built, oracle, case):
    """Generators that are not time-reversible but have a real spectrum (Const / m_file / m_custom,
    lib/mlModel.ml:473-520): the general solver (Hessenberg + shifted QR + back-substitution) must
    reproduce the P(t) of the reference's dgeev path (golden vectors made by oracle/_ref)."""
    from scipy.linalg import expm

    g = np.load(os.path.join(GOLD, "compose_nonrev_ref.npz"))
    Q = g[case + "_Q"]
    U, D, Ui = engine.diagonalize(Q, False)
    assert np.abs(U @ Ui - np.eye(Q.shape[0])).max() <= 1e-9
    for t, P in zip(g[case + "_t"], g[case + "_P"]):
        got = oracle.compose(U, D, Ui, t)
        # the triangular generators have an ill-conditioned eigenbasis (cond ~ 7e3 for irrev20): the
        # reference itself is 1.3e-11 away from expm there, ours 2.6e-12
        assert np.abs(got - P).max() <= 1e-10, (case, t)
        assert np.abs(got - expm(Q * t)).max() <= 1e-11, (case, t)


@pytest.mark.parametrize("alpha", [0.02, 0.3, 0.5, 1.0, 2.5, 10.0, 300.0])
def test_gamma_rates_against_scipy(built, alpha):
    """phylo_gamma_rates (self-contained incomplete gamma + inverse) against SciPy, for the two
    definitions: the reference's literal quantiles (lib/mlModel.ml:93-99, r_0 = 0; Pareto/GSL
    there -- parity unpinned by the reference, pinned here to SciPy) and Yang's class means."""
    for k in (1, 2, 4, 8, 16):
        r0, p0 = engine.gamma_rates(alpha, k, "ref_literal")
        w0 = mlmodel.gamma_rates_ref_literal(alpha, k)
        assert r0[0] == 0.0 and np.allclose(p0, 1.0 / k)
        assert np.all(np.abs(r0 - w0) <= 1e-11 * np.abs(w0))
        r1, _ = engine.gamma_rates(alpha, k, "yang_mean")
        w1 = mlmodel.gamma_rates_yang_mean(alpha, k)
        assert np.all(np.abs(r1 - w1) <= 1e-11 * np.abs(w1))
        assert abs(r1.mean() - 1.0) <= 1e-12 and np.all(np.diff(r1) >= 0)
    lib = engine.load()
    assert lib.phylo_gamma_rates(-1.0, 4, 0, None, None) == -2


def _interpret_plan(steps, T):
    """Run a compiled plan on symbolic values: a value is the nested tuple of what it was computed
    from. Returns (root expression, {slot: expression} of every median, deepest stack)."""
    cur, stack, out, deepest = None, [], {}, 0

    def operand(kind, idx):
        nonlocal cur
        if kind == 0:
            assert idx < T
            return ("tip", int(idx))
        if kind == 3:
            assert idx >= T
            return ("stored", int(idx))
        if kind == 1:
            assert cur is not None
            v, cur = cur, None
            return v
        assert kind == 2 and stack
        return stack.pop()

    for lk, li, rk, ri, push, slot in steps.tolist():
        if push:
            assert cur is not None
            stack.append(cur)
            cur = None
            deepest = max(deepest, len(stack))
        # a popped operand is taken before the in-register one is consumed, whichever side it is on
        if lk == 2:
            left = operand(lk, li)
            right = operand(rk, ri)
        else:
            right = operand(rk, ri) if rk == 2 else None
            left = operand(lk, li)
            right = operand(rk, ri) if right is None else right
        assert cur is None, "a live value would be overwritten"
        cur = ("median", left, right)
        if slot >= 0:
            out[int(slot)] = cur
    assert not stack
    return cur, out, deepest


def _meaning(ops, T, ra, rb):
    val = {}
    produced = {int(o["parent"]) for o in ops}

    def get(s):
        s = int(s)
        if s < T:
            return ("tip", s)
        return val[s] if s in val else ("stored", s)

    for o in ops:
        val[int(o["parent"])] = ("median", get(o["left"]), get(o["right"]))
    assert produced == set(val)
    return ("median", get(ra), get(rb)), val


@pytest.mark.parametrize("kind,T,seed", [("random", 5, 1), ("random", 64, 2), ("random", 257, 3), ("caterpillar", 300, 0),
                                         ("random", 2, 4), ("random", 3, 5)])
def test_plan_compiler_preserves_the_schedule(built, kind, T, seed):
    """phylo_plan_compile (the program the tree-fused kernels and the Fitch register walk execute):
    interpreting it on symbolic values reproduces exactly the expression the post-order schedule
    denotes, for every root edge; the stack never exceeds the reported depth, which is
    logarithmic for random trees and 1 for a caterpillar."""
    tr = tree.random_tree(T, seed) if kind == "random" else tree.caterpillar_tree(T)
    edges = tr.edges()
    for e in edges[:: max(1, len(edges) // 7)]:
        ops, ra, rb, rt, n_nodes = tree.schedule(tr, root_edge=e)
        got = engine.plan_compile(ops, T, n_nodes, ra, rb)
        assert got is not None
        steps, depth = got
        root, out, deepest = _interpret_plan(steps, T)
        want_root, want = _meaning(ops, T, ra, rb)
        assert root == want_root
        assert out == want
        assert deepest <= depth <= max(1, int(np.ceil(np.log2(max(T, 2)))) + 1)
        if kind == "caterpillar":
            assert depth == 1


def test_plan_compiler_partial_schedules_and_rejections(built):
    """Incremental re-scoring: a path-to-root schedule whose other operands are resident CLVs
    (kind 3). Non-trees (a result used twice, a slot written twice) are refused -> per-node path."""
    T = 24
    tr = tree.random_tree(T, 12)
    ops, ra, rb, rt, n_nodes = tree.schedule(tr)
    parent_of = {int(o["left"]): j for j, o in enumerate(ops)}
    parent_of.update({int(o["right"]): j for j, o in enumerate(ops)})
    path, j = [4], 4
    while int(ops[j]["parent"]) in parent_of:
        j = parent_of[int(ops[j]["parent"])]
        path.append(j)
    sub = ops[sorted(path)]
    steps, depth = engine.plan_compile(sub, T, n_nodes, ra, rb)
    root, out, deepest = _interpret_plan(steps, T)
    want_root, want = _meaning(sub, T, ra, rb)
    assert root == want_root and out == want and depth == 1
    assert (steps[:, [0, 2]] == 3).any()
    twice = ops.copy()
    twice[-1]["left"] = twice[-2]["left"]  # some result feeds two medians
    assert engine.plan_compile(twice, T, n_nodes, ra, rb) is None
    dup = ops.copy()
    dup[1]["parent"] = dup[0]["parent"]
    assert engine.plan_compile(dup, T, n_nodes, ra, rb) is None


def test_integerize_matrix_is_the_reference_formula(built, oracle):
    """lib/mlModel.ml:639-660 restated with Python ints (int_of_float truncates toward zero), on
    P(t) from the oracle's compose (pinned to the reference's compose_gtr)."""
    import math

    g = np.load(os.path.join(GOLD, "compose_ref.npz"))
    for case, with_priors in (("dna_gtr", True), ("dna_jc69", False), ("aa20", True)):
        Q = g[case + "_Q"]
        n = Q.shape[0]
        U, D = g[case + "_U"], g[case + "_D"]
        Ui = g[case + "_Ui"] if case + "_Ui" in g else None
        pri = np.full(n, 1.0 / n) if not with_priors else np.abs(np.linalg.svd(Q.T)[2][-1]) / np.abs(np.linalg.svd(Q.T)[2][-1]).sum()
        for t in (0.05, 0.5):
            P = oracle.compose(U, D, Ui, t)
            for sigma in (2, 3, 4):
                want = [[-int(10.0 ** sigma * math.log((pri[i] if with_priors else 1.0) * P[i, j])) for j in range(n)]
                        for i in range(n)]
                got = engine.integerize_matrix(P, pri if with_priors else None, sigma)
                assert got.tolist() == want, (case, t, sigma)
    bad = np.eye(4)
    with pytest.raises(engine.PhyloError) as ei:
        engine.integerize_matrix(bad, None, 3)  # ln 0
    assert ei.value.code == -5


def test_non_reversible_model_through_the_pruning_oracle(built, oracle):
    """End to end on the CPU side: a non-reversible 4-state generator, decomposed by the product's
    general solver, scored by the pruning oracle, against brute-force summation over all interior
    states with expm (no eigensystem, no pruning). Without reversibility the likelihood depends on
    where the root sits: the engine's convention is pi at the first end of the root edge and every
    P(t) applied parent -> child away from it, which is what the enumeration does with root = ra."""
    import importlib.util

    spec = importlib.util.spec_from_file_location("make_golden", os.path.join(GOLD, "make_golden.py"))
    mg = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mg)
    g = np.load(os.path.join(GOLD, "compose_nonrev_ref.npz"))
    Q = g["pert4_Q"] / 3.0
    U, D, Ui = engine.diagonalize(Q, False)
    pi = np.array([0.1, 0.2, 0.3, 0.4])  # root distribution: any, it need not be stationary
    rates, probs = np.array([0.3, 1.7]), np.array([0.5, 0.5])
    model = dict(S=4, K=2, Q=Q, U=U, D=D, Ui=Ui, pi=pi, rates=rates, probs=probs, pinvar=None, sym=False)
    tr = tree.random_tree(5, 9, mean_bl=0.4)
    tips = tree.random_tips(5, 12, 4, seed=3, missing_frac=0.1)
    checked = 0
    for e in tr.edges():
        ops, ra, rb, rt, n_nodes = tree.schedule(tr, root_edge=e)
        if ra < 5:
            continue  # root end is a tip: the enumeration roots at interior nodes
        got = oracle.lk_score_tree(model, tips, None, ops, n_nodes, ra, rb, rt)["site_lnl"]
        want = mg.brute_lnl(model, tr, tips, root=ra)
        assert np.abs(got - want).max() <= 1e-11 * np.abs(want).max(), e
        checked += 1
    assert checked >= 2
    # and the root position matters for this model (so the test above is not vacuous)
    vals = []
    for e in tr.edges():
        ops, ra, rb, rt, n_nodes = tree.schedule(tr, root_edge=e)
        vals.append(oracle.lk_score_tree(model, tips, None, ops, n_nodes, ra, rb, rt)["lnl"])
    assert max(vals) - min(vals) > 1e-6


@pytest.mark.parametrize("case", ["cyc3", "cyc4", "cyc7"])
def test_diagonalize_gtr_rejects_complex_spectra_like_the_reference(built, case):
    """lib/mlmodel.c:248-250 raises "Imaginary eigenvalues"; ours returns PHYLO_ERR_NUMERIC."""
    g = np.load(os.path.join(GOLD, "compose_nonrev_ref.npz"))
    with pytest.raises(engine.PhyloError) as ei:
        engine.diagonalize(g[case + "_Q"], False)
    assert ei.value.code == -5


@pytest.mark.parametrize("case", ["dna_gtr", "dna_f81", "aa20", "codon61"])
def test_diagonalize_gtr_reproduces_reference_P(built, oracle, case):
    """phylo_diagonalize_gtr (Jacobi on the symmetrised generator) vs the reference's LAPACK
    path (lib/mlmodel.c:208-262): eigenvector scaling differs, P(t) must not."""
    g = np.load(os.path.join(GOLD, "compose_ref.npz"))
    Q = g[case + "_Q"]
    U, D, Ui = engine.diagonalize(Q, False)
    assert np.abs(U @ D @ Ui - Q).max() <= 1e-13 * np.abs(Q).max() * Q.shape[0]
    assert np.abs(U @ Ui - np.eye(Q.shape[0])).max() <= 1e-12
    for t, P in zip(g[case + "_t"], g[case + "_P"]):
        got = oracle.compose(U, D, Ui, t)
        assert np.abs(got - P).max() <= 1e-12, (case, t)


@pytest.mark.parametrize("case", ["dna_jc69", "dna_k2p", "five_jc69"])
def test_diagonalize_sym_reproduces_reference_P(built, oracle, case):
    g = np.load(os.path.join(GOLD, "compose_ref.npz"))
    Q = g[case + "_Q"]
    U, D, Ui = engine.diagonalize(Q, True)
    assert Ui is None
    assert np.abs(U.T @ D @ U - Q).max() <= 1e-13  # rows of U are the eigenvectors (mlmodel.c:155-158)
    assert np.all(np.diff(np.diag(D)) >= 0)        # ascending like dsyev
    for t, P in zip(g[case + "_t"], g[case + "_P"]):
        assert np.abs(oracle.compose(U, D, None, t) - P).max() <= 1e-12


def test_diagonalize_errors(built):
    Q = np.zeros((4, 4))
    Q[0, 1] = np.nan
    with pytest.raises(engine.PhyloError) as ei:
        engine.diagonalize(Q, False)
    assert ei.value.code == -5
    # a rotation generator has complex eigenvalues: the reference fails ("Imaginary
    # eigenvalues", lib/mlmodel.c:248-250); so do we
    R = np.array([[0.0, 1.0, 0.0], [0.0, 0.0, 1.0], [1.0, 0.0, 0.0]]) - np.eye(3)
    with pytest.raises(engine.PhyloError):
        engine.diagonalize(R, False)


def test_reduce_partials_matches_oracle(built, oracle):
    eng_lib = engine.load()
    rng = np.random.default_rng(0)
    for n in (1, 5, 1024, 1025, 3907, 1024 * 1024 + 3):
        v = rng.normal(-1e4, 50.0, n)
        got = eng_lib.phylo_reduce_partials(v.ctypes.data_as(ctypes.POINTER(ctypes.c_double)), n)
        assert got == oracle.reduce(v), n


# ---------------------------------------------------------------- harness mirrors ----
def test_q_builders_follow_mlmodel_ml():
    pi = np.array([0.1, 0.2, 0.3, 0.4])
    for Q in (mlmodel.m_jc69(4), mlmodel.m_k2p(0.5, 4), mlmodel.m_f81(pi, 4), mlmodel.m_hky85(pi, 2.0, 4),
              mlmodel.m_f84(pi, 1.5, 4), mlmodel.m_tn93(pi, 2.0, 3.0, 4),
              mlmodel.m_gtr(pi, [1.0, 2.5, 0.8, 1.2, 3.0], 4)):
        assert np.abs(Q.sum(1)).max() < 1e-14            # rows sum to 0
    for Q in (mlmodel.m_f81(pi, 4), mlmodel.m_gtr(pi, [1.0, 2.5, 0.8, 1.2, 3.0], 4)):
        assert abs(-(np.diag(Q) * pi).sum() - 1.0) < 1e-14   # mean rate 1 (mlModel.ml:204-213)
        F = pi[:, None] * Q
        assert np.abs(F - F.T).max() < 1e-15                 # reversible: q_ij = c_ij pi_j
    with pytest.raises(ValueError):                          # mlModel.ml:392-395
        mlmodel.m_gtr(pi, [1.0, 2.0], 4)
    p = mlmodel.priors([0.5, 0.5, 0.0, 0.0], 4)              # clamp 1e-13 + renormalise (:697-712)
    assert p.min() > 0 and abs(p.sum() - 1) < 1e-15
    r = mlmodel.gamma_rates_yang_mean(0.5, 4)
    assert abs(r.mean() - 1.0) < 1e-12
    assert mlmodel.gamma_rates_ref_literal(0.5, 4)[0] == 0.0  # mlModel.ml:93-99: p = 0/k


@pytest.mark.parametrize("T", [2, 3, 10, 50])
def test_tree_and_schedule_invariants(T):
    """2T-3 edges / 2T-2 nodes (test/treeTest.ml:52-59, lib/tree.ml:552-568); the schedule
    has T-2 medians in post-order, whatever the root edge."""
    tr = tree.random_tree(T, seed=T)
    assert len(tr.edges()) == 2 * T - 3
    assert len(tr.adj) == 2 * T - 2
    for e in tr.edges()[:5]:
        ops, ra, rb, rt, n_nodes = tree.schedule(tr, root_edge=e)
        assert len(ops) == T - 2 and n_nodes == 2 * T - 2 and rt == tr.length(*e)
        done = set(range(T))
        for op in ops:
            assert int(op["left"]) in done and int(op["right"]) in done and int(op["parent"]) >= T
            done.add(int(op["parent"]))
        assert ra in done and rb in done


def test_shard_bounds_cover_and_align():
    import bench

    for N, world in ((4_000_000, 8), (4_000_000, 3), (1000, 4), (1, 2)):
        spans = [bench.shard_bounds(N, world, r) for r in range(world)]
        assert spans[0][0] == 0 and spans[-1][1] == N
        for (a, b), (c, d) in zip(spans, spans[1:]):
            assert b == c and (b % 1024 == 0 or b == N)


def test_ocaml_stubs_compile_and_cover_the_plugin_surface(tmp_path):
    """stubs/phylo_stubs.c compiles against the (fake) caml headers and defines the stubs the
    reference's externs name (lib/mlmodel.h:39-43) plus one stub per new external in ocaml/."""
    obj = tmp_path / "phylo_stubs.o"
    subprocess.run(["gcc", "-O2", "-Wall", "-Werror", "-fPIC", "-I" + os.path.join(ROOT, "oracle", "shim"),
                    "-I" + os.path.join(ROOT, "include"), "-c", os.path.join(ROOT, "stubs", "phylo_stubs.c"),
                    "-o", str(obj)], check=True)
    syms = subprocess.run(["nm", str(obj)], capture_output=True, text=True).stdout
    defined = set(re.findall(r" T (\w+)", syms))
    for name in ("likelihood_CAML_compose_gtr", "likelihood_CAML_compose_sym",
                 "likelihood_CAML_diagonalize_gtr", "likelihood_CAML_diagonalize_sym"):
        assert name in defined, name
    externs = set()
    for fn in ("likelihood_c.ml", "nonAdditive_c.ml"):
        with open(os.path.join(ROOT, "ocaml", fn)) as f:
            externs |= set(re.findall(r'=\s*"(\w+_CAML_\w+)"', f.read()))
    assert externs and externs <= defined, externs - defined
    # every C-ABI function the stubs call is declared in the header
    called = set(re.findall(r" U (phylo_\w+)", syms))
    assert called and called <= set(_declared_symbols())


def test_host_packers_match_numpy(built):
    """phylo_pack_nibbles / phylo_fitch_pack_planes (the compact upload formats) against numpy."""
    rng = np.random.default_rng(5)
    for N in (1, 2, 31, 32, 33, 1000, 1025):
        tips = rng.integers(1, 16, size=(5, N), dtype=np.uint8)
        got = engine.pack_nibbles(tips)
        pad = np.concatenate([tips, np.full((5, N % 2), 15, np.uint8)], axis=1)
        assert np.array_equal(got, pad[:, 0::2] | (pad[:, 1::2] << 4))
    lib = engine.load()
    assert [lib.phylo_fitch_plane_count(s) for s in (1, 4, 7, 9, 20, 33, 64)] == [1, 4, 8, 12, 24, 64, 64]
    for N, ns, dt in ((70, 4, np.uint8), (64, 6, np.uint8), (33, 13, np.uint16), (100, 20, np.uint32), (40, 61, np.uint64)):
        codes = rng.integers(1, 1 << min(ns, 62), size=(3, N), dtype=np.uint64).astype(dt)
        planes = engine.fitch_pack_planes(codes, ns)
        NP = lib.phylo_fitch_plane_count(ns)
        W = (N + 31) // 32
        assert planes.shape == (3, W * NP)
        p3 = planes.reshape(3, W, NP)
        for t in range(3):
            for i in range(N):
                v = sum(((int(p3[t, i // 32, s]) >> (i % 32)) & 1) << s for s in range(NP))
                assert v == int(codes[t, i]) & ((1 << ns) - 1)


def test_fitch_reroot_keeps_the_length_and_halves_the_height(oracle):
    """phylo_fitch_reroot (what a length-only phylo_fitch_score_tree evaluates): the same unrooted tree, every
    interior node with exactly one median, in post-order, the oracle's Fitch length unchanged, and the chain
    of dependent medians no longer than before (much shorter for a caterpillar)."""
    def height(ops, ra, rb):
        h = {}
        for op in ops:
            h[int(op["parent"])] = 1 + max(h.get(int(op["left"]), 0), h.get(int(op["right"]), 0))
        return max(h.get(ra, 0), h.get(rb, 0))

    cases = [(tree.random_tree(T, seed), T) for T, seed in ((5, 1), (16, 2), (64, 1), (64, 7), (33, 3))]
    cases += [(tree.caterpillar_tree(40), 40)]
    for tr, T in cases:
        ops, ra, rb, rt, n_nodes = tree.schedule(tr)
        ops2, ra2, rb2 = engine.fitch_reroot(ops, n_nodes, ra, rb)
        assert sorted(int(p) for p in ops2["parent"]) == sorted(int(p) for p in ops["parent"])
        done = set(range(T))
        for op in ops2:
            assert int(op["left"]) in done and int(op["right"]) in done and int(op["parent"]) not in done
            done.add(int(op["parent"]))
        assert ra2 in done and rb2 in done and ra2 != rb2
        chars = tree.random_fitch_chars(T, 1500, 4, seed=T, dtype=np.uint8)
        want = oracle.fitch_score_tree(chars, None, ops, n_nodes, ra, rb)["length"]
        assert oracle.fitch_score_tree(chars, None, ops2, n_nodes, ra2, rb2)["length"] == want
        assert height(ops2, ra2, rb2) <= height(ops, ra, rb)
    tr = tree.caterpillar_tree(40)
    ops, ra, rb, rt, n_nodes = tree.schedule(tr)
    ops2, ra2, rb2 = engine.fitch_reroot(ops, n_nodes, ra, rb)
    assert height(ops2, ra2, rb2) <= height(ops, ra, rb) // 2 + 1

