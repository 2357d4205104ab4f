#!/usr/bin/env python
"""Generates the golden fixtures in tests/golden/ (run here, where /root/reference exists).

Sources of truth, none of which is our oracle restatement or the CUDA engine:
  * the reference's own C compiled unmodified (oracle/_ref): diagonalize_gtr/_sym,
    compose_gtr/_sym (lib/mlmodel.c), bv_fitch, bv_distance, bv_union, bv_inter, bv_popcount,
    bv_saturation, bv_poly_saturation, bv_compare (lib/bitvector/bv.c);
  * the Fitch truth table the reference's own test asserts for every pair of Alphabet.dna
    codes (test/costMatrixTest.ml:83-108: median = inter if non-empty else union, cost 0/1;
    codes A=1,C=2,G=4,T=8,-=16 from test/alphabetTest.ml:12-18, X=32 lib/alphabet.ml:301-307);
  * independent mathematics for what the reference never wrote: brute-force summation over
    all internal-state assignments with scipy.linalg.expm (lnL), brute-force minimum-change
    search (Fitch length and MPR sets), the JC69 two-taxon closed form.

    python tests/golden/make_golden.py      # rewrites tests/golden/*.npz, *.json
"""
import itertools
import json
import os
import sys

import numpy as np
import scipy.linalg as sl

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)

from oracle.oracle import Ref  # noqa: E402
from phylocaml_b200 import mlmodel, tree  # noqa: E402

TS = [1e-4, 0.01, 0.1, 0.5, 2.0, 100.0, -1.0, 0.0, 1e-11]


def q_dna_gtr():
    return mlmodel.m_gtr(np.array([0.30, 0.20, 0.25, 0.25]), [1.0, 2.5, 0.8, 1.2, 3.0], 4)


def golden_compose(ref):
    out = {}
    cases = {"dna_gtr": (q_dna_gtr(), False), "dna_f81": (mlmodel.m_f81(np.array([.1, .2, .3, .4]), 4), False),
             "dna_jc69": (mlmodel.m_jc69(4), True), "dna_k2p": (mlmodel.m_k2p(0.4, 4), True),
             "five_jc69": (mlmodel.m_jc69(5), True)}
    R, pi = mlmodel.synthetic_reversible(20, 4)
    cases["aa20"] = (mlmodel.m_file(pi, R), False)
    R, pi = mlmodel.gy94(2.0, 0.5, 5)
    cases["codon61"] = (mlmodel.m_file(pi, R), False)
    for name, (Q, sym) in cases.items():
        U, D, Ui = ref.diagonalize(Q, sym)
        out[name + "_Q"] = Q
        out[name + "_U"] = U
        out[name + "_D"] = D
        if Ui is not None:
            out[name + "_Ui"] = Ui
        out[name + "_t"] = np.array(TS)
        out[name + "_P"] = np.stack([ref.compose(U, D, Ui, t) for t in TS])
    np.savez_compressed(os.path.join(HERE, "compose_ref.npz"), **out)
    return sorted(cases)


def nonreversible_cases():
    """Generators that are NOT time-reversible but have a real spectrum (what Const / m_file /
    m_custom can hand to diagonalize_gtr, lib/mlModel.ml:473-520), and cyclic ones whose spectrum
    is complex (the reference raises "Imaginary eigenvalues", lib/mlmodel.c:248-250)."""
    rng = np.random.default_rng(7)

    def gen(R):
        R = np.array(R, dtype=float)
        np.fill_diagonal(R, 0.0)
        Q = R.copy()
        np.fill_diagonal(Q, -R.sum(1))
        return Q

    cases = {"irrev4": gen(np.triu(rng.uniform(0.1, 2, (4, 4)), 1)),
             "irrev20": gen(np.triu(rng.uniform(0.1, 2, (20, 20)), 1))}
    for n in (4, 20, 61):
        R, pi = mlmodel.synthetic_reversible(n, 3)
        cases["pert%d" % n] = gen(np.array(R) * np.asarray(pi)[None, :] * (1 + 0.02 * rng.uniform(-1, 1, (n, n))))
    complex_cases = {}
    for n in (3, 4, 7):
        R = np.zeros((n, n))
        for i in range(n):
            R[i, (i + 1) % n] = 1.0
        complex_cases["cyc%d" % n] = gen(R)
    return cases, complex_cases


def golden_compose_nonreversible(ref):
    cases, complex_cases = nonreversible_cases()
    out = {}
    ts = [0.01, 0.3, 2.0]
    for name, Q in cases.items():
        U, D, Ui = ref.diagonalize(Q, False)
        out[name + "_Q"] = Q
        out[name + "_t"] = np.array(ts)
        out[name + "_P"] = np.stack([ref.compose(U, D, Ui, t) for t in ts])
        assert max(np.abs(P - sl.expm(Q * t)).max() for P, t in zip(out[name + "_P"], ts)) < 1e-10
    for name, Q in complex_cases.items():
        try:
            ref.diagonalize(Q, False)
            raise AssertionError("the reference accepted a complex spectrum: " + name)
        except RuntimeError as ex:
            assert "Imaginary" in str(ex)
        out[name + "_Q"] = Q
    np.savez_compressed(os.path.join(HERE, "compose_nonrev_ref.npz"), **out)
    return sorted(cases), sorted(complex_cases)


def golden_bv(ref):
    out = {}
    rng = np.random.default_rng(20261017)
    for w, dt in ((8, np.uint8), (16, np.uint16), (32, np.uint32), (64, np.uint64)):
        n = 997
        hi = min(w, 62)
        a = rng.integers(1, 1 << hi, n, dtype=np.uint64).astype(dt)
        b = rng.integers(1, 1 << hi, n, dtype=np.uint64).astype(dt)
        # make intersections common: copy / overlap a third of the entries
        b[::3] = a[::3]
        b[1::7] |= a[1::7]
        c, cost = ref.bv_fitch(a, b)
        out["w%d_a" % w], out["w%d_b" % w], out["w%d_fitch" % w] = a, b, c
        out["w%d_scalars" % w] = np.array(
            [cost, ref.bv_distance(a, b), ref.bv_popcount(a), ref.bv_saturation(a, 0b101),
             ref.bv_poly_saturation(a, 2), ref.bv_compare(a, b) & 0xffffffff, ref.bv_compare(a, a) & 0xffffffff],
            dtype=np.uint64)
        out["w%d_union" % w] = ref.bv_binop("union", a, b)
        out["w%d_inter" % w] = ref.bv_binop("inter", a, b)
    # the reference's own enabled test: all ordered pairs of the Alphabet.dna codes
    codes = np.array([1, 2, 4, 8, 16, 32], dtype=np.uint8)
    # extended with every polymorphic combination of A,C,G,T,- (Alphabet.dna combinations)
    allc = np.arange(1, 64, dtype=np.uint8)
    for name, cs in (("dna36", codes), ("dna_all", allc)):
        x = np.repeat(cs, len(cs))
        y = np.tile(cs, len(cs))
        med, _ = ref.bv_fitch(x, y)
        cst = np.array([ref.bv_fitch(x[i:i + 1], y[i:i + 1])[1] for i in range(len(x))], dtype=np.uint8)
        # the rule asserted by test/costMatrixTest.ml:88-94
        exp_med = np.where((x & y) != 0, x & y, x | y)
        exp_cst = ((x & y) == 0).astype(np.uint8)
        assert np.array_equal(med, exp_med) and np.array_equal(cst, exp_cst)
        out[name + "_x"], out[name + "_y"], out[name + "_median"], out[name + "_cost"] = x, y, med, cst
    np.savez_compressed(os.path.join(HERE, "bv_ref.npz"), **out)


def brute_lnl(model, tr, tips, root=None):
    """sum over all joint assignments of interior-node states (no pruning, no oracle)."""
    T = tr.T
    interior = sorted(v for v in tr.adj if v >= T)
    S, K = model["S"], model["K"]
    Q = model["Q"]
    edges = tr.edges()
    N = tips.shape[1]
    site = np.zeros(N)
    root = interior[0] if root is None else root
    for s in range(N):
        tot = 0.0
        for k in range(K):
            # symmetric models go through compose_sym, which takes `const float t`
            # (lib/mlmodel.c:280): the reference semantics include that rounding
            tau = lambda e: (float(np.float32(tr.blen[e] * model["rates"][k])) if model["sym"]
                             else tr.blen[e] * model["rates"][k])
            Pe = {e: sl.expm(Q * tau(e)) for e in edges}
            lk = 0.0
            for assign in itertools.product(range(S), repeat=len(interior)):
                st = dict(zip(interior, assign))
                # orient edges away from `root`
                p = model["pi"][st[root]]
                stack = [(None, root)]
                while stack and p != 0.0:
                    par, cur = stack.pop()
                    for nb in tr.adj[cur]:
                        if nb == par:
                            continue
                        P = Pe[(min(cur, nb), max(cur, nb))]
                        if nb < T:
                            m = int(tips[nb, s])
                            p *= sum(P[st[cur], j] for j in range(S) if (m >> j) & 1)
                        else:
                            p *= P[st[cur], st[nb]]
                            stack.append((cur, nb))
                lk += p
            tot += model["probs"][k] * lk
        site[s] = np.log(tot)
    return site


def strip(model):
    return {k: (v.tolist() if isinstance(v, np.ndarray) else v) for k, v in model.items()
            if k in ("S", "K", "Q", "pi", "rates", "probs", "pinvar", "sym")}


def golden_lnl():
    """Brute-force site log-likelihoods on tiny trees. The model is stored by Q (not by an
    eigensystem) so each consumer diagonalises with its own solver."""
    cases = []
    specs = [
        ("gtr_g2_T4", ("GTR", [1.0, 2.5, 0.8, 1.2, 3.0]), [0.3, 0.2, 0.25, 0.25], ("custom", [0.3, 1.7], [0.5, 0.5]), 4, 4),
        ("hky_T5", ("HKY85", 2.0), [0.1, 0.2, 0.3, 0.4], None, 5, 4),
        ("jc_g3_T5", ("JC69",), None, ("custom", [0.2, 1.0, 1.8], [1 / 3, 1 / 3, 1 / 3]), 5, 4),
        ("f81_5state_T4", ("F81",), [0.2, 0.2, 0.25, 0.25, 0.1], None, 4, 5),
    ]
    for name, subst, pi, sv, T, S in specs:
        q = {"GTR": lambda: mlmodel.m_gtr(np.array(pi), subst[1], S),
             "HKY85": lambda: mlmodel.m_hky85(np.array(pi), subst[1], S),
             "JC69": lambda: mlmodel.m_jc69(S), "F81": lambda: mlmodel.m_f81(np.array(pi), S)}[subst[0]]()
        pri = mlmodel.priors(pi, S)
        rates, probs = (np.array([1.0]), np.array([1.0])) if sv is None else (np.array(sv[1]), np.array(sv[2]))
        model = dict(S=S, K=len(rates), Q=q, pi=pri, rates=rates, probs=probs, pinvar=None,
                     sym=subst[0] in ("JC69",))
        tr = tree.random_tree(T, seed=100 + T)
        rng = np.random.default_rng(7 + T)
        N = 12
        st = rng.integers(0, S, size=(T, N))
        tips = (1 << st).astype(np.uint8)
        tips[0, 0] = (1 << S) - 1          # missing
        tips[1, 1] = 0b0101                # ambiguity
        site = brute_lnl(model, tr, tips)
        cases.append(dict(name=name, model=strip(model), T=T, adj={str(k): v for k, v in tr.adj.items()},
                          blen=[[a, b, l] for (a, b), l in tr.blen.items()], tips=tips.tolist(),
                          site_lnl=site.tolist()))
    # JC69 two taxa, closed form: P_same = 1/4 + 3/4 e^{-4t/3}, P_diff = 1/4 - 1/4 e^{-4t/3}
    t = 0.37
    same = np.log(0.25 * (0.25 + 0.75 * np.exp(-4 * t / 3)))
    diff = np.log(0.25 * (0.25 - 0.25 * np.exp(-4 * t / 3)))
    cases.append(dict(name="jc69_two_taxa_closed_form", t=t, site_lnl_same=float(same), site_lnl_diff=float(diff)))
    with open(os.path.join(HERE, "lnl_bruteforce.json"), "w") as f:
        json.dump(cases, f, indent=1)


def brute_fitch(tr, chars):
    """minimum number of changes over all interior assignments, and the MPR state sets."""
    T = tr.T
    interior = sorted(v for v in tr.adj if v >= T)
    edges = tr.edges()
    N = chars.shape[1]
    nst = int(chars.max()).bit_length()
    length = np.zeros(N, dtype=np.int64)
    mpr = np.zeros((len(interior), N), dtype=np.uint8)
    for s in range(N):
        best, sets = None, None
        tip_opts = [[j for j in range(nst) if (int(chars[t, s]) >> j) & 1] for t in range(T)]
        for assign in itertools.product(range(nst), repeat=len(interior)):
            st = dict(zip(interior, assign))
            cost = 0
            for a, b in edges:
                if a < T and b < T:
                    cost += 0 if set(tip_opts[a]) & set(tip_opts[b]) else 1
                elif a < T:
                    cost += 0 if st[b] in tip_opts[a] else 1
                elif b < T:
                    cost += 0 if st[a] in tip_opts[b] else 1
                else:
                    cost += st[a] != st[b]
            if best is None or cost < best:
                best, sets = cost, [0] * len(interior)
            if cost == best:
                for i, v in enumerate(interior):
                    sets[i] |= 1 << st[v]
        length[s] = best
        mpr[:, s] = sets
    return length, interior, mpr


def golden_fitch():
    cases = []
    for T, seed in ((4, 1), (5, 2), (6, 3), (6, 4)):
        tr = tree.random_tree(T, seed=200 + seed)
        rng = np.random.default_rng(seed)
        N = 40
        chars = (1 << rng.integers(0, 4, size=(T, N))).astype(np.uint8)
        amb = rng.random((T, N)) < 0.15
        chars[amb] |= (1 << rng.integers(0, 4, size=(T, N)))[amb].astype(np.uint8)
        length, interior, mpr = brute_fitch(tr, chars)
        cases.append(dict(T=T, adj={str(k): v for k, v in tr.adj.items()},
                          blen=[[a, b, l] for (a, b), l in tr.blen.items()], chars=chars.tolist(),
                          char_length=length.tolist(), interior=interior, mpr=mpr.tolist()))
    with open(os.path.join(HERE, "fitch_bruteforce.json"), "w") as f:
        json.dump(cases, f)


if __name__ == "__main__":
    ref = Ref()
    print("compose:", golden_compose(ref))
    print("compose, non-reversible:", golden_compose_nonreversible(ref))
    golden_bv(ref)
    golden_lnl()
    golden_fitch()
    for fn in sorted(os.listdir(HERE)):
        print(fn, os.path.getsize(os.path.join(HERE, fn)))
