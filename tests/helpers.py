"""Shared builders for the parity tests (seeded, deterministic)."""
import numpy as np

from phylocaml_b200 import mlmodel, tree

GTR_CO = [1.0, 2.5, 0.8, 1.2, 3.0]  # last exchangeability fixed to 1.0 (lib/mlModel.ml:396-400)
GTR_PI = [0.30, 0.20, 0.25, 0.25]


def dna_gtr_g4(rates="yang_mean", alpha=0.5, pinvar=None):
    sv = ("gamma", 4, alpha) if pinvar is None else ("theta", 4, alpha, pinvar)
    return mlmodel.create(("GTR", GTR_CO), 4, pi=GTR_PI, site_var=sv, rates=rates)


def aa_model(K=4, seed=4):
    R, pi = mlmodel.synthetic_reversible(20, seed)
    return mlmodel.create(("Const", R), 20, pi=pi, site_var=("gamma", K, 0.7) if K > 1 else None)


def codon_model(seed=5):
    R, pi = mlmodel.gy94(2.0, 0.5, seed)
    return mlmodel.create(("Const", R), 61, pi=pi)


def mask_dtype(S):
    return np.uint8 if S <= 8 else (np.uint32 if S <= 32 else np.uint64)


def setup_lk(T, N, model, seed=1, tree_kind="random", mean_bl=0.1, missing=0.01):
    tr = tree.random_tree(T, seed, mean_bl=mean_bl) if tree_kind == "random" else tree.caterpillar_tree(T, mean_bl)
    ops, ra, rb, rt, n_nodes = tree.schedule(tr)
    tips = tree.evolve_tips(tr, model, N, seed + 2, missing_frac=missing, dtype=mask_dtype(model["S"]))
    return tr, ops, ra, rb, rt, n_nodes, tips


def rel_err(a, b):
    return abs(a - b) / max(abs(b), 1e-300)


def dna_gtr_g4_numpy(diag):
    """Same DNA GTR+G4 record as dna_gtr_g4() but with the eigensystem from `diag` (e.g.
    numpy LAPACK), so that oracle-only tests do not depend on the product's own solver."""
    pi = mlmodel.priors(GTR_PI, 4)
    Q = mlmodel.m_gtr(pi, GTR_CO, 4)
    U, D, Ui = diag(Q, False)
    return dict(S=4, K=4, Q=Q, U=U, D=D, Ui=Ui, pi=pi, rates=mlmodel.gamma_rates_yang_mean(0.5, 4),
                probs=np.full(4, 0.25), pinvar=None)
