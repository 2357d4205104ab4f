"""Harness-side mirror of MlModel.create (lib/mlModel.ml:664-746): builds the model record
the engine consumes (S, K, U, D, Ui, pi, rates, probs, pinvar).

In a phylocaml deployment this record comes from the (unchanged) OCaml MlModel module; this
module only exists so tests and bench.py can make the same records without OCaml. The
eigen-decomposition goes through the product's own phylo_diagonalize_* (engine.diagonalize),
i.e. the replacement of the externs at lib/mlModel.ml:73-79.
"""
import numpy as np

MINIMUM = 1e-13  # lib/mlModel.ml:8


def m_meanrate(srm, pi):
    """lib/mlModel.ml:204-213: divide by the mean rate sum_i pi_i * -q_ii."""
    mr = 0.0
    for i in range(srm.shape[0]):
        mr += (-srm[i, i]) * pi[i]
    return srm / mr


def _fill_diag(srm):
    for i in range(srm.shape[0]):
        srm[i, i] = 0.0
        srm[i, i] = -srm[i].sum()
    return srm


def m_jc69(a_size):
    """lib/mlModel.ml:215-243 (no gap state)."""
    srm = np.full((a_size, a_size), 1.0)
    for i in range(a_size):
        srm[i, i] = -1.0 * (a_size - 1)
    return m_meanrate(srm, np.full(a_size, 1.0 / a_size))


def m_k2p(beta, a_size=4):
    """lib/mlModel.ml:246-290 (no gap state): transitions A<->G, C<->T at rate alpha=1."""
    assert a_size in (4, 5)
    beta = max(beta, MINIMUM)
    srm = np.full((a_size, a_size), beta)
    srm[1, 3] = srm[3, 1] = srm[2, 0] = srm[0, 2] = 1.0
    diag = -1.0 - 2 * beta if a_size == 4 else -1.0 - 3.0 * beta
    for i in range(a_size):
        srm[i, i] = diag
    return m_meanrate(srm, np.full(a_size, 1.0 / a_size))


def m_tn93(pi, alpha, beta, a_size=4):
    """lib/mlModel.ml:293-326 (no gap state)."""
    assert a_size in (4, 5)
    srm = np.full((a_size, a_size), 1.0)
    srm[0, 2] = srm[2, 0] = alpha
    srm[1, 3] = srm[3, 1] = beta
    srm = srm * np.asarray(pi)[None, :]
    return m_meanrate(_fill_diag(srm), pi)


def m_f81(pi, a_size):
    """lib/mlModel.ml:329-358 (no gap state)."""
    srm = np.tile(np.asarray(pi, dtype=float)[None, :], (a_size, 1))
    return m_meanrate(_fill_diag(srm), pi)


def m_hky85(pi, kappa, a_size=4):
    """lib/mlModel.ml:361-362."""
    return m_tn93(pi, kappa, kappa, a_size)


def m_f84(pi, kappa, a_size=4):
    """lib/mlModel.ml:365-370."""
    y = pi[1] + pi[3]
    r = pi[0] + pi[2]
    return m_tn93(pi, 1.0 + kappa / r, 1.0 + kappa / y, a_size)


def m_gtr(pi, co, a_size):
    """lib/mlModel.ml:390-421: co = upper triangle row-major WITHOUT the last entry (fixed
    to 1.0, :396-400); q_ij = c_ij * pi_j."""
    need = ((a_size + 1) * (a_size - 2)) // 2
    if len(co) != need:
        raise ValueError("Length of GTR parameters is incorrect: expected %d, got %d" % (need, len(co)))
    co = list(co) + [1.0]
    srm = np.zeros((a_size, a_size))
    n = 0
    for i in range(a_size):
        for j in range(i + 1, a_size):
            srm[i, j] = co[n] * pi[j]
            srm[j, i] = co[n] * pi[i]
            n += 1
    return m_meanrate(_fill_diag(srm), pi)


def m_file(pi, f_rr):
    """lib/mlModel.ml:473-488: arbitrary rate matrix (the only route to WAG/LG/GY94 in the
    reference): off-diagonals as given, diagonal recomputed, mean-rate normalised."""
    srm = np.array(f_rr, dtype=float)
    return m_meanrate(_fill_diag(srm), pi)


def priors(base, a_size):
    """lib/mlModel.ml:697-712: Equal, or Empirical clamped at 1e-13 and renormalised."""
    if base is None:
        return np.full(a_size, 1.0 / a_size)
    p = np.maximum(np.asarray(base, dtype=float), MINIMUM)
    s = p.sum()
    if abs(s - 1.0) >= np.finfo(float).eps:  # Internal.(=.) lib/internal.ml:12
        p = p / s
    if len(p) != a_size:
        raise ValueError("Priors (length %d) don't match alphabet (length %d)" % (len(p), a_size))
    return p


def gamma_rates_ref_literal(alpha, k):
    """lib/mlModel.ml:93-99 + :679 literally: quantiles at p = i/k of Gamma(shape=alpha,
    scale=alpha); r_0 = 0 (PARITY UNPINNED: Pareto/GSL is absent; scipy.stats stands in)."""
    from scipy.stats import gamma
    return np.array([gamma.ppf(i / k, a=alpha, scale=alpha) for i in range(k)])


def gamma_rates_yang_mean(alpha, k):
    """Yang 1994 category means of Gamma(shape=alpha, rate=alpha): what lib/mlModel.mli:12
    documents ("means of the Gamma distribution"). sum r_i / k = 1."""
    from scipy.special import gammainc
    from scipy.stats import gamma
    cuts = np.array([0.0] + [gamma.ppf(i / k, a=alpha, scale=1.0 / alpha) for i in range(1, k)] + [np.inf])
    upper = gammainc(alpha + 1.0, cuts * alpha)
    return (upper[1:] - upper[:-1]) * k


def create(subst, a_size, pi=None, site_var=None, rates="yang_mean", diagonalize=None):
    """subst: ("JC69",) | ("K2P", b) | ("F81",) | ("HKY85", k) | ("F84", k) | ("TN93", a, b) |
    ("GTR", co) | ("Const", matrix).  site_var: None (Constant) | ("gamma", k, alpha) |
    ("theta", k, alpha, pinvar) | ("custom", rates, probs).  diagonalize(Q, sym) -> (U, D, Ui):
    the eigen-solver; default = the product's phylo_diagonalize_* (bench.py's CPU arm passes the
    reference's own so that it does not load the product library)."""
    if diagonalize is None:
        from . import engine as _engine

        diagonalize = _engine.diagonalize

    pri = priors(pi, a_size)
    kind = subst[0]
    if kind == "JC69":
        sym, q = True, m_jc69(a_size)
    elif kind == "K2P":
        sym, q = True, m_k2p(subst[1], a_size)
    elif kind == "F81":
        sym, q = False, m_f81(pri, a_size)
    elif kind == "HKY85":
        sym, q = False, m_hky85(pri, subst[1], a_size)
    elif kind == "F84":
        sym, q = False, m_f84(pri, subst[1], a_size)
    elif kind == "TN93":
        sym, q = False, m_tn93(pri, subst[1], subst[2], a_size)
    elif kind == "GTR":
        sym, q = False, m_gtr(pri, subst[1], a_size)
    elif kind == "Const":
        sym, q = False, m_file(pri, subst[1])
    else:
        raise ValueError(kind)
    gen = gamma_rates_yang_mean if rates == "yang_mean" else gamma_rates_ref_literal
    pinvar = None
    if site_var is None:
        r, p = np.array([1.0]), np.array([1.0])
    elif site_var[0] == "gamma":
        k = site_var[1]
        r, p = gen(site_var[2], k), np.full(k, 1.0 / k)
    elif site_var[0] == "theta":
        k = site_var[1]
        pinvar = site_var[3]
        if k == 1:
            r, p = np.array([1.0]), np.array([1.0])
        else:
            r, p = gen(site_var[2], k), np.full(k, 1.0 / k)
    elif site_var[0] == "custom":
        r, p = np.asarray(site_var[1], float), np.asarray(site_var[2], float)
    else:
        raise ValueError(site_var)
    U, D, Ui = diagonalize(q, sym)
    return dict(S=a_size, K=len(r), Q=q, U=U, D=D, Ui=Ui, pi=pri, rates=r, probs=p, pinvar=pinvar, sym=sym)


def synthetic_reversible(S, seed):
    """cfg4-style 20-state model: symmetric exchangeabilities U(0.1,2), pi from
    Dirichlet(5), fed through the Const/m_file route."""
    rng = np.random.default_rng(seed)
    ex = rng.uniform(0.1, 2.0, size=(S, S))
    ex = np.triu(ex, 1)
    ex = ex + ex.T
    pi = rng.dirichlet(np.full(S, 5.0))
    return ex * pi[None, :], pi


_CODONS = [a + b + c for a in "TCAG" for b in "TCAG" for c in "TCAG"]
_AA = "FFLLSSSSYY**CC*WLLLLPPPPHHQQRRRRIIIMTTTTNNKKSSRRVVVVAAAADDEEGGGG"


def gy94(kappa, omega, seed):
    """cfg5-style 61-state codon model: GY94-structured Q (single-nucleotide changes only,
    kappa for transitions, omega for non-synonymous), F3x4-style pi from `seed`."""
    rng = np.random.default_rng(seed)
    sense = [i for i, a in enumerate(_AA) if a != "*"]
    f = rng.dirichlet(np.full(4, 10.0), size=3)
    nuc = "TCAG"
    pi = np.array([f[0][nuc.index(_CODONS[c][0])] * f[1][nuc.index(_CODONS[c][1])] *
                   f[2][nuc.index(_CODONS[c][2])] for c in sense])
    pi /= pi.sum()
    S = len(sense)
    R = np.zeros((S, S))
    transitions = {("T", "C"), ("C", "T"), ("A", "G"), ("G", "A")}
    for a, ca in enumerate(sense):
        for b, cb in enumerate(sense):
            if a == b:
                continue
            diff = [(x, y) for x, y in zip(_CODONS[ca], _CODONS[cb]) if x != y]
            if len(diff) != 1:
                continue
            r = 1.0
            if diff[0] in transitions:
                r *= kappa
            if _AA[ca] != _AA[cb]:
                r *= omega
            R[a, b] = r * pi[b]
    return R, pi
