"""TEST INFRASTRUCTURE ONLY -- ctypes access to the CPU oracle.

`Oracle`  wraps oracle/liboracle.so (our C restatement, phylo_oracle.c).
`Ref`     wraps oracle/_ref/*.so: the reference's own lib/mlmodel.c and lib/bitvector/bv.c
          compiled unmodified (oracle/Makefile); it is used to pin the restatement.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
import this module. Nothing here is on the product path.
"""
import ctypes as C
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
REF_DIR = os.path.join(HERE, "_ref")

OP_DTYPE = np.dtype(
    [("parent", "<i4"), ("left", "<i4"), ("right", "<i4"), ("pad_", "<i4"),
     ("t_left", "<f8"), ("t_right", "<f8")], align=True)
assert OP_DTYPE.itemsize == 32

_dp = C.POINTER(C.c_double)


def build(ref=None):
    """Compile liboracle.so, and oracle/_ref when the reference sources are present."""
    if ref is None:
        ref = os.path.isdir("/root/reference/lib")
    targets = ["oracle"] + (["ref"] if ref else [])
    subprocess.run(["make", "-s", "-C", HERE] + targets, check=True)


def _ptr(a, typ=C.c_void_p):
    if a is None:
        return None
    return a.ctypes.data_as(typ)


def make_ops(parent, left, right, t_left=None, t_right=None):
    n = len(parent)
    ops = np.zeros(n, dtype=OP_DTYPE)
    ops["parent"], ops["left"], ops["right"] = parent, left, right
    ops["t_left"] = 0.0 if t_left is None else t_left
    ops["t_right"] = 0.0 if t_right is None else t_right
    return ops


def compress_patterns(masks, weights=None):
    """CPU statement of site-pattern compression (test oracle for phylo_compress_patterns):
    identical columns of the tip-major alignment `masks` [T x N] are merged, patterns are
    numbered by first occurrence, weights are summed. Returns (patterns [T x P], weights [P],
    site_to_pattern [N]). The reference carries such weights (lib/nonAdditive_c.ml:3) but has no
    code that produces them."""
    masks = np.ascontiguousarray(masks)
    cols = np.ascontiguousarray(masks.T)
    view = cols.view(np.dtype((np.void, cols.dtype.itemsize * cols.shape[1]))).ravel()
    _, first, inverse = np.unique(view, return_index=True, return_inverse=True)
    order = np.argsort(first, kind="stable")
    rank = np.empty_like(order)
    rank[order] = np.arange(len(order))
    s2p = rank[np.asarray(inverse).ravel()].astype(np.int32)
    w = np.bincount(s2p, weights=None if weights is None else np.asarray(weights, dtype=float),
                    minlength=len(order)).astype(float)
    return np.ascontiguousarray(masks[:, first[order]]), w, s2p


def tcm_median_table(M, metric=False):
    """CPU restatement of CostMatrix.find_median_general (lib/costMatrix.ml:68-86) and
    find_median_metric (:107-124) for every pair of state sets: the literal fold over
    (istate in a, jstate in b, k in candidates) with find_median_pair (:55-66) collecting every
    k that reaches the minimum. Returns (cost[a][b], median[a][b]) indexed by bit masks."""
    M = np.asarray(M)
    S = M.shape[0]
    sets = 1 << S
    cost = np.zeros((sets, sets), dtype=np.int64)
    med = np.zeros((sets, sets), dtype=np.int64)
    for a in range(1, sets):
        for b in range(1, sets):
            cand = [k for k in range(S) if (not metric) or ((a | b) >> k) & 1]
            best, assign = None, 0
            for i in range(S):
                if not (a >> i) & 1:
                    continue
                for j in range(S):
                    if not (b >> j) & 1:
                        continue
                    for k in cand:  # find_median_pair
                        c = int(M[i][k]) + int(M[j][k])
                        if best is None or c < best:
                            best, assign = c, 1 << k
                        elif c == best:
                            assign |= 1 << k
            cost[a][b], med[a][b] = best, assign
    return cost, med


def numpy_diagonalize(Q, sym):
    """(U, D, Ui) in the reference's conventions (lib/mlmodel.c:155-158,208-262) from numpy.linalg:
    gtr: Q = U D Ui; sym: the ROWS of U are the eigenvectors, Ui = None. For harness legs that must
    not touch the product library and run where oracle/_ref is not built."""
    Q = np.asarray(Q, dtype=np.float64)
    if sym:
        w, V = np.linalg.eigh(Q)
        return np.ascontiguousarray(V.T), np.diag(w), None
    w, V = np.linalg.eig(Q)
    if np.abs(np.imag(w)).max() > 0:
        raise RuntimeError("Imaginary eigenvalues")
    V = np.real(V)
    return np.ascontiguousarray(V), np.diag(np.real(w)), np.ascontiguousarray(np.linalg.inv(V))


class Oracle:
    # variant "o2" = liboracle.so (-O2 -ffp-contract=off: the parity oracle); "o3" = liboracle_o3.so
    # (-O3 -march=x86-64-v3, FMA contraction allowed: a faster CPU baseline for bench.py, never a checker)
    FILES = {"o2": "liboracle.so", "o3": "liboracle_o3.so"}

    @classmethod
    def available(cls, variant="o2"):
        return os.path.exists(os.path.join(HERE, cls.FILES[variant]))

    def __init__(self, variant="o2"):
        path = os.path.join(HERE, self.FILES[variant])
        if not os.path.exists(path):
            build(ref=False)
        L = self.lib = C.CDLL(path)
        L.oracle_compose.argtypes = [_dp, _dp, _dp, _dp, C.c_double, C.c_int]
        L.oracle_compose.restype = None
        L.oracle_reduce.argtypes = [_dp, C.c_long]
        L.oracle_reduce.restype = C.c_double
        L.oracle_reduce_blocks.argtypes = [_dp, C.c_long, _dp]
        L.oracle_reduce_blocks.restype = C.c_long
        L.oracle_lk_score_tree.argtypes = [
            C.c_int, C.c_int, _dp, _dp, _dp, _dp, _dp, _dp, C.c_double, C.c_int, C.c_long,
            C.c_void_p, C.c_int, _dp, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_double,
            _dp, C.c_void_p, _dp, C.c_int]
        L.oracle_lk_score_tree.restype = C.c_double
        L.oracle_lk_median2.argtypes = [C.c_int, C.c_int, C.c_long, _dp, _dp, _dp, C.c_void_p,
                                        _dp, C.c_void_p, _dp, C.c_void_p]
        L.oracle_lk_median2.restype = None
        L.oracle_fitch_median2.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_long, C.c_int]
        L.oracle_fitch_median2.restype = C.c_uint64
        L.oracle_fitch_distance.argtypes = [C.c_void_p, C.c_void_p, C.c_long, C.c_int]
        L.oracle_fitch_distance.restype = C.c_uint64
        L.oracle_fitch_score_tree.argtypes = [
            C.c_int, C.c_long, C.c_int, C.c_void_p, _dp, C.c_void_p, C.c_int, C.c_int, C.c_int,
            C.c_int, C.c_void_p, C.c_void_p, C.c_int]
        L.oracle_fitch_score_tree.restype = C.c_uint64
        L.oracle_fitch_uppass.argtypes = [C.c_int, C.c_long, C.c_int, C.c_void_p, C.c_void_p,
                                          C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p]
        L.oracle_fitch_uppass.restype = None

    # -- P(t)
    def compose(self, U, D, Ui, t):
        U = np.ascontiguousarray(U, dtype=np.float64)
        D = np.ascontiguousarray(D, dtype=np.float64)
        n = U.shape[0]
        if D.ndim == 1:
            D = np.diag(D)
        Ui_ = None if Ui is None else np.ascontiguousarray(Ui, dtype=np.float64)
        P = np.empty((n, n))
        self.lib.oracle_compose(_ptr(P, _dp), _ptr(U, _dp), _ptr(D, _dp), _ptr(Ui_, _dp), float(t), n)
        return P

    def reduce(self, v):
        v = np.ascontiguousarray(v, dtype=np.float64)
        return self.lib.oracle_reduce(_ptr(v, _dp), v.size)

    def reduce_blocks(self, v):
        v = np.ascontiguousarray(v, dtype=np.float64)
        out = np.empty((v.size + 1023) // 1024)
        self.lib.oracle_reduce_blocks(_ptr(v, _dp), v.size, _ptr(out, _dp))
        return out

    # -- likelihood
    def lk_score_tree(self, model, tips, weights, ops, n_nodes, root_a, root_b, root_t,
                      want_clv=False, nthreads=1):
        """model: dict with S,K,U,D,Ui,pi,rates,probs,pinvar. tips: (T,N) unsigned masks."""
        S, K = int(model["S"]), int(model["K"])
        U = np.ascontiguousarray(model["U"], dtype=np.float64)
        D = np.ascontiguousarray(model["D"], dtype=np.float64)
        if D.ndim == 1:
            D = np.ascontiguousarray(np.diag(D))
        Ui = model.get("Ui")
        Ui = None if Ui is None else np.ascontiguousarray(Ui, dtype=np.float64)
        pi = np.ascontiguousarray(model["pi"], dtype=np.float64)
        rates = np.ascontiguousarray(model["rates"], dtype=np.float64)
        probs = np.ascontiguousarray(model["probs"], dtype=np.float64)
        pinvar = model.get("pinvar")
        pinvar = -1.0 if pinvar is None else float(pinvar)
        tips = np.ascontiguousarray(tips)
        T, N = tips.shape
        w = None if weights is None else np.ascontiguousarray(weights, dtype=np.float64)
        ops = np.ascontiguousarray(ops, dtype=OP_DTYPE)
        clv = np.empty((n_nodes, N, K, S)) if want_clv else None
        scale = np.empty((n_nodes, N), dtype=np.int32) if want_clv else None
        site = np.empty(N)
        lnl = self.lib.oracle_lk_score_tree(
            S, K, _ptr(U, _dp), _ptr(D, _dp), _ptr(Ui, _dp), _ptr(pi, _dp), _ptr(rates, _dp),
            _ptr(probs, _dp), pinvar, T, N, _ptr(tips), tips.dtype.itemsize, _ptr(w, _dp),
            _ptr(ops), len(ops), n_nodes, root_a, root_b, float(root_t), _ptr(clv, _dp),
            _ptr(scale), _ptr(site, _dp), nthreads)
        return dict(lnl=lnl, clv=clv, scale=scale, site_lnl=site)

    def lk_median2(self, S, K, Pl, Pr, clv_l, sc_l, clv_r, sc_r):
        N = clv_l.shape[0]
        Pl = np.ascontiguousarray(Pl, dtype=np.float64)
        Pr = np.ascontiguousarray(Pr, dtype=np.float64)
        clv_l = np.ascontiguousarray(clv_l, dtype=np.float64)
        clv_r = np.ascontiguousarray(clv_r, dtype=np.float64)
        sc_l = np.ascontiguousarray(sc_l, dtype=np.int32)
        sc_r = np.ascontiguousarray(sc_r, dtype=np.int32)
        out = np.empty((N, K, S))
        sc = np.empty(N, dtype=np.int32)
        self.lib.oracle_lk_median2(S, K, N, _ptr(Pl, _dp), _ptr(Pr, _dp), _ptr(clv_l, _dp),
                                   _ptr(sc_l), _ptr(clv_r, _dp), _ptr(sc_r), _ptr(out, _dp), _ptr(sc))
        return out, sc

    # -- Fitch
    def fitch_median2(self, a, b):
        a = np.ascontiguousarray(a)
        b = np.ascontiguousarray(b, dtype=a.dtype)
        c = np.empty_like(a)
        cost = self.lib.oracle_fitch_median2(_ptr(c), _ptr(a), _ptr(b), a.size, a.dtype.itemsize)
        return c, int(cost)

    def fitch_distance(self, a, b):
        a = np.ascontiguousarray(a)
        b = np.ascontiguousarray(b, dtype=a.dtype)
        return int(self.lib.oracle_fitch_distance(_ptr(a), _ptr(b), a.size, a.dtype.itemsize))

    def fitch_score_tree(self, tips, weights, ops, n_nodes, root_a, root_b, want_sets=False,
                         nthreads=1):
        tips = np.ascontiguousarray(tips)
        T, N = tips.shape
        w = None if weights is None else np.ascontiguousarray(weights, dtype=np.float64)
        ops = np.ascontiguousarray(ops, dtype=OP_DTYPE)
        prelim = np.zeros((n_nodes, N), dtype=tips.dtype) if want_sets else None
        costs = np.zeros(n_nodes, dtype=np.uint64)
        total = self.lib.oracle_fitch_score_tree(
            T, N, tips.dtype.itemsize, _ptr(tips), _ptr(w, _dp), _ptr(ops), len(ops), n_nodes,
            root_a, root_b, _ptr(prelim), _ptr(costs), nthreads)
        return dict(length=int(total), prelim=prelim, node_cost=costs)

    def fitch_uppass(self, T, prelim, ops, root_a, root_b):
        prelim = np.ascontiguousarray(prelim)
        n_nodes, N = prelim.shape
        ops = np.ascontiguousarray(ops, dtype=OP_DTYPE)
        final = np.empty_like(prelim)
        self.lib.oracle_fitch_uppass(T, N, prelim.dtype.itemsize, _ptr(prelim), _ptr(ops), len(ops),
                                     n_nodes, root_a, root_b, _ptr(final))
        return final


class _Vect(C.Structure):
    # struct vect_t, lib/bitvector/bv.h:57-63
    _fields_ = [("length", C.c_ulong), ("chars", C.c_ulong), ("code", C.c_uint),
                ("msize", C.c_uint), ("data", C.c_void_p)]


class Ref:
    """The reference's own C, compiled unmodified into oracle/_ref (not available unless built
    in a container that has /root/reference; the built .so files travel to the GPU box)."""

    @staticmethod
    def available(variant="o2"):
        return os.path.exists(os.path.join(REF_DIR, "libmlmodel_ref.so")) and (
            variant == "o2" or os.path.exists(os.path.join(REF_DIR, "libbv8_ref_o3.so")))

    def __init__(self, variant="o2"):
        """variant "o3": bv.c for WIDTH=8 built with -O3 -march=x86-64-v3 instead of the reference's
        -O2 (bench.py's faster CPU baseline); everything else is the -O2 build either way."""
        if not self.available(variant):
            raise RuntimeError("oracle/_ref not built (run `make -C oracle ref` where /root/reference exists)")
        m = self.ml = C.CDLL(os.path.join(REF_DIR, "libmlmodel_ref.so"))
        m.compose_gtr.argtypes = [_dp, _dp, _dp, _dp, C.c_double, C.c_int, _dp]
        m.compose_gtr.restype = None
        m.compose_sym.argtypes = [_dp, _dp, _dp, C.c_float, C.c_int, _dp]
        m.compose_sym.restype = None
        m.diagonalize_gtr.argtypes = [_dp, _dp, _dp, C.c_int]
        m.diagonalize_gtr.restype = C.c_int
        m.diagonalize_sym.argtypes = [_dp, _dp, C.c_int]
        m.diagonalize_sym.restype = C.c_int
        m.shim_last_failure.restype = C.c_char_p
        self.bv = {}
        for w in (8, 16, 32, 64):
            b = C.CDLL(os.path.join(REF_DIR, "libbv%d_ref%s.so" % (w, "_o3" if (variant == "o3" and w == 8) else "")))
            b.bv_fitch.argtypes = [C.POINTER(_Vect)] * 3
            b.bv_fitch.restype = C.c_ulong
            b.bv_distance.argtypes = [C.POINTER(_Vect)] * 2
            b.bv_distance.restype = C.c_ulong
            for name in ("bv_union", "bv_inter"):
                getattr(b, name).argtypes = [C.POINTER(_Vect)] * 3
                getattr(b, name).restype = None
            b.bv_popcount.argtypes = [C.POINTER(_Vect)]
            b.bv_popcount.restype = C.c_ulong
            b.bv_saturation.restype = C.c_ulong
            b.bv_poly_saturation.argtypes = [C.POINTER(_Vect), C.c_int]
            b.bv_poly_saturation.restype = C.c_ulong
            b.bv_compare.argtypes = [C.POINTER(_Vect)] * 2
            b.bv_compare.restype = C.c_int
            self.bv[w] = b

    # lib/mlmodel.c:208-262 / :163-197. Returns (U, D full matrix, Ui or None) as the OCaml
    # side would hold them (row-major view of the LAPACK buffers).
    def diagonalize(self, Q, sym):
        n = Q.shape[0]
        U = np.array(Q, dtype=np.float64, order="C", copy=True)
        D = np.zeros((n, n))
        if sym:
            info = self.ml.diagonalize_sym(_ptr(U, _dp), _ptr(D, _dp), n)
            Ui = None
        else:
            Ui = np.zeros((n, n))
            info = self.ml.diagonalize_gtr(_ptr(U, _dp), _ptr(D, _dp), _ptr(Ui, _dp), n)
        fail = self.ml.shim_last_failure()
        if info != 0 or fail:
            raise RuntimeError("reference diagonalize failed: info=%d %s" % (info, fail))
        return U, D, Ui

    # lib/mlmodel.c:325-342 / :280-302
    def compose(self, U, D, Ui, t):
        n = U.shape[0]
        U = np.ascontiguousarray(U, dtype=np.float64)
        D = np.ascontiguousarray(D, dtype=np.float64)
        P = np.empty((n, n))
        tmp = np.empty((n, n))
        if Ui is None:
            self.ml.compose_sym(_ptr(P, _dp), _ptr(U, _dp), _ptr(D, _dp), float(t), n, _ptr(tmp, _dp))
        else:
            Ui = np.ascontiguousarray(Ui, dtype=np.float64)
            self.ml.compose_gtr(_ptr(P, _dp), _ptr(U, _dp), _ptr(D, _dp), _ptr(Ui, _dp), float(t), n,
                                _ptr(tmp, _dp))
        return P

    def _vect(self, a):
        return _Vect(a.size, a.size, 0, a.dtype.itemsize * 8, a.ctypes.data)

    # lib/bitvector/bv.c:148-160
    def bv_fitch(self, a, b):
        a = np.ascontiguousarray(a)
        b = np.ascontiguousarray(b, dtype=a.dtype)
        c = np.zeros_like(a)
        va, vb, vc = self._vect(a), self._vect(b), self._vect(c)
        cost = self.bv[a.dtype.itemsize * 8].bv_fitch(C.byref(vc), C.byref(va), C.byref(vb))
        return c, int(cost)

    # lib/bitvector/bv.c:46-55
    def bv_distance(self, a, b):
        a = np.ascontiguousarray(a)
        b = np.ascontiguousarray(b, dtype=a.dtype)
        va, vb = self._vect(a), self._vect(b)
        return int(self.bv[a.dtype.itemsize * 8].bv_distance(C.byref(va), C.byref(vb)))

    def fitch_score_tree(self, chars, ops, n_nodes, root_a, root_b, nthreads=1):
        """Whole-tree Fitch length out of the reference's own kernels: one bv_fitch per schedule
        entry (lib/bitvector/bv.c:148-160, the call Node.median_2 makes per node, lib/node.ml:183-198)
        plus bv_distance across the root edge (bv.c:46-55), the characters cut into `nthreads`
        contiguous slabs walked by one host thread each (ctypes releases the GIL during the calls).
        This is bench.py's `--impl reference` arm for the Fitch workload."""
        from concurrent.futures import ThreadPoolExecutor

        chars = np.ascontiguousarray(chars)
        T, N = chars.shape
        lib = self.bv[chars.dtype.itemsize * 8]
        cuts = [(N * i) // nthreads for i in range(nthreads + 1)]

        def walk(lo, hi):
            if hi <= lo:
                return 0
            sets = {t: chars[t, lo:hi] for t in range(T)}
            cost = 0
            for op in ops:
                a, b = sets[int(op["left"])], sets[int(op["right"])]
                c = np.empty(hi - lo, dtype=chars.dtype)
                va, vb, vc = self._vect(a), self._vect(b), self._vect(c)
                cost += int(lib.bv_fitch(C.byref(vc), C.byref(va), C.byref(vb)))
                sets[int(op["parent"])] = c
            va, vb = self._vect(sets[root_a]), self._vect(sets[root_b])
            return cost + int(lib.bv_distance(C.byref(va), C.byref(vb)))

        if nthreads <= 1:
            return walk(0, N)
        with ThreadPoolExecutor(nthreads) as ex:
            return sum(ex.map(lambda i: walk(cuts[i], cuts[i + 1]), range(nthreads)))

    def bv_binop(self, name, a, b):
        a = np.ascontiguousarray(a)
        b = np.ascontiguousarray(b, dtype=a.dtype)
        c = np.zeros_like(a)
        va, vb, vc = self._vect(a), self._vect(b), self._vect(c)
        getattr(self.bv[a.dtype.itemsize * 8], "bv_" + name)(C.byref(vc), C.byref(va), C.byref(vb))
        return c

    def bv_popcount(self, a):
        a = np.ascontiguousarray(a)
        va = self._vect(a)
        return int(self.bv[a.dtype.itemsize * 8].bv_popcount(C.byref(va)))

    def bv_saturation(self, a, n):
        a = np.ascontiguousarray(a)
        va = self._vect(a)
        w = a.dtype.itemsize * 8
        f = self.bv[w].bv_saturation
        f.argtypes = [C.POINTER(_Vect), {8: C.c_uint8, 16: C.c_uint16, 32: C.c_uint32, 64: C.c_uint64}[w]]
        return int(f(C.byref(va), int(n)))

    def bv_poly_saturation(self, a, n):
        a = np.ascontiguousarray(a)
        va = self._vect(a)
        return int(self.bv[a.dtype.itemsize * 8].bv_poly_saturation(C.byref(va), int(n)))

    def bv_compare(self, a, b):
        a = np.ascontiguousarray(a)
        b = np.ascontiguousarray(b, dtype=a.dtype)
        va, vb = self._vect(a), self._vect(b)
        return int(self.bv[a.dtype.itemsize * 8].bv_compare(C.byref(va), C.byref(vb)))


SANKOFF_INF = 1 << 28


def sankoff_score_tree(tips, M, weights, ops, n_nodes, root_a, root_b):
    """Cost-vector (Sankoff) parsimony, numpy restatement of the rule in include/phylo_engine.h (the
    reference only names the passes: commented-out bv_CAML_sankoff_median2_*, lib/bitvector/bv.h:97-98):
    c_p[s] = min_i (M[s][i] + c_l[i]) + min_j (M[s][j] + c_r[j]); tips 0 / infinity by mask; length =
    sum_chars w * min_{i,j} (c_a[i] + M[i][j] + c_b[j]). Pinned by exhaustive enumeration in
    tests/test_oracle_cpu.py. Returns dict(length, vec = {slot: (N, S) int64})."""
    tips = np.asarray(tips)
    M = np.asarray(M, dtype=np.int64)
    T, N = tips.shape
    S = M.shape[0]
    vec = {}

    def v(slot):
        if slot < T:
            bits = (tips[slot].astype(np.int64)[:, None] >> np.arange(S)) & 1
            return np.where(bits == 1, 0, SANKOFF_INF).astype(np.int64)
        return vec[slot]

    def relax(c):  # [n, s] = min_i M[s, i] + c[n, i]
        return (M[None, :, :] + c[:, None, :]).min(axis=2)

    for op in ops:
        vec[int(op["parent"])] = np.minimum(relax(v(int(op["left"]))) + relax(v(int(op["right"]))), SANKOFF_INF)
    join = (v(root_a) + relax(v(root_b))).min(axis=1)
    w = np.ones(N, dtype=np.int64) if weights is None else np.asarray(weights).astype(np.int64)
    return dict(length=int((join * w).sum()), vec=vec)
