"""ctypes binding of the C ABI in include/phylo_engine.h (libphyloc_b200.so).

This is the same surface the OCaml stubs (stubs/phylo_stubs.c) bind; Python is only the
harness language of this repository (tests, bench). There is no fallback of any kind: if the
shared library is missing, importing `load()` raises; if no CUDA device is usable,
`Engine()` raises with the library's own message.
"""
import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "lib", "libphyloc_b200.so")

OP_DTYPE = np.dtype(
    [("parent", "<i4"), ("left", "<i4"), ("right", "<i4"), ("pad_", "<i4"),
     ("t_left", "<f8"), ("t_right", "<f8")], align=True)
assert OP_DTYPE.itemsize == 32

PHYLO_OK = 0
ERR_NAMES = {-1: "PHYLO_ERR_CUDA", -2: "PHYLO_ERR_ARG", -3: "PHYLO_ERR_STATE", -4: "PHYLO_ERR_DATA",
             -5: "PHYLO_ERR_NUMERIC", -6: "PHYLO_ERR_UNSUPPORTED"}

_dp = C.POINTER(C.c_double)
_vp = C.c_void_p
_i64 = C.c_int64
_u64p = C.POINTER(C.c_uint64)

# name -> (restype, argtypes); every symbol include/phylo_engine.h declares
SIGNATURES = {
    "phylo_engine_create": (C.c_int, [C.c_int, C.POINTER(_vp)]),
    "phylo_engine_destroy": (None, [_vp]),
    "phylo_last_error": (C.c_char_p, [_vp]),
    "phylo_engine_set_stream": (C.c_int, [_vp, _vp]),
    "phylo_engine_sync": (C.c_int, [_vp]),
    "phylo_engine_launch_count": (C.c_uint64, [_vp]),
    "phylo_engine_set_symbol_table": (C.c_int, [_vp, _u64p]),
    "phylo_engine_set_option": (C.c_int, [_vp, C.c_int, _i64]),
    "phylo_engine_get_option": (C.c_int, [_vp, C.c_int, C.POINTER(_i64)]),
    "phylo_engine_profile": (C.c_int, [_vp, C.c_int]),
    "phylo_engine_profile_reset": (C.c_int, [_vp]),
    "phylo_engine_profile_get": (C.c_int, [_vp, C.c_int, _dp, _u64p]),
    "phylo_kernel_class_count": (C.c_int, []),
    "phylo_kernel_class_name": (C.c_char_p, [C.c_int]),
    "phylo_host_alloc": (C.c_int, [C.POINTER(_vp), C.c_uint64]),
    "phylo_host_free": (C.c_int, [_vp]),
    "phylo_diagonalize_sym": (C.c_int, [_dp, _dp, C.c_int]),
    "phylo_diagonalize_gtr": (C.c_int, [_dp, _dp, _dp, C.c_int]),
    "phylo_gamma_rates": (C.c_int, [C.c_double, C.c_int, C.c_int, _dp, _dp]),
    "phylo_compose_sym": (C.c_int, [_vp, _dp, _dp, C.c_double, C.c_int, _dp]),
    "phylo_compose_gtr": (C.c_int, [_vp, _dp, _dp, _dp, C.c_double, C.c_int, _dp]),
    "phylo_lk_set_model": (C.c_int, [_vp, C.c_int, C.c_int, _dp, _dp, _dp, _dp, _dp, _dp, C.c_double]),
    "phylo_lk_set_tips": (C.c_int, [_vp, C.c_int, _i64, _vp, C.c_int, _dp, C.c_int]),
    "phylo_pack_nibbles": (C.c_int, [_vp, C.c_int, _i64, _vp]),
    "phylo_fitch_pack_planes": (C.c_int, [_vp, C.c_int, C.c_int, C.c_int, _i64, _vp]),
    "phylo_fitch_plane_count": (C.c_int, [C.c_int]),
    "phylo_lk_set_tips_pitched": (C.c_int, [_vp, C.c_int, _i64, _vp, C.c_int, C.c_uint64, _dp, C.c_int]),
    "phylo_lk_node_alloc": (C.c_int, [_vp, C.POINTER(C.c_int), _u64p]),
    "phylo_lk_node_release": (C.c_int, [_vp, C.c_int, C.c_uint64]),
    "phylo_lk_node_stats": (C.c_int, [_vp, C.POINTER(C.c_int), C.POINTER(C.c_int), C.POINTER(C.c_int)]),
    "phylo_fitch_node_alloc": (C.c_int, [_vp, C.POINTER(C.c_int), _u64p]),
    "phylo_fitch_node_release": (C.c_int, [_vp, C.c_int, C.c_uint64]),
    "phylo_fitch_node_stats": (C.c_int, [_vp, C.POINTER(C.c_int), C.POINTER(C.c_int), C.POINTER(C.c_int)]),
    "phylo_lk_median_2": (C.c_int, [_vp, C.c_int, C.c_int, C.c_double, C.c_int, C.c_double]),
    "phylo_lk_median_3": (C.c_int, [_vp, C.c_int, C.c_int, C.c_double, C.c_int, C.c_double, C.c_int, C.c_double]),
    "phylo_fitch_median_3": (C.c_int, [_vp, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int]),
    "phylo_bv_eltcount": (C.c_int, [_vp, C.c_int, _i64, C.POINTER(C.c_int)]),
    "phylo_lk_score_tree": (C.c_int, [_vp, _vp, C.c_int, C.c_int, C.c_int, C.c_double, _dp]),
    "phylo_lk_uppass": (C.c_int, [_vp, _vp, C.c_int, C.c_int, C.c_int, C.c_double, _vp]),
    "phylo_lk_param_gradient": (C.c_int, [_vp, _vp, C.c_int, C.c_int, C.c_int, C.c_double, _vp, C.c_int, _dp, _dp, _dp,
                                          _dp, _dp]),
    "phylo_sankoff_set_matrix": (C.c_int, [_vp, C.c_int, _vp]),
    "phylo_sankoff_set_tips": (C.c_int, [_vp, C.c_int, _i64, C.c_int, C.c_int, _vp, _dp, C.c_int]),
    "phylo_sankoff_median_2": (C.c_int, [_vp, C.c_int, C.c_int, C.c_int, C.POINTER(C.c_uint64)]),
    "phylo_sankoff_score_tree": (C.c_int, [_vp, _vp, C.c_int, C.c_int, C.c_int, C.POINTER(C.c_uint64)]),
    "phylo_sankoff_get_costs": (C.c_int, [_vp, C.c_int, _vp]),
    "phylo_lk_shape": (C.c_int, [_vp, C.POINTER(C.c_int), C.POINTER(_i64), C.POINTER(C.c_int)]),
    "phylo_lk_edge_lnl_batch": (C.c_int, [_vp, C.c_int, _vp, _vp, _dp, _dp]),
    "phylo_exchange_alloc": (C.c_int, [_vp, C.POINTER(_vp), _vp]),
    "phylo_exchange_open": (C.c_int, [_vp, _vp, C.POINTER(_vp)]),
    "phylo_exchange_set": (C.c_int, [_vp, C.c_int, C.c_int, C.POINTER(_vp)]),
    "phylo_lk_exchange_reduce": (C.c_int, [_vp, _dp]),
    "phylo_exchange_sum_u64": (C.c_int, [_vp, C.c_uint64, C.POINTER(C.c_uint64)]),
    "phylo_plan_compile": (C.c_int, [_vp, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, _vp, C.POINTER(C.c_int)]),
    "phylo_fitch_reroot": (C.c_int, [_vp, C.c_int, C.c_int, C.c_int, C.c_int, _vp, C.POINTER(C.c_int), C.POINTER(C.c_int)]),
    "phylo_lk_score_alignment": (C.c_int, [_vp, C.c_int, _i64, _vp, C.c_int, _dp, C.c_int, _vp, C.c_int,
                                           C.c_int, C.c_int, C.c_double, _dp]),
    "phylo_lk_edge_lnl": (C.c_int, [_vp, C.c_int, C.c_int, _dp, C.c_int, _dp]),
    "phylo_tcm_set_matrix": (C.c_int, [_vp, C.c_int, _vp, C.c_int]),
    "phylo_integerize_matrix": (C.c_int, [_dp, _dp, C.c_int, C.c_int, _vp]),
    "phylo_tcm_median_2": (C.c_int, [_vp, C.c_int, C.c_int, C.c_int, C.POINTER(C.c_uint64)]),
    "phylo_tcm_score_tree": (C.c_int, [_vp, _vp, C.c_int, C.c_int, C.c_int, C.POINTER(C.c_uint64)]),
    "phylo_compress_patterns": (C.c_int, [_vp, C.c_int, _i64, _vp, C.c_int, _dp, _vp, _dp, _vp, C.POINTER(_i64)]),
    "phylo_compress_patterns_pitched": (C.c_int, [_vp, C.c_int, _i64, _vp, C.c_int, C.c_uint64, _dp, _vp, _dp, _vp, C.POINTER(_i64)]),
    "phylo_group_compress_patterns": (C.c_int, [_vp, C.c_int, _i64, _vp, C.c_int, _dp, _vp, _dp, _vp, C.POINTER(_i64)]),
    "phylo_lk_edge_prepare": (C.c_int, [_vp, C.c_int, C.c_int]),
    "phylo_lk_edge_eval": (C.c_int, [_vp, _dp, C.c_int, _dp, _dp, _dp]),
    "phylo_lk_optimize_branch": (C.c_int, [_vp, C.c_int, C.c_int, C.c_double, C.c_double, C.c_double, C.c_double,
                                           C.c_int, _dp, _dp, C.POINTER(C.c_int)]),
    "phylo_lk_get_clv": (C.c_int, [_vp, C.c_int, _dp, _vp]),
    "phylo_lk_get_site_lnl": (C.c_int, [_vp, _dp]),
    "phylo_lk_get_block_partials": (C.c_int, [_vp, _dp, C.POINTER(_i64)]),
    "phylo_reduce_partials": (C.c_double, [_dp, _i64]),
    "phylo_fitch_set_tips": (C.c_int, [_vp, C.c_int, _i64, C.c_int, C.c_int, _vp, _dp, C.c_int]),
    "phylo_fitch_set_tips_pitched": (C.c_int, [_vp, C.c_int, _i64, C.c_int, C.c_int, _vp, C.c_uint64, _dp, C.c_int]),
    "phylo_fitch_median_2": (C.c_int, [_vp, C.c_int, C.c_int, C.c_int, _u64p]),
    "phylo_fitch_distance": (C.c_int, [_vp, C.c_int, C.c_int, _u64p]),
    "phylo_fitch_score_tree": (C.c_int, [_vp, _vp, C.c_int, C.c_int, C.c_int, _u64p]),
    "phylo_fitch_get_node_costs": (C.c_int, [_vp, _u64p]),
    "phylo_fitch_uppass": (C.c_int, [_vp, _vp, C.c_int, C.c_int, C.c_int]),
    "phylo_fitch_get_states": (C.c_int, [_vp, C.c_int, C.c_int, _vp]),
    "phylo_fitch_set_states": (C.c_int, [_vp, C.c_int, _vp]),
    "phylo_bv_union": (C.c_int, [_vp, C.c_int, C.c_int, C.c_int]),
    "phylo_bv_inter": (C.c_int, [_vp, C.c_int, C.c_int, C.c_int]),
    "phylo_bv_popcount": (C.c_int, [_vp, C.c_int, _u64p]),
    "phylo_bv_saturation": (C.c_int, [_vp, C.c_int, C.c_uint64, _u64p]),
    "phylo_bv_poly_saturation": (C.c_int, [_vp, C.c_int, C.c_int, _u64p]),
    "phylo_bv_compare": (C.c_int, [_vp, C.c_int, C.c_int, C.POINTER(C.c_int)]),
    # several GPUs behind one handle
    "phylo_group_create": (C.c_int, [C.POINTER(C.c_int), C.c_int, C.POINTER(_vp)]),
    "phylo_group_destroy": (None, [_vp]),
    "phylo_group_last_error": (C.c_char_p, [_vp]),
    "phylo_group_size": (C.c_int, [_vp]),
    "phylo_group_engine": (_vp, [_vp, C.c_int]),
    "phylo_group_shard": (C.c_int, [_vp, C.c_int, C.c_int, C.POINTER(_i64), C.POINTER(_i64)]),
    "phylo_group_set_option": (C.c_int, [_vp, C.c_int, _i64]),
    "phylo_group_set_symbol_table": (C.c_int, [_vp, _u64p]),
    "phylo_group_lk_set_model": (C.c_int, [_vp, C.c_int, C.c_int, _dp, _dp, _dp, _dp, _dp, _dp, C.c_double]),
    "phylo_group_lk_set_tips": (C.c_int, [_vp, C.c_int, _i64, _vp, C.c_int, _dp, C.c_int]),
    "phylo_group_lk_score_tree": (C.c_int, [_vp, _vp, C.c_int, C.c_int, C.c_int, C.c_double, _dp]),
    "phylo_group_lk_edge_lnl": (C.c_int, [_vp, C.c_int, C.c_int, _dp, C.c_int, _dp]),
    "phylo_group_lk_optimize_branch": (C.c_int, [_vp, C.c_int, C.c_int, C.c_double, C.c_double, C.c_double,
                                                 C.c_double, C.c_int, _dp, _dp, C.POINTER(C.c_int)]),
    "phylo_group_lk_get_site_lnl": (C.c_int, [_vp, _dp]),
    "phylo_group_lk_get_clv": (C.c_int, [_vp, C.c_int, _dp, _vp]),
    "phylo_group_fitch_set_tips": (C.c_int, [_vp, C.c_int, _i64, C.c_int, C.c_int, _vp, _dp, C.c_int]),
    "phylo_group_fitch_score_tree": (C.c_int, [_vp, _vp, C.c_int, C.c_int, C.c_int, _u64p]),
    "phylo_group_fitch_uppass": (C.c_int, [_vp, _vp, C.c_int, C.c_int, C.c_int]),
    "phylo_group_fitch_get_states": (C.c_int, [_vp, C.c_int, C.c_int, _vp]),
}

_lib = None


class PhyloError(RuntimeError):
    """A C-ABI call returned non-zero (the OCaml stubs raise `Failure msg` for the same)."""

    def __init__(self, code, msg):
        super().__init__("%s: %s" % (ERR_NAMES.get(code, str(code)), msg))
        self.code = code


def load():
    """dlopen the product library and declare every entry point. Raises if it is missing."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise ImportError(
            "%s not found: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
            "(there is no CPU or PyTorch fallback)" % LIB_PATH)
    lib = C.CDLL(LIB_PATH)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib


def _p(a, typ=_vp):
    return None if a is None else a.ctypes.data_as(typ)


def _f64(a):
    return np.ascontiguousarray(a, dtype=np.float64)


def make_ops(parent, left, right, t_left=None, t_right=None):
    ops = np.zeros(len(parent), dtype=OP_DTYPE)
    ops["parent"], ops["left"], ops["right"] = parent, left, right
    ops["t_left"] = 0.0 if t_left is None else t_left
    ops["t_right"] = 0.0 if t_right is None else t_right
    return ops


def diagonalize(Q, sym):
    """MlModel.diagonalize (lib/mlModel.ml:554-583): returns (U, D, Ui or None)."""
    lib = load()
    n = Q.shape[0]
    U = np.array(Q, dtype=np.float64, order="C", copy=True)
    D = np.zeros((n, n))
    if sym:
        rc = lib.phylo_diagonalize_sym(_p(U, _dp), _p(D, _dp), n)
        Ui = None
    else:
        Ui = np.zeros((n, n))
        rc = lib.phylo_diagonalize_gtr(_p(U, _dp), _p(D, _dp), _p(Ui, _dp), n)
    if rc != PHYLO_OK:
        raise PhyloError(rc, "diagonalize failed (complex eigenvalues / not convergent / NaN)")
    return U, D, Ui


def gamma_rates(alpha, k, mode="yang_mean"):
    """(rates, probs) of the discrete Gamma: mode "ref_literal" (lib/mlModel.ml:93-99) or "yang_mean"."""
    rates, probs = np.empty(k), np.empty(k)
    rc = load().phylo_gamma_rates(float(alpha), int(k), 0 if mode == "ref_literal" else 1, _p(rates, _dp), _p(probs, _dp))
    if rc != PHYLO_OK:
        raise PhyloError(rc, "phylo_gamma_rates(alpha=%r, k=%r) failed" % (alpha, k))
    return rates, probs


def integerize_matrix(P, priors=None, sigma=4):
    """MlModel.integerized_model's conversion (lib/mlModel.ml:639-660) of P(t) to integer costs."""
    P = _f64(P)
    n = P.shape[0]
    pri = None if priors is None else _f64(priors)
    out = np.empty((n, n), dtype=np.int32)
    rc = load().phylo_integerize_matrix(_p(P, _dp), _p(pri, _dp), n, int(sigma), _p(out))
    if rc != PHYLO_OK:
        raise PhyloError(rc, "phylo_integerize_matrix failed (non-positive entry of P, or bad arguments)")
    return out


def plan_compile(ops, T, capacity, root_a, root_b):
    """(steps [n_ops + 1, 6] int32, stack depth) of the compiled schedule, or None when the schedule is
    not a plain tree (host-only; see phylo_plan_compile)."""
    ops = np.ascontiguousarray(ops, dtype=OP_DTYPE)
    steps = np.zeros((len(ops) + 1, 6), dtype=np.int32)
    depth = C.c_int()
    rc = load().phylo_plan_compile(_p(ops), len(ops), T, capacity, root_a, root_b, _p(steps), C.byref(depth))
    if rc == -6:
        return None
    if rc != PHYLO_OK:
        raise PhyloError(rc, "phylo_plan_compile: bad arguments")
    return steps, depth.value


def fitch_reroot(ops, capacity, root_a, root_b):
    """(ops', root_a', root_b'): the schedule a length-only phylo_fitch_score_tree evaluates -- the same
    unrooted tree rooted on its centre edge (host-only; see phylo_fitch_reroot)."""
    ops = np.ascontiguousarray(ops, dtype=OP_DTYPE)
    out = np.zeros(len(ops), dtype=OP_DTYPE)
    a, b = C.c_int(), C.c_int()
    rc = load().phylo_fitch_reroot(_p(ops), len(ops), capacity, root_a, root_b, _p(out), C.byref(a), C.byref(b))
    if rc != PHYLO_OK:
        raise PhyloError(rc, "phylo_fitch_reroot: bad arguments")
    return out, a.value, b.value


def pack_nibbles(tips, out=None):
    """T x N one-byte masks -> T x ceil(N/2) packed nibbles (phylo_pack_nibbles), the compact upload
    format of lk_set_tips(..., packed_n=N). `out`: optional (e.g. pinned) destination."""
    tips = np.ascontiguousarray(tips, dtype=np.uint8)
    T, N = tips.shape
    if out is None:
        out = np.empty((T, (N + 1) // 2), dtype=np.uint8)
    assert out.shape == (T, (N + 1) // 2) and out.dtype == np.uint8 and out.flags.c_contiguous
    rc = load().phylo_pack_nibbles(_p(tips), T, N, _p(out))
    if rc != PHYLO_OK:
        raise PhyloError(rc, "phylo_pack_nibbles: bad arguments")
    return out


def fitch_pack_planes(codes, n_states, out=None):
    """T x N characters (one per element) -> the bit-sliced plane layout (T x ceil(N/32)*NP uint32),
    the compact upload format of fitch_set_tips_planes."""
    codes = np.ascontiguousarray(codes)
    T, N = codes.shape
    NP = load().phylo_fitch_plane_count(int(n_states))
    shape = (T, ((N + 31) // 32) * NP)
    if out is None:
        out = np.empty(shape, dtype=np.uint32)
    assert out.shape == shape and out.dtype == np.uint32 and out.flags.c_contiguous
    rc = load().phylo_fitch_pack_planes(_p(codes), codes.dtype.itemsize, int(n_states), T, N, _p(out))
    if rc != PHYLO_OK:
        raise PhyloError(rc, "phylo_fitch_pack_planes: bad arguments")
    return out


def pinned_empty(shape, dtype):
    """numpy array over page-locked host memory from phylo_host_alloc."""
    lib = load()
    dtype = np.dtype(dtype)
    nbytes = int(np.prod(shape)) * dtype.itemsize
    ptr = _vp()
    rc = lib.phylo_host_alloc(C.byref(ptr), max(nbytes, 1))
    if rc != PHYLO_OK:
        raise PhyloError(rc, "phylo_host_alloc(%d bytes) failed" % nbytes)
    buf = (C.c_char * max(nbytes, 1)).from_address(ptr.value)
    arr = np.frombuffer(buf, dtype=dtype, count=int(np.prod(shape))).reshape(shape)
    _PINNED[arr.ctypes.data] = ptr.value
    return arr


_PINNED = {}


def pinned_free(arr):
    ptr = _PINNED.pop(arr.ctypes.data, None)
    if ptr is not None:
        load().phylo_host_free(_vp(ptr))


class Engine:
    """One engine handle = one GPU. Thin, 1:1 over the C ABI."""

    def __init__(self, device=0):
        self.lib = load()
        h = _vp()
        rc = self.lib.phylo_engine_create(device, C.byref(h))
        if rc != PHYLO_OK:
            raise PhyloError(rc, self.lib.phylo_last_error(None).decode())
        self.h = h
        self.lk_shape = None
        self.fitch_shape = None

    def close(self):
        if getattr(self, "h", None):
            self.lib.phylo_engine_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _ck(self, rc):
        if rc != PHYLO_OK:
            raise PhyloError(rc, self.lib.phylo_last_error(self.h).decode())

    def set_stream(self, cuda_stream_ptr):
        self._ck(self.lib.phylo_engine_set_stream(self.h, _vp(cuda_stream_ptr)))

    def sync(self):
        self._ck(self.lib.phylo_engine_sync(self.h))

    @property
    def launch_count(self):
        return int(self.lib.phylo_engine_launch_count(self.h))

    OPT_FUSED_TREE, OPT_RETAIN_CLV, OPT_FITCH_WALK, OPT_DEFER_SCALAR = 1, 2, 3, 4

    # ---- Sankoff cost-vector parsimony
    def sankoff_set_matrix(self, M):
        M = np.ascontiguousarray(M, dtype=np.int32)
        self._ck(self.lib.phylo_sankoff_set_matrix(self.h, M.shape[0], _p(M)))

    def sankoff_set_tips(self, codes, n_states, weights=None, capacity=None):
        codes = np.ascontiguousarray(codes)
        T, N = codes.shape
        capacity = 2 * T if capacity is None else capacity
        w = None if weights is None else _f64(weights)
        self._ck(self.lib.phylo_sankoff_set_tips(self.h, T, N, codes.dtype.itemsize, n_states, _p(codes), _p(w, _dp), capacity))
        self.sankoff_shape = (T, N, capacity, n_states)

    def sankoff_median_2(self, parent, left, right):
        out = C.c_uint64()
        self._ck(self.lib.phylo_sankoff_median_2(self.h, parent, left, right, C.byref(out)))
        return int(out.value)

    def sankoff_score_tree(self, ops, root_a, root_b):
        ops = np.ascontiguousarray(ops, dtype=OP_DTYPE)
        out = C.c_uint64()
        self._ck(self.lib.phylo_sankoff_score_tree(self.h, _p(ops), len(ops), root_a, root_b, C.byref(out)))
        return int(out.value)

    def sankoff_get_costs(self, node):
        T, N, cap, S = self.sankoff_shape
        out = np.empty((N, S), dtype=np.int32)
        self._ck(self.lib.phylo_sankoff_get_costs(self.h, node, _p(out)))
        return out

    # ---- device-side scalar exchange (include/phylo_engine.h)
    def exchange_alloc(self):
        """-> (mailbox device pointer, 64-byte cudaIpcMemHandle for other processes)"""
        ptr = _vp()
        handle = (C.c_ubyte * 64)()
        self._ck(self.lib.phylo_exchange_alloc(self.h, C.byref(ptr), handle))
        return ptr.value, bytes(handle)

    def exchange_open(self, handle):
        ptr = _vp()
        buf = (C.c_ubyte * 64).from_buffer_copy(handle)
        self._ck(self.lib.phylo_exchange_open(self.h, buf, C.byref(ptr)))
        return ptr.value

    def exchange_set(self, rank, mailboxes):
        arr = (_vp * len(mailboxes))(*mailboxes)
        self._ck(self.lib.phylo_exchange_set(self.h, len(mailboxes), rank, arr))

    def lk_exchange_reduce(self):
        out = C.c_double()
        self._ck(self.lib.phylo_lk_exchange_reduce(self.h, C.byref(out)))
        return out.value

    def exchange_sum_u64(self, value):
        out = C.c_uint64()
        self._ck(self.lib.phylo_exchange_sum_u64(self.h, C.c_uint64(int(value)), C.byref(out)))
        return int(out.value)

    def set_symbol_table(self, table):
        """256 state masks by symbol byte (phylocaml_b200.alphabet), or None for plain masks."""
        if table is None:
            self._ck(self.lib.phylo_engine_set_symbol_table(self.h, None))
            return
        t = np.ascontiguousarray(table, dtype=np.uint64)
        assert t.shape == (256,)
        self._ck(self.lib.phylo_engine_set_symbol_table(self.h, _p(t, _u64p)))

    def set_option(self, option, value):
        self._ck(self.lib.phylo_engine_set_option(self.h, option, int(value)))

    def get_option(self, option):
        v = _i64()
        self._ck(self.lib.phylo_engine_get_option(self.h, option, C.byref(v)))
        return v.value

    def profile(self, enable=True, reset=False):
        if reset:
            self._ck(self.lib.phylo_engine_profile_reset(self.h))
        self._ck(self.lib.phylo_engine_profile(self.h, 1 if enable else 0))

    def profile_get(self):
        """{kernel class name: (total ms, launches)} accumulated while profiling was on."""
        out = {}
        for c in range(self.lib.phylo_kernel_class_count()):
            ms, n = C.c_double(), C.c_uint64()
            self._ck(self.lib.phylo_engine_profile_get(self.h, c, C.byref(ms), C.byref(n)))
            if n.value:
                out[self.lib.phylo_kernel_class_name(c).decode()] = (ms.value, int(n.value))
        return out

    # ---- MlModel
    def compose(self, U, D, Ui, t):
        U, D = _f64(U), _f64(D)
        n = U.shape[0]
        P = np.empty((n, n))
        if Ui is None:
            self._ck(self.lib.phylo_compose_sym(self.h, _p(U, _dp), _p(D, _dp), float(t), n, _p(P, _dp)))
        else:
            Ui = _f64(Ui)
            self._ck(self.lib.phylo_compose_gtr(self.h, _p(U, _dp), _p(D, _dp), _p(Ui, _dp), float(t), n,
                                                _p(P, _dp)))
        return P

    # ---- Likelihood
    def lk_set_model(self, model):
        S, K = int(model["S"]), int(model["K"])
        U, D = _f64(model["U"]), _f64(model["D"])
        if D.ndim == 1:
            D = _f64(np.diag(D))
        Ui = model.get("Ui")
        Ui = None if Ui is None else _f64(Ui)
        pi, rates, probs = _f64(model["pi"]), _f64(model["rates"]), _f64(model["probs"])
        pinvar = model.get("pinvar")
        pinvar = -1.0 if pinvar is None else float(pinvar)
        self._ck(self.lib.phylo_lk_set_model(self.h, S, K, _p(U, _dp), _p(D, _dp), _p(Ui, _dp), _p(pi, _dp),
                                             _p(rates, _dp), _p(probs, _dp), pinvar))
        self.S, self.K = S, K

    def lk_set_tips(self, tips, weights=None, capacity=None, packed_n=None):
        """packed_n: `tips` is the packed form (pack_nibbles) of an alignment of packed_n patterns."""
        tips = np.ascontiguousarray(tips)
        T, N = tips.shape
        mask_bytes = tips.dtype.itemsize
        if packed_n is not None:
            assert tips.dtype == np.uint8 and N == (packed_n + 1) // 2
            N, mask_bytes = int(packed_n), 0
        capacity = 2 * T if capacity is None else capacity
        w = None if weights is None else _f64(weights)
        self._ck(self.lib.phylo_lk_set_tips(self.h, T, N, _p(tips), mask_bytes, _p(w, _dp), capacity))
        self.lk_shape = (T, N, capacity)

    def node_alloc(self, fitch=False):
        """(slot, generation) of a fresh interior node slot (phylo_lk_node_alloc / phylo_fitch_node_alloc)."""
        slot, gen = C.c_int(), C.c_uint64()
        fn = self.lib.phylo_fitch_node_alloc if fitch else self.lib.phylo_lk_node_alloc
        self._ck(fn(self.h, C.byref(slot), C.byref(gen)))
        return slot.value, gen.value

    def node_release(self, slot, generation, fitch=False):
        fn = self.lib.phylo_fitch_node_release if fitch else self.lib.phylo_lk_node_release
        self._ck(fn(self.h, int(slot), int(generation)))

    def node_stats(self, fitch=False):
        """(interior slots in the table, slots handed out, interior slots owning a device buffer)."""
        a, b, c = C.c_int(), C.c_int(), C.c_int()
        fn = self.lib.phylo_fitch_node_stats if fitch else self.lib.phylo_lk_node_stats
        self._ck(fn(self.h, C.byref(a), C.byref(b), C.byref(c)))
        return a.value, b.value, c.value

    def lk_median_2(self, parent, left, t_left, right, t_right):
        self._ck(self.lib.phylo_lk_median_2(self.h, parent, left, float(t_left), right, float(t_right)))

    def lk_median_3(self, parent, a, t_a, b, t_b, c, t_c):
        self._ck(self.lib.phylo_lk_median_3(self.h, parent, a, float(t_a), b, float(t_b), c, float(t_c)))

    def lk_uppass(self, ops, root_a, root_b, root_t, up_slot):
        """3-directional CLVs: fill up_slot[v] (one int32 per node slot, -1 = skip) with the CLV of the
        rest of the tree above v; needs the retained CLVs of a preceding lk_score_tree over `ops`."""
        ops = np.ascontiguousarray(ops, dtype=OP_DTYPE)
        up_slot = np.ascontiguousarray(up_slot, dtype=np.int32)
        self._ck(self.lib.phylo_lk_uppass(self.h, _p(ops), len(ops), root_a, root_b, float(root_t), _p(up_slot)))

    def lk_param_gradient(self, ops, root_a, root_b, root_t, up_slot, dQ=None, drates=None, dpi=None):
        """d lnL / d theta for every parameter described by (dQ[p], drates[p], dpi[p]); needs lk_score_tree +
        lk_uppass (all up slots) first. Returns the gradient vector."""
        ops = np.ascontiguousarray(ops, dtype=OP_DTYPE)
        up_slot = np.ascontiguousarray(up_slot, dtype=np.int32)
        arrs = [None if a is None else _f64(np.asarray(a)) for a in (dQ, drates, dpi)]
        n_params = next(a.shape[0] for a in arrs if a is not None)
        grad = np.empty(n_params)
        self._ck(self.lib.phylo_lk_param_gradient(
            self.h, _p(ops), len(ops), root_a, root_b, float(root_t), _p(up_slot), n_params,
            *[None if a is None else _p(a, _dp) for a in arrs], None, _p(grad, _dp)))
        return grad

    def lk_score_tree(self, ops, root_a, root_b, root_t):
        ops = np.ascontiguousarray(ops, dtype=OP_DTYPE)
        out = C.c_double()
        self._ck(self.lib.phylo_lk_score_tree(self.h, _p(ops), len(ops), root_a, root_b, float(root_t),
                                              C.byref(out)))
        return out.value

    def lk_score_alignment(self, tips, ops, root_a, root_b, root_t, weights=None, capacity=None, packed_n=None):
        """set_tips + score_tree with the upload overlapped with the scoring."""
        tips = np.ascontiguousarray(tips)
        T, N = tips.shape
        mask_bytes = tips.dtype.itemsize
        if packed_n is not None:
            assert tips.dtype == np.uint8 and N == (packed_n + 1) // 2
            N, mask_bytes = int(packed_n), 0
        capacity = 2 * T if capacity is None else capacity
        w = None if weights is None else _f64(weights)
        ops = np.ascontiguousarray(ops, dtype=OP_DTYPE)
        out = C.c_double()
        self._ck(self.lib.phylo_lk_score_alignment(self.h, T, N, _p(tips), mask_bytes, _p(w, _dp), capacity,
                                                   _p(ops), len(ops), root_a, root_b, float(root_t), C.byref(out)))
        self.lk_shape = (T, N, capacity)
        return out.value

    def tcm_set_matrix(self, M, metric=False):
        """General-TCM median table from an S x S integer cost matrix (CostMatrix, lib/costMatrix.ml)."""
        M = np.ascontiguousarray(M, dtype=np.int32)
        self._ck(self.lib.phylo_tcm_set_matrix(self.h, M.shape[0], _p(M), 1 if metric else 0))

    def tcm_median_2(self, parent, left, right):
        out = C.c_uint64()
        self._ck(self.lib.phylo_tcm_median_2(self.h, parent, left, right, C.byref(out)))
        return out.value

    def tcm_score_tree(self, ops, root_a, root_b):
        ops = np.ascontiguousarray(ops, dtype=OP_DTYPE)
        out = C.c_uint64()
        self._ck(self.lib.phylo_tcm_score_tree(self.h, _p(ops), len(ops), root_a, root_b, C.byref(out)))
        return out.value

    def compress_patterns(self, masks, weights=None):
        """Merge identical alignment columns: (patterns [T x P], weights [P], site_to_pattern [N])."""
        masks = np.ascontiguousarray(masks)
        T, N = masks.shape
        out = np.empty((T, N), dtype=masks.dtype)
        w_out = np.empty(N)
        s2p = np.empty(N, dtype=np.int32)
        n = _i64()
        w_in = None if weights is None else _f64(weights)
        self._ck(self.lib.phylo_compress_patterns(self.h, T, N, _p(masks), masks.dtype.itemsize,
                                                  None if w_in is None else _p(w_in, _dp), _p(out), _p(w_out, _dp),
                                                  _p(s2p), C.byref(n)))
        P = n.value
        return np.ascontiguousarray(out.reshape(-1)[:T * P].reshape(T, P)), w_out[:P].copy(), s2p

    def lk_edge_prepare(self, a, b):
        """Sum table of edge (a, b) for the branch-length loop (Likelihood.adjust_3)."""
        self._ck(self.lib.phylo_lk_edge_prepare(self.h, a, b))

    def lk_edge_eval(self, ts):
        """lnL(t), dlnL/dt, d2lnL/dt2 of the prepared edge for every t in ts."""
        ts = _f64(np.atleast_1d(ts))
        out = np.empty((3, ts.size))
        self._ck(self.lib.phylo_lk_edge_eval(self.h, _p(ts, _dp), ts.size, _p(out[0], _dp), _p(out[1], _dp),
                                             _p(out[2], _dp)))
        return out[0], out[1], out[2]

    def lk_optimize_branch(self, a, b, t0=0.1, t_min=1e-8, t_max=100.0, tol=1e-9, max_iter=50):
        """Maximum-likelihood length of edge (a, b): (t_opt, lnL(t_opt), iterations)."""
        t, l, it = C.c_double(), C.c_double(), C.c_int()
        self._ck(self.lib.phylo_lk_optimize_branch(self.h, a, b, t0, t_min, t_max, tol, max_iter, C.byref(t),
                                                   C.byref(l), C.byref(it)))
        return t.value, l.value, it.value

    def lk_edge_lnl_batch(self, a_slots, b_slots, ts):
        a = np.ascontiguousarray(a_slots, dtype=np.int32)
        b = np.ascontiguousarray(b_slots, dtype=np.int32)
        ts = _f64(ts)
        out = np.empty(len(a))
        self._ck(self.lib.phylo_lk_edge_lnl_batch(self.h, len(a), _p(a), _p(b), _p(ts, _dp), _p(out, _dp)))
        return out

    def lk_edge_lnl(self, a, b, ts):
        ts = _f64(np.atleast_1d(ts))
        out = np.empty(ts.size)
        self._ck(self.lib.phylo_lk_edge_lnl(self.h, a, b, _p(ts, _dp), ts.size, _p(out, _dp)))
        return out

    def lk_get_clv(self, node):
        T, N, _ = self.lk_shape
        clv = np.empty((N, self.K, self.S))
        sc = np.empty(N, dtype=np.int32)
        self._ck(self.lib.phylo_lk_get_clv(self.h, node, _p(clv, _dp), _p(sc)))
        return clv, sc

    def lk_get_site_lnl(self):
        out = np.empty(self.lk_shape[1])
        self._ck(self.lib.phylo_lk_get_site_lnl(self.h, _p(out, _dp)))
        return out

    def lk_get_block_partials(self):
        n = _i64()
        self._ck(self.lib.phylo_lk_get_block_partials(self.h, None, C.byref(n)))
        out = np.empty(n.value)
        self._ck(self.lib.phylo_lk_get_block_partials(self.h, _p(out, _dp), C.byref(n)))
        return out

    def reduce_partials(self, partials):
        partials = _f64(partials)
        return self.lib.phylo_reduce_partials(_p(partials, _dp), partials.size)

    # ---- NonAdditive / Bitvector
    def fitch_set_tips(self, codes, n_states, weights=None, capacity=None):
        codes = np.ascontiguousarray(codes)
        T, N = codes.shape
        capacity = 2 * T if capacity is None else capacity
        w = None if weights is None else _f64(weights)
        self._ck(self.lib.phylo_fitch_set_tips(self.h, T, N, codes.dtype.itemsize, n_states, _p(codes),
                                               _p(w, _dp), capacity))
        self.fitch_shape = (T, N, capacity)
        self.fitch_dtype = codes.dtype

    def fitch_set_tips_planes(self, planes, N, n_states, weights=None, capacity=None):
        """Characters already in the bit-sliced device layout (fitch_pack_planes): elt_bytes = 0."""
        planes = np.ascontiguousarray(planes, dtype=np.uint32)
        T = planes.shape[0]
        capacity = 2 * T if capacity is None else capacity
        w = None if weights is None else _f64(weights)
        self._ck(self.lib.phylo_fitch_set_tips(self.h, T, int(N), 0, n_states, _p(planes), _p(w, _dp), capacity))
        self.fitch_shape = (T, int(N), capacity)
        self.fitch_dtype = np.dtype(np.uint8 if n_states <= 8 else (np.uint16 if n_states <= 16 else
                                                                    (np.uint32 if n_states <= 32 else np.uint64)))

    def fitch_median_2(self, parent, left, right):
        out = C.c_uint64()
        self._ck(self.lib.phylo_fitch_median_2(self.h, parent, left, right, C.byref(out)))
        return out.value

    def fitch_median_3(self, dst, prelim, parent_final, left, right):
        self._ck(self.lib.phylo_fitch_median_3(self.h, dst, prelim, parent_final, left, right))

    def bv_eltcount(self, a, i):
        out = C.c_int()
        self._ck(self.lib.phylo_bv_eltcount(self.h, a, int(i), C.byref(out)))
        return out.value

    def fitch_distance(self, a, b):
        out = C.c_uint64()
        self._ck(self.lib.phylo_fitch_distance(self.h, a, b, C.byref(out)))
        return out.value

    def fitch_score_tree(self, ops, root_a, root_b):
        ops = np.ascontiguousarray(ops, dtype=OP_DTYPE)
        out = C.c_uint64()
        self._ck(self.lib.phylo_fitch_score_tree(self.h, _p(ops), len(ops), root_a, root_b, C.byref(out)))
        return out.value

    def fitch_get_node_costs(self):
        out = np.zeros(max(self.fitch_shape[2], self.node_stats(fitch=True)[0] + self.fitch_shape[0]), dtype=np.uint64)
        self._ck(self.lib.phylo_fitch_get_node_costs(self.h, _p(out, _u64p)))
        return out

    def fitch_uppass(self, ops, root_a, root_b):
        ops = np.ascontiguousarray(ops, dtype=OP_DTYPE)
        self._ck(self.lib.phylo_fitch_uppass(self.h, _p(ops), len(ops), root_a, root_b))

    def fitch_get_states(self, node, final=False):
        out = np.empty(self.fitch_shape[1], dtype=self.fitch_dtype)
        self._ck(self.lib.phylo_fitch_get_states(self.h, node, 1 if final else 0, _p(out)))
        return out

    def fitch_set_states(self, node, codes):
        codes = np.ascontiguousarray(codes, dtype=self.fitch_dtype)
        assert codes.size == self.fitch_shape[1]
        self._ck(self.lib.phylo_fitch_set_states(self.h, node, _p(codes)))

    def bv_union(self, dst, a, b):
        self._ck(self.lib.phylo_bv_union(self.h, dst, a, b))

    def bv_inter(self, dst, a, b):
        self._ck(self.lib.phylo_bv_inter(self.h, dst, a, b))

    def bv_popcount(self, a):
        out = C.c_uint64()
        self._ck(self.lib.phylo_bv_popcount(self.h, a, C.byref(out)))
        return out.value

    def bv_saturation(self, a, state_mask):
        out = C.c_uint64()
        self._ck(self.lib.phylo_bv_saturation(self.h, a, int(state_mask), C.byref(out)))
        return out.value

    def bv_poly_saturation(self, a, n):
        out = C.c_uint64()
        self._ck(self.lib.phylo_bv_poly_saturation(self.h, a, int(n), C.byref(out)))
        return out.value

    def bv_compare(self, a, b):
        out = C.c_int()
        self._ck(self.lib.phylo_bv_compare(self.h, a, b, C.byref(out)))
        return out.value


class Group:
    """Several GPUs behind one handle, in one process (phylo_group_*): one engine + one host
    worker thread per listed device, patterns / characters in contiguous 1024-aligned shards,
    lnL bit-identical to a single engine. `devices` may repeat a device id."""

    OPT_FUSED_TREE, OPT_RETAIN_CLV, OPT_FITCH_WALK = 1, 2, 3

    def __init__(self, devices):
        self.lib = load()
        devs = (C.c_int * len(devices))(*[int(d) for d in devices])
        h = _vp()
        rc = self.lib.phylo_group_create(devs, len(devices), C.byref(h))
        if rc != PHYLO_OK:
            raise PhyloError(rc, self.lib.phylo_group_last_error(None).decode())
        self.h = h
        self.size = self.lib.phylo_group_size(h)
        self.lk_shape = self.fitch_shape = None

    def close(self):
        if getattr(self, "h", None):
            self.lib.phylo_group_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _ck(self, rc):
        if rc != PHYLO_OK:
            raise PhyloError(rc, self.lib.phylo_group_last_error(self.h).decode())

    def shard(self, i, fitch=False):
        lo, hi = _i64(), _i64()
        self._ck(self.lib.phylo_group_shard(self.h, 1 if fitch else 0, i, C.byref(lo), C.byref(hi)))
        return lo.value, hi.value

    @property
    def launch_count(self):
        return sum(self.lib.phylo_engine_launch_count(self.lib.phylo_group_engine(self.h, i)) for i in range(self.size))

    def set_symbol_table(self, table):
        t = None if table is None else np.ascontiguousarray(table, dtype=np.uint64)
        self._ck(self.lib.phylo_group_set_symbol_table(self.h, None if t is None else _p(t, _u64p)))

    def set_option(self, option, value):
        self._ck(self.lib.phylo_group_set_option(self.h, option, int(value)))

    def lk_set_model(self, model):
        S, K = int(model["S"]), int(model["K"])
        U, D = _f64(model["U"]), _f64(model["D"])
        if D.ndim == 1:
            D = _f64(np.diag(D))
        Ui = model.get("Ui")
        Ui = None if Ui is None else _f64(Ui)
        pi, rates, probs = _f64(model["pi"]), _f64(model["rates"]), _f64(model["probs"])
        pinvar = model.get("pinvar")
        pinvar = -1.0 if pinvar is None else float(pinvar)
        self._ck(self.lib.phylo_group_lk_set_model(self.h, S, K, _p(U, _dp), _p(D, _dp), _p(Ui, _dp), _p(pi, _dp),
                                                   _p(rates, _dp), _p(probs, _dp), pinvar))
        self.S, self.K = S, K

    def lk_set_tips(self, tips, weights=None, capacity=None):
        tips = np.ascontiguousarray(tips)
        T, N = tips.shape
        capacity = 2 * T if capacity is None else capacity
        w = None if weights is None else _f64(weights)
        self._ck(self.lib.phylo_group_lk_set_tips(self.h, T, N, _p(tips), tips.dtype.itemsize, _p(w, _dp), capacity))
        self.lk_shape = (T, N, capacity)

    def lk_score_tree(self, ops, root_a, root_b, root_t):
        ops = np.ascontiguousarray(ops, dtype=OP_DTYPE)
        out = C.c_double()
        self._ck(self.lib.phylo_group_lk_score_tree(self.h, _p(ops), len(ops), root_a, root_b, float(root_t),
                                                    C.byref(out)))
        return out.value

    def lk_edge_lnl(self, a, b, ts):
        ts = _f64(np.atleast_1d(ts))
        out = np.empty(ts.size)
        self._ck(self.lib.phylo_group_lk_edge_lnl(self.h, a, b, _p(ts, _dp), ts.size, _p(out, _dp)))
        return out

    def lk_optimize_branch(self, a, b, t0=0.1, t_min=1e-8, t_max=100.0, tol=1e-9, max_iter=50):
        t, l, it = C.c_double(), C.c_double(), C.c_int()
        self._ck(self.lib.phylo_group_lk_optimize_branch(self.h, a, b, t0, t_min, t_max, tol, max_iter, C.byref(t),
                                                         C.byref(l), C.byref(it)))
        return t.value, l.value, it.value

    def lk_get_site_lnl(self):
        out = np.empty(self.lk_shape[1])
        self._ck(self.lib.phylo_group_lk_get_site_lnl(self.h, _p(out, _dp)))
        return out

    def lk_get_clv(self, node):
        N = self.lk_shape[1]
        clv = np.empty((N, self.K, self.S))
        sc = np.empty(N, dtype=np.int32)
        self._ck(self.lib.phylo_group_lk_get_clv(self.h, node, _p(clv, _dp), _p(sc)))
        return clv, sc

    def fitch_set_tips(self, codes, n_states, weights=None, capacity=None):
        codes = np.ascontiguousarray(codes)
        T, N = codes.shape
        capacity = 2 * T if capacity is None else capacity
        w = None if weights is None else _f64(weights)
        self._ck(self.lib.phylo_group_fitch_set_tips(self.h, T, N, codes.dtype.itemsize, n_states, _p(codes),
                                                     _p(w, _dp), capacity))
        self.fitch_shape = (T, N, capacity)
        self.fitch_dtype = codes.dtype

    def fitch_score_tree(self, ops, root_a, root_b):
        ops = np.ascontiguousarray(ops, dtype=OP_DTYPE)
        out = C.c_uint64()
        self._ck(self.lib.phylo_group_fitch_score_tree(self.h, _p(ops), len(ops), root_a, root_b, C.byref(out)))
        return out.value

    def fitch_uppass(self, ops, root_a, root_b):
        ops = np.ascontiguousarray(ops, dtype=OP_DTYPE)
        self._ck(self.lib.phylo_group_fitch_uppass(self.h, _p(ops), len(ops), root_a, root_b))

    def fitch_get_states(self, node, final=False):
        out = np.empty(self.fitch_shape[1], dtype=self.fitch_dtype)
        self._ck(self.lib.phylo_group_fitch_get_states(self.h, node, 1 if final else 0, _p(out)))
        return out

    def compress_patterns(self, masks, weights=None):
        """Engine.compress_patterns over all devices of the group (slabs of sites, merged on device 0)."""
        masks = np.ascontiguousarray(masks)
        T, N = masks.shape
        out = np.empty((T, N), dtype=masks.dtype)
        w_out = np.empty(N)
        s2p = np.empty(N, dtype=np.int32)
        n = _i64()
        w_in = None if weights is None else _f64(weights)
        self._ck(self.lib.phylo_group_compress_patterns(self.h, T, N, _p(masks), masks.dtype.itemsize,
                                                        None if w_in is None else _p(w_in, _dp), _p(out), _p(w_out, _dp),
                                                        _p(s2p), C.byref(n)))
        P = n.value
        return np.ascontiguousarray(out.reshape(-1)[:T * P].reshape(T, P)), w_out[:P].copy(), s2p
