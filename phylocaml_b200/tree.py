"""Harness-side tree helpers: random unrooted binary trees, the flattened post-order
schedule the engine consumes, and synthetic alignments.

The traversal mirrors Tree.post_order_edges (lib/tree.ml:171-187): starting from a root
edge (a, b) both directed subtrees are walked post-order; every interior node visited
becomes one phylo_op {parent, left, right, t_left, t_right}. Tree.random
(lib/tree.ml:508-518) is mirrored by sequential random edge insertion. In a phylocaml
deployment the OCaml Tree module does this; Python here is bench/test scaffolding.
"""
import numpy as np

from .engine import OP_DTYPE


class Tree:
    """Unrooted binary tree on leaves 0..T-1 and interior nodes T..2T-3; adj[v] lists the
    neighbours; blen[(min(u,v), max(u,v))] is the branch length."""

    def __init__(self, T, adj, blen):
        self.T, self.adj, self.blen = T, adj, blen

    def length(self, u, v):
        return self.blen[(min(u, v), max(u, v))]

    def edges(self):
        return sorted(self.blen.keys())


def random_tree(T, seed, mean_bl=0.1, bl_clamp=(1e-4, 2.0)):
    """Random topology by sequential random edge insertion (lib/tree.ml:508-518), branch
    lengths Exp(mean_bl) clamped."""
    assert T >= 2
    rng = np.random.default_rng(seed)
    adj = {0: [1], 1: [0]}
    edges = [(0, 1)]
    nxt = T
    for leaf in range(2, T):
        a, b = edges.pop(int(rng.integers(len(edges))))
        v = nxt
        nxt += 1
        adj[a].remove(b)
        adj[b].remove(a)
        adj[a].append(v)
        adj[b].append(v)
        adj[v] = [a, b, leaf]
        adj[leaf] = [v]
        edges += [(a, v), (b, v), (leaf, v)]
    rng2 = np.random.default_rng(seed + 1)
    blen = {}
    for (a, b) in sorted((min(a, b), max(a, b)) for a, b in edges):
        blen[(a, b)] = float(np.clip(rng2.exponential(mean_bl), *bl_clamp))
    return Tree(T, adj, blen)


def caterpillar_tree(T, bl=0.1):
    """Fully unbalanced tree (deepest possible): exercises rescaling and deep recursion."""
    adj = {i: [] for i in range(2 * T - 2)}
    blen = {}

    def link(a, b):
        adj[a].append(b)
        adj[b].append(a)
        blen[(min(a, b), max(a, b))] = bl

    if T == 2:
        link(0, 1)
        return Tree(T, adj, blen)
    link(0, T)
    link(1, T)
    prev = T
    for leaf in range(2, T - 1):
        v = T + leaf - 1
        link(prev, v)
        link(leaf, v)
        prev = v
    link(prev, T - 1)
    return Tree(T, {k: v for k, v in adj.items() if v}, blen)


def schedule(tree, root_edge=None, order="dfs"):
    """Flatten to (ops, root_a, root_b, root_t, n_nodes). Node ids are the tree's own
    (leaves 0..T-1). Iterative post-order of both sides of the root edge."""
    if root_edge is None:
        root_edge = (0, tree.adj[0][0])
    a, b = root_edge
    ops = []

    def walk(prev, start):
        # iterative post-order: children of `curr` are its neighbours except `prev`
        stack = [(prev, start, False)]
        while stack:
            p, c, done = stack.pop()
            nbrs = [x for x in tree.adj[c] if x != p]
            if len(tree.adj[c]) == 1:
                continue  # leaf
            if done:
                l, r = nbrs
                ops.append((c, l, r, tree.length(c, l), tree.length(c, r)))
            else:
                stack.append((p, c, True))
                for x in reversed(nbrs):
                    stack.append((c, x, False))

    walk(b, a)
    walk(a, b)
    arr = np.zeros(len(ops), dtype=OP_DTYPE)
    for i, (p, l, r, tl, tr) in enumerate(ops):
        arr[i] = (p, l, r, 0, tl, tr)
    n_nodes = max(tree.adj.keys()) + 1
    return arr, a, b, tree.length(a, b), n_nodes


def evolve_tips(tree, model, N, seed, missing_frac=0.01, dtype=np.uint8):
    """Evolve N sites down the tree under `model` from pi (one rate class drawn per site),
    return T x N one-hot state masks with `missing_frac` of the cells set to all-ones."""
    from scipy.linalg import expm

    rng = np.random.default_rng(seed)
    S, K = model["S"], model["K"]
    Q = model["Q"]
    cat = rng.choice(K, size=N, p=model["probs"] / model["probs"].sum())
    start = 0 if len(tree.adj[0]) > 1 else tree.adj[0][0]
    states = {start: rng.choice(S, size=N, p=model["pi"] / model["pi"].sum())}
    stack = [(None, start)]
    while stack:
        p, c = stack.pop()
        for x in tree.adj[c]:
            if x == p:
                continue
            t = tree.length(c, x)
            child = np.empty(N, dtype=np.int64)
            for k in range(K):
                idx = np.nonzero(cat == k)[0]
                if idx.size == 0:
                    continue
                P = expm(Q * (t * model["rates"][k]))
                P = np.clip(P, 0, None)
                cdf = np.cumsum(P / P.sum(1, keepdims=True), axis=1)
                u = rng.random(idx.size)
                child[idx] = np.minimum((u[:, None] > cdf[states[c][idx]]).sum(1), S - 1)
            states[x] = child
            stack.append((c, x))
        if c != start and len(tree.adj[c]) > 1:
            del states[c]
    T = tree.T
    one = np.ones(1, dtype=np.uint64)
    tips = np.empty((T, N), dtype=dtype)
    allones = (1 << S) - 1
    for t in range(T):
        m = (one << states[t].astype(np.uint64))
        miss = rng.random(N) < missing_frac
        m[miss] = allones
        tips[t] = m.astype(dtype)
    return tips


def random_tips(T, N, S, seed, missing_frac=0.01, dtype=np.uint8):
    """Cheap synthetic tips for throughput runs (independent uniform states)."""
    rng = np.random.default_rng(seed)
    st = rng.integers(0, S, size=(T, N), dtype=np.uint8)
    tips = (np.ones((), dtype=np.uint64) << st.astype(np.uint64))
    if missing_frac > 0:
        tips[rng.random((T, N)) < missing_frac] = (1 << S) - 1
    return tips.astype(dtype)


def random_fitch_chars(T, N, n_states, seed, ambiguity=0.02, dtype=np.uint8):
    """cfg2-style characters: singletons + `ambiguity` fraction of random 2-state sets."""
    rng = np.random.default_rng(seed)
    a = rng.integers(0, n_states, size=(T, N))
    codes = (1 << a).astype(np.uint64)
    amb = rng.random((T, N)) < ambiguity
    b = rng.integers(0, n_states, size=(T, N))
    codes[amb] |= (1 << b[amb]).astype(np.uint64)
    return codes.astype(dtype)


def uppass_plan(ops, root_a, root_b, root_t, first_free_slot):
    """Slots and schedule of the pre-order pass that phylo_lk_uppass runs (3-directional CLVs,
    lib/node.ml:363-477): every node below the root edge gets a slot for up[v] = the CLV of the rest of
    the tree above it. Returns (up_slot int32 array sized `capacity`, capacity, up_ops, edges):
    up_ops is the same pass as a plain op list (parent = up slot, left = sibling, right = the parent's up
    value), which the oracle can execute after the down-pass ops; edges = [(v, up_slot[v], t_v)] is every
    branch below the root edge as a directional pair."""
    ops = np.asarray(ops)
    n_up = 2 * len(ops)
    cap = first_free_slot + n_up
    up_slot = np.full(cap, -1, dtype=np.int32)
    nxt = first_free_slot
    upsrc = {int(root_a): (int(root_b), float(root_t)), int(root_b): (int(root_a), float(root_t))}
    up_ops = np.zeros(n_up, dtype=ops.dtype)
    edges, i = [], 0
    for op in ops[::-1]:
        p = int(op["parent"])
        kids = ((int(op["left"]), float(op["t_left"])), (int(op["right"]), float(op["t_right"])))
        src, t_src = upsrc[p]
        for c in (0, 1):
            v, tv = kids[c]
            sib, ts = kids[1 - c]
            up_slot[v] = nxt
            up_ops[i] = (nxt, sib, src, 0, ts, t_src)
            upsrc[v] = (nxt, tv)
            edges.append((v, nxt, tv))
            nxt += 1
            i += 1
    return up_slot, cap, up_ops, edges
