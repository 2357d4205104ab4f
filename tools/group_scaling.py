#!/usr/bin/env python
"""Single-process multi-GPU check of phylo_group (run under `gpurun --gpus N`): BASELINE config 3
(DNA GTR+G4, 256 taxa x 4M patterns) scored through one group handle over 1, 2, ... visible GPUs.
Prints one JSON line per device count: wall time per evaluation (host clock around the call: the
group call is synchronous and spans several devices, so there is no single CUDA stream to put
events on), site-updates/s, and whether lnL is bit-identical to the 1-GPU value. Also the Fitch
config (64 taxa x 64M characters)."""
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402  (workload definitions)
from phylocaml_b200 import engine, tree  # noqa: E402


def main():
    import torch

    nd = torch.cuda.device_count()
    retain = int(os.environ.get("GROUP_RETAIN", "1"))
    wl = bench.WORKLOADS["dna"]
    T, N = wl["T"], wl["N"]
    model = bench.make_model(wl)
    tr = tree.random_tree(T, seed=1)
    ops, ra, rb, rt, n_nodes = tree.schedule(tr)
    tips = bench.build_tips(tree, tr, model, T, N, 4, 0)
    chars = tree.random_fitch_chars(64, 1 << 20, 4, seed=5)
    chars = np.ascontiguousarray(np.tile(chars, (1, 61)))
    ftr = tree.random_tree(64, seed=1)
    fops, fa, fb, _, fn = tree.schedule(ftr)
    ref_lnl = ref_len = None
    counts = [n for n in (1, 2, 4, 8) if n <= nd]
    for n in counts:
        g = engine.Group(list(range(n)))
        g.set_option(g.OPT_RETAIN_CLV, retain)
        g.lk_set_model(model)
        t0 = time.perf_counter()
        g.lk_set_tips(tips, capacity=n_nodes)
        t_up = time.perf_counter() - t0
        for _ in range(3):
            lnl = g.lk_score_tree(ops, ra, rb, rt)
        reps = 10
        t0 = time.perf_counter()
        for _ in range(reps):
            lnl = g.lk_score_tree(ops, ra, rb, rt)
        dt = (time.perf_counter() - t0) / reps
        ref_lnl = lnl if ref_lnl is None else ref_lnl
        g.fitch_set_tips(chars, 4, capacity=fn)
        for _ in range(3):
            ln = g.fitch_score_tree(fops, fa, fb)
        t0 = time.perf_counter()
        for _ in range(reps):
            ln = g.fitch_score_tree(fops, fa, fb)
        fdt = (time.perf_counter() - t0) / reps
        ref_len = ln if ref_len is None else ref_len
        print(json.dumps({
            "devices": n, "mode": "retain" if retain else "lnl-only", "lk_ms_per_eval": 1e3 * dt,
            "site_updates_per_s": (T - 1) * N / dt, "lnl": lnl, "lnl_bit_identical_to_1gpu": lnl == ref_lnl,
            "set_tips_ms": 1e3 * t_up, "fitch_chars": int(chars.shape[1]), "fitch_ms_per_eval": 1e3 * fdt,
            "fitch_char_ops_per_s": 63 * chars.shape[1] / fdt, "fitch_length": ln, "fitch_exact": ln == ref_len,
            "timing": "host perf_counter around synchronous group calls", "kernel_launches": g.launch_count}))
        g.close()


if __name__ == "__main__":
    main()
