"""Run under gpurun: where the time of phylo_lk_param_gradient goes (256 taxa x 131072 patterns, GTR+G4, six
parameters): wall time per call, kernel-class times from the engine's event profiler."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from phylocaml_b200 import engine, mlmodel, tree

T, N, K = 256, int(os.environ.get("GRAD_N", 131072)), 4
co, pi = [1.0, 2.5, 0.8, 1.2, 3.0], [0.3, 0.2, 0.25, 0.25]
model = mlmodel.create(("GTR", co), 4, pi=pi, site_var=("gamma", K, 0.5))
tr = tree.random_tree(T, seed=1)
ops, ra, rb, rt, n_nodes = tree.schedule(tr)
tips = tree.evolve_tips(tr, model, N, seed=3)
up_slot, cap, up_ops, edges = tree.uppass_plan(ops, ra, rb, rt, n_nodes)
rng = np.random.default_rng(0)
dQ = rng.standard_normal((6, 4, 4)); dQ -= dQ.sum(axis=2, keepdims=True) * np.eye(4)
drates = rng.standard_normal((6, K))
eng = engine.Engine(0)
eng.lk_set_model(model); eng.lk_set_tips(tips, capacity=cap)
eng.lk_score_tree(ops, ra, rb, rt); eng.lk_uppass(ops, ra, rb, rt, up_slot)
for mode in ("1", "0", "1"):
    os.environ["PHYLO_GRAD_MMA"] = mode
    g = eng.lk_param_gradient(ops, ra, rb, rt, up_slot, dQ=dQ, drates=drates)
    torch.cuda.synchronize(); t0 = time.perf_counter()
    for _ in range(5): g = eng.lk_param_gradient(ops, ra, rb, rt, up_slot, dQ=dQ, drates=drates)
    wall = (time.perf_counter() - t0) / 5
    eng.profile(True, reset=True)
    g = eng.lk_param_gradient(ops, ra, rb, rt, up_slot, dQ=dQ, drates=drates)
    prof = eng.profile_get(); eng.profile(False)
    print("PHYLO_GRAD_MMA=%s patterns %d wall ms/call %.3f kernels %s grad[0] %.12g" % (mode, N, wall * 1e3, prof, g[0]), flush=True)
eng.close()
