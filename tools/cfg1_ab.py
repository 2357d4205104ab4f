"""Run under gpurun: config 1 (16 taxa x 10 k patterns) per-call latency under the tree-kernel variants."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from phylocaml_b200 import engine, mlmodel, tree

T, N = 16, 10000
model = mlmodel.create(("GTR", [1.0, 2.5, 0.8, 1.2, 3.0]), 4, pi=[0.3, 0.2, 0.25, 0.25], site_var=("gamma", 4, 0.5))
tr = tree.random_tree(T, seed=1)
ops, ra, rb, rt, n_nodes = tree.schedule(tr)
tips = tree.evolve_tips(tr, model, N, seed=3)
for tune in (None, "0,1,1", "0,2,1", "1,1,1", "2,1,1"):
    if tune: os.environ["PHYLO_TREEW_TUNE"] = tune
    for fused in (1, 2, 0):
        for retain in (1, 0):
            eng = engine.Engine(0)
            eng.lk_set_model(model); eng.lk_set_tips(tips, capacity=n_nodes)
            eng.set_option(eng.OPT_FUSED_TREE, fused); eng.set_option(eng.OPT_RETAIN_CLV, retain)
            for _ in range(20): v = eng.lk_score_tree(ops, ra, rb, rt)
            torch.cuda.synchronize(); t0 = time.perf_counter()
            for _ in range(200): v = eng.lk_score_tree(ops, ra, rb, rt)
            torch.cuda.synchronize(); dt = (time.perf_counter() - t0) / 200
            print("tune", tune, "fused", fused, "retain", retain, "us/call %.1f" % (dt * 1e6), "lnl", v)
            eng.close()
    if tune is None: continue
