"""Run under ncu (tools/directions_launches.sh): one down pass, one up pass, all edge joins and one six-parameter
gradient of a 256-taxon tree at 131072 patterns, so that the launch list shows what each call launches."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from phylocaml_b200 import engine, mlmodel, tree

T, N, K = 256, 131072, 4
model = mlmodel.create(("GTR", [1.0, 2.5, 0.8, 1.2, 3.0]), 4, pi=[0.3, 0.2, 0.25, 0.25], site_var=("gamma", K, 0.5))
tr = tree.random_tree(T, seed=1)
ops, ra, rb, rt, n_nodes = tree.schedule(tr)
tips = tree.evolve_tips(tr, model, N, seed=3)
up_slot, cap, up_ops, edges = tree.uppass_plan(ops, ra, rb, rt, n_nodes)
rng = np.random.default_rng(0)
dQ = rng.standard_normal((6, 4, 4)); dQ -= dQ.sum(axis=2, keepdims=True) * np.eye(4)
drates = rng.standard_normal((6, K))
eng = engine.Engine(0)
eng.lk_set_model(model); eng.lk_set_tips(tips, capacity=cap)
n0 = eng.launch_count
lnl = eng.lk_score_tree(ops, ra, rb, rt); n1 = eng.launch_count
eng.lk_uppass(ops, ra, rb, rt, up_slot); n2 = eng.launch_count
j = eng.lk_edge_lnl_batch([e[0] for e in edges], [e[1] for e in edges], [e[2] for e in edges]); n3 = eng.launch_count
g = eng.lk_param_gradient(ops, ra, rb, rt, up_slot, dQ=dQ, drates=drates); n4 = eng.launch_count
print("launches: down pass %d, up pass %d (%d directional CLVs), %d joins %d, gradient %d; lnL %.6f max |join - lnL| %.2e" %
      (n1 - n0, n2 - n1, len(edges), len(edges), n3 - n2, n4 - n3, lnl, float(np.max(np.abs(j - lnl)))))
eng.close()
