"""Run under gpurun with PHYLO_FITCH_TIMING=1: the per-CTA timeline of a length-only Fitch call (64 taxa x 1 M)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from phylocaml_b200 import engine, tree
T, N = 64, 1_000_000
tr = tree.random_tree(T, 1)
ops, ra, rb, rt, n_nodes = tree.schedule(tr)
chars = tree.random_fitch_chars(T, N, 4, seed=5)
eng = engine.Engine(0)
eng.fitch_set_tips(chars, 4, capacity=n_nodes)
eng.set_option(eng.OPT_RETAIN_CLV, 0)
for _ in range(24):
    v = eng.fitch_score_tree(ops, ra, rb)
print("length", v)
