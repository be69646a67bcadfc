import os, sys, numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import trajectory_planner_b200 as tp
g = np.load("tests/golden/minsnap_osqp_golden.npz")
k = 2
e = tp.Engine(0)
pt = tp.PolyTraj(e)
sols, st = pt.corridor_solve_batch([g[f"path_{k}"]], [g[f"corridor_{k}"]], 8.0, g[f"bc_{k}"].reshape(1, 12))
print(st)
