import os, sys, numpy as np
sys.path.insert(0,'/root/repo')
import trajectory_planner_b200 as tp, bench
pmap = tp.OccMap.from_tpm(bench.MAP_TPM); eng = tp.Engine(0); eng.set_map(pmap); p = tp.default_params()
B = int(os.environ.get("PROBE_B", "256"))
off, ctrl = bench.make_workload(tp, pmap, eng.query_points, B, bench.SEED, p)
eng.make_plan_batch(p, off, ctrl)
