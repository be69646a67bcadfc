import os, sys, numpy as np
sys.path.insert(0,'/root/repo')
os.environ["TP_PROF_DUMP"]="1"
import trajectory_planner_b200 as tp, bench
pmap = tp.OccMap.from_tpm(bench.MAP_TPM); eng = tp.Engine(0); eng.set_map(pmap); p = tp.default_params()
off, ctrl = bench.make_workload(tp, pmap, eng.query_points, 4096, bench.SEED, p)
eng.make_plan_batch(p, off, ctrl)
eng.profile_enable(True); eng.profile_get()
out, res = eng.make_plan_batch(p, off, ctrl)
prof = eng.profile_get()
print({k: round(v,2) for k,v in prof['ms'].items()})
