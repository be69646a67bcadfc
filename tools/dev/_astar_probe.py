import os, sys, numpy as np
sys.path.insert(0,'/root/repo'); 
import trajectory_planner_b200 as tp, bench
pmap = tp.OccMap.from_tpm(bench.MAP_TPM); eng = tp.Engine(0); eng.set_map(pmap); p = tp.default_params()
off, ctrl = bench.make_workload(tp, pmap, eng.query_points, 512, bench.SEED, p)
segs = eng.find_collision_seg_batch(p, off, ctrl)
st, en = [], []
for b in range(512):
    c = ctrl[off[b]:off[b+1]]
    for s0, s1 in segs[b]: st.append(c[s0]); en.append(c[s1])
st, en = np.array(st), np.array(en)
paths, ex = eng.astar_batch(p, st, en)
order = np.argsort(-ex)[:6]
print('top expansions', ex[order])
for i in order[:4]:
    eng.astar_batch(p, st[i:i+1], en[i:i+1])
# a failing full-pool search
eng.astar_batch(p, np.array([[0.0,0.0,1.0]]), np.array([[4.9,0.0,1.0]]))
