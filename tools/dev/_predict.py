import os, sys, numpy as np, heapq
sys.path.insert(0,'/root/repo')
import trajectory_planner_b200 as tp, bench
pmap = tp.OccMap.from_tpm(bench.MAP_TPM); eng = tp.Engine(0); eng.set_map(pmap); p = tp.default_params()
off, ctrl = bench.make_workload(tp, pmap, eng.query_points, 4096, bench.SEED, p)
N = np.diff(off)
p0 = tp.default_params(); p0.max_outer_rounds = 0
out0, r0 = eng.make_plan_batch(p0, off, ctrl)
os.environ["TP_TIMELINE"] = "/tmp/tl.bin"
out, res = eng.make_plan_batch(p, off, ctrl)
tl = np.fromfile("/tmp/tl.bin", dtype=np.int64).reshape(-1, 4)
dur = (tl[:,1]-tl[:,0])/1e6
hit = eng.has_collision_batch(p, off, out0)
print("after round 0: colliding %.3f, astar-fail %.3f" % (hit.mean(), (r0['status']==0).mean()))
eng_key = np.array([0.0]*len(N))
try:
    eng_key = eng.query_points(ctrl).astype(float)
    eng_key = np.array([eng_key[off[b]:off[b+1]].sum() for b in range(len(N))])
except Exception as ex: print('key failed', ex)
feats = dict(N=N.astype(float), iters0=r0['lbfgs_iters'].astype(float), exp0=r0['astar_expansions'].astype(float), pairs0=r0['n_guide_pairs'].astype(float),
             hit0=hit.astype(float), cost0=r0['final_cost'], key=eng_key, exp_or_key=r0['astar_expansions'].astype(float)+50*eng_key, hard=(r0['astar_expansions']>1500).astype(float)*1e6+eng_key, combo=hit*(r0['lbfgs_iters']+0.3*r0['astar_expansions']+5*r0['n_guide_pairs']) )
def makespan(order, d, P=444):
    h=[0.0]*P; heapq.heapify(h); end=0
    for i in order:
        t=heapq.heappop(h); t2=t+d[i]; end=max(end,t2); heapq.heappush(h,t2)
    return end
print("sum/P %.1f max %.1f" % (dur.sum()/444, dur.max()))
for k,f in feats.items():
    print("%-8s corr %.2f  simulated makespan %.1f ms" % (k, np.corrcoef(f,dur)[0,1], makespan(np.argsort(-f,kind='stable'), dur)))
print("oracle order makespan %.1f" % makespan(np.argsort(-dur), dur))
top = np.argsort(-dur)[:12]
print([(int(N[i]), int(r0['lbfgs_iters'][i]), int(r0['astar_expansions'][i]), int(hit[i]), round(float(r0['final_cost'][i]),0), round(dur[i],1)) for i in top])
