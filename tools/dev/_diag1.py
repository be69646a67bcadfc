import sys, numpy as np
sys.path.insert(0,'/root/repo'); sys.path.insert(0,'/root/repo/tests')
import trajectory_planner_b200 as tp
from oracle import oracle as O
from helpers import *
m = tp.OccMap.from_tpm('data/maps/square_static.tpm'); om = oracle_map_from(O, m)
pr = make_problems(tp, m, om, 96, 20261018)
e = tp.Engine(0); e.set_map(m)
off = pr['offsets']; B=len(off)-1
per=[];pls=[]
for b in range(B):
    pl=O.Planner(om); pl.set_ctrl(traj(pr,b)); pl.init_guides(); per.append(pl.get_guides()); pls.append(pl)
for strict in (0,1):
    p = tp.default_params(); p.strict_order=strict
    co, res, xf = e.optimize_batch(p, off, pr['ctrl'], flat_guides(per))
    print('--- optimize strict', strict)
    rows=[]
    for b in range(B):
        pl=O.Planner(om); pl.set_ctrl(traj(pr,b)); pl.add_guides(*per[b]); o=pl.optimize()
        d=np.abs(co[off[b]:off[b+1]]-pl.get_ctrl()).max()
        rows.append((res['iters'][b], o['iters'], res['ret'][b], o['ret'], res['fx'][b], o['fx'], d))
    rows=np.array(rows)
    eq = rows[:,0]==rows[:,1]
    print('equal iters', eq.sum(), 'of', B, ' maxdiff(eq)', rows[eq,6].max() if eq.any() else None, 'bit-identical', (rows[:,6]==0).sum())
    print('rel fx diff: median %.2e max %.2e' % (np.median(np.abs(rows[:,4]-rows[:,5])/rows[:,5]), np.max(np.abs(rows[:,4]-rows[:,5])/rows[:,5])))
    print('ctrl diff when iters differ: median %.2e max %.2e' % (np.median(rows[~eq,6]) if (~eq).any() else 0, np.max(rows[~eq,6]) if (~eq).any() else 0))
    for r in rows[:12]: print('  it %3d/%3d ret %5d/%5d fx %.6f/%.6f d=%.2e'%tuple(r))
# make plan
po = om.lib.default_params()
ok_o,out_o,st_o = O.make_plan_batch(om, po, off, pr['ctrl'], nthreads=4)
for strict in (0,1):
    p = tp.default_params(); p.strict_order=strict
    out,res = e.make_plan_batch(p, off, pr['ctrl'])
    print('--- make_plan strict', strict, 'gpu success', (res['status']==1).mean(), 'oracle', st_o['success'].mean())
    keys=["outer_rounds","fail_count","lbfgs_runs","lbfgs_iters","lbfgs_evals","astar_searches","astar_expansions","n_guide_pairs"]
    n=0
    for b in range(B):
        same = all(res[k][b]==st_o[k][b] for k in keys)
        d=np.abs(out[off[b]:off[b+1]]-out_o[off[b]:off[b+1]]).max()
        if n<30: print(b, 'same' if same else 'DIFF', 'st %d/%d'%(res['status'][b], st_o['success'][b]), 'd=%.2e'%d, [ (int(res[k][b]),int(st_o[k][b])) for k in keys], 'cost %.4f/%.4f'%(res['final_cost'][b], st_o['final_cost'][b]))
        n+=1
