# which phase-A features predict a trajectory's total work?  (scheduling study for the park buckets)
import os, sys, numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import trajectory_planner_b200 as tp, bench
pmap = tp.OccMap.from_tpm(bench.MAP_TPM); eng = tp.Engine(0); eng.set_map(pmap); p = tp.default_params()
X, Y = [], []
for off_ in (0, 2, 7, 3):
    offsets, ctrl = bench.make_workload(tp, pmap, eng.query_points, 4096, bench.SEED + off_, p)
    N = np.diff(offsets)
    g = eng.init_guides_batch(p, offsets, ctrl)
    pairs = np.array([len(x["cp"]) for x in g]); nseg = np.array([len(x["segs"]) for x in g])
    seglen = np.array([int((x["segs"][:, 1] - x["segs"][:, 0]).sum()) if len(x["segs"]) else 0 for x in g])
    hit = eng.query_points(ctrl)
    key = np.add.reduceat(hit.astype(int), offsets[:-1])
    out, r = eng.make_plan_batch(p, offsets, ctrl)
    t = r["lbfgs_iters"] * (4.0 + 0.05 * N) + r["astar_expansions"] * 1.8   # us, rough
    X.append(np.column_stack([N, pairs, nseg, seglen, key])); Y.append(np.column_stack([t, r["status"], r["outer_rounds"], r["fail_count"]]))
X = np.vstack(X); Y = np.vstack(Y); t = Y[:, 0]
names = ["N", "pairs", "nseg", "seglen", "key"]
print("n", len(t), "mean work us", t.mean(), "p99", np.percentile(t, 99), "max", t.max())
for i, nm in enumerate(names):
    print(nm, "corr with work %.3f" % np.corrcoef(X[:, i], t)[0, 1], "spearman-ish %.3f" % np.corrcoef(np.argsort(np.argsort(X[:, i])), np.argsort(np.argsort(t)))[0, 1])
for nm, score in (("pairs", X[:, 1]), ("seglen", X[:, 3]), ("key", X[:, 4]), ("pairs*N", X[:, 1] * X[:, 0]), ("seglen/N", X[:, 3] / X[:, 0]), ("key/N", X[:, 4] / X[:, 0]), ("nseg", X[:, 2])):
    order = np.argsort(-score)
    top = np.argsort(-t)[:int(0.02 * len(t))]   # the 2 % longest
    rank = np.empty(len(t), int); rank[order] = np.arange(len(t))
    print("%-10s: mean start rank (0..1) of the 2%% longest trajectories %.3f, worst %.3f" % (nm, rank[top].mean() / len(t), rank[top].max() / len(t)))
top = np.argsort(-t)[:20]
print("longest:", [(int(X[i, 0]), int(X[i, 1]), int(X[i, 2]), int(X[i, 3]), int(X[i, 4]), int(t[i]), int(Y[i, 1]), int(Y[i, 2])) for i in top])
