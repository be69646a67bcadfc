"""Shared test utilities: oracle map construction, seeded workloads, flat guide lists."""
import numpy as np


def oracle_map_from(O, pmap, ref=False):
    info = pmap.info()
    om = O.Map(info["res"], info["origin"], info["dims"], info["inflate"], ref=ref)
    occ = pmap.grid("occupied")
    known = pmap.grid("known")
    om.add_cells(np.argwhere(occ != 0), occupied=True)
    free = np.argwhere((known != 0) & (occ == 0))
    if len(free):
        om.add_cells(free, occupied=False)
    return om


def random_pairs(omap, B, rng, lo=-9.5, hi=9.5, z=1.0, min_dist=2.0, max_dist=1e9):
    S, G = [], []
    while len(S) < B:
        n = 4 * (B - len(S)) + 16
        s = np.column_stack([rng.uniform(lo, hi, (n, 2)), np.full(n, z)])
        g = np.column_stack([rng.uniform(lo, hi, (n, 2)), np.full(n, z)])
        d = np.linalg.norm(g - s, axis=1)
        ok = (omap.query(s) == 0) & (omap.query(g) == 0) & (d >= min_dist) & (d <= max_dist)
        S.extend(s[ok])
        G.extend(g[ok])
    return np.array(S[:B]), np.array(G[:B])


def make_problems(tp, pmap, omap, B, seed, params=None):
    """Seeded start/goal pairs -> product front end -> (offsets, ctrl) with only valid trajectories."""
    rng = np.random.default_rng(seed)
    p = params if params is not None else tp.default_params()
    S, G = random_pairs(omap, int(B * 1.3) + 8, rng)
    off, ctrl, valid = tp.frontend_batch(pmap, p, S, G)
    keep = [b for b in range(len(S)) if valid[b] and off[b + 1] - off[b] >= 7][:B]
    new_off = [0]
    chunks = []
    for b in keep:
        chunks.append(ctrl[off[b]:off[b + 1]])
        new_off.append(new_off[-1] + len(chunks[-1]))
    return dict(offsets=np.array(new_off, np.int32), ctrl=np.concatenate(chunks, 0), starts=S[keep], goals=G[keep])


def flat_guides(per_traj):
    """list of (cp, p, v) per trajectory -> (g_offsets, g_cp, g_p, g_v)."""
    g_off = [0]
    cps, ps, vs = [], [], []
    for cp, p, v in per_traj:
        g_off.append(g_off[-1] + len(cp))
        cps.append(np.asarray(cp, np.int32))
        ps.append(np.asarray(p, float).reshape(-1, 3))
        vs.append(np.asarray(v, float).reshape(-1, 3))
    return (np.array(g_off, np.int32), np.concatenate(cps) if cps else np.zeros(0, np.int32),
            np.concatenate(ps, 0) if ps else np.zeros((0, 3)), np.concatenate(vs, 0) if vs else np.zeros((0, 3)))


def traj(problems, b):
    o = problems["offsets"]
    return problems["ctrl"][o[b]:o[b + 1]]


def rel_err(a, b):
    a, b = np.asarray(a, float), np.asarray(b, float)
    return float(np.max(np.abs(a - b) / np.maximum(1e-300, np.maximum(np.abs(a), np.abs(b))))) if a.size else 0.0
