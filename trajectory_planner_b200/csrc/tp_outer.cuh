// The ViGO outer loop on the device: collision segments, A* detours, shortcut, semicircle guide
// points, re-guide decision and weight escalation.  One WARP per trajectory; lane-parallel where
// the reference's loop body is a pure map query (26-neighbour expansion, line samples, per-
// control-point occupancy), lane 0 for the inherently serial bookkeeping.
//
// Replaces: bsplineTraj::findCollisionSeg (bsplineTraj.cpp:403-445), pathSearch (:447-514),
// assignGuidePointsSemiCircle (:517-571), isReguideRequired (:573-608), the body of
// optimizeTrajectory's loop (:619-681), checkCollisionLine / shortcutPath /
// findGuidePointSemiCircle / compareCollisionSeg / isControlPointRequireNewGuide
// (bsplineTraj.h:196-304, 370-429) and AStar::AstarSearch / getPath / getDiagHeu /
// ConvertToIndexAndAdjustStartEndPoints / Coord2Index / Index2Coord (astarOcc.cpp:39-254,
// astarOcc.h:84-105).
#pragma once
#include "tp_vigo.cuh"

#define ERR_SEG_OVERFLOW 1
#define ERR_PAIR_OVERFLOW 2
#define ERR_HEAP_OVERFLOW 4
#define ERR_PATH_OVERFLOW 8
#define ERR_SC_OVERFLOW 16
#define ERR_BAND 32

#define NODE_NONE 0xFFFFFFFFu
#define ST_OPEN 1u
#define ST_CLOSED 2u

// per-warp working set
struct Worker {
  ANode* nodes;
  HeapEnt* heap_sm;  // shared, TP_HEAP_SMEM entries
  HeapEnt* heap_gl;  // HBM spill for entries >= TP_HEAP_SMEM
  double* path;      // xyz, path_cap points
  double* sc;        // max_seg x TP_SC_CAP x 3
  int* sc_len;       // max_seg
  uint32_t* round_ptr;
  int lane;
};

struct AStarFrame {  // per-search constants (uniform across the warp)
  D3 center;
  double step, inv_step;
  int k_lo;          // first stored z layer (pool index) of the height band
  uint32_t round;
};

__device__ __forceinline__ D3 as_index2coord(const VigoConst& C, const AStarFrame& F, int i, int j, int k) {
  // astarOcc.h:89-92
  return d3((double)(i - C.pool[0] / 2) * F.step + F.center.x, (double)(j - C.pool[1] / 2) * F.step + F.center.y,
            (double)(k - C.pool[2] / 2) * F.step + F.center.z);
}
__device__ __forceinline__ bool as_coord2index(const VigoConst& C, const AStarFrame& F, const D3& pt, int& i, int& j,
                                               int& k) {
  // astarOcc.h:94-105: cast<int>() truncates toward zero
  i = (int)((pt.x - F.center.x) * F.inv_step + 0.5) + C.pool[0] / 2;
  j = (int)((pt.y - F.center.y) * F.inv_step + 0.5) + C.pool[1] / 2;
  k = (int)((pt.z - F.center.z) * F.inv_step + 0.5) + C.pool[2] / 2;
  return !(i < 0 || i >= C.pool[0] || j < 0 || j >= C.pool[1] || k < 0 || k >= C.pool[2]);
}
// astarOcc.cpp:39-63 (override-style ifs) x tie breaker (astarOcc.h:65,84-87)
__device__ __forceinline__ double as_heu(int i, int j, int k, int ei, int ej, int ek) {
  double dx = (double)abs(i - ei), dy = (double)abs(j - ej), dz = (double)abs(k - ek);
  double h = 0.0;
  const int diag = (int)fmin(fmin(dx, dy), dz);
  dx -= diag;
  dy -= diag;
  dz -= diag;
  if (dx == 0) h = 1.0 * sqrt(3.0) * diag + sqrt(2.0) * fmin(dy, dz) + 1.0 * fabs(dy - dz);
  if (dy == 0) h = 1.0 * sqrt(3.0) * diag + sqrt(2.0) * fmin(dx, dz) + 1.0 * fabs(dx - dz);
  if (dz == 0) h = 1.0 * sqrt(3.0) * diag + sqrt(2.0) * fmin(dx, dy) + 1.0 * fabs(dx - dy);
  const double tie = 1.0 + 1.0 / 10000;
  return tie * h;
}

// ---- libstdc++ heap mechanics (std::priority_queue without decrease-key, astarOcc.cpp:150-228).
// comp(a,b) = a->fScore > b->fScore with keys read at comparison time; here the key lives in the
// heap entry and is patched whenever the reference mutates fScore in place, so every comparison
// sees exactly the value the reference would read through the node pointer.
struct Heap {
  HeapEnt* sm;
  HeapEnt* gl;
  ANode* nodes;
  int size;
  __device__ __forceinline__ HeapEnt get(int i) const { return i < TP_HEAP_SMEM ? sm[i] : gl[i - TP_HEAP_SMEM]; }
  __device__ __forceinline__ double key(int i) const { return i < TP_HEAP_SMEM ? sm[i].f : gl[i - TP_HEAP_SMEM].f; }
  __device__ __forceinline__ void put(int i, const HeapEnt& e) {
    if (i < TP_HEAP_SMEM) sm[i] = e; else gl[i - TP_HEAP_SMEM] = e;
    nodes[e.node].heap_pos = (uint32_t)i;
  }
  __device__ __forceinline__ void set_key(int i, double f) {
    if (i < TP_HEAP_SMEM) sm[i].f = f; else gl[i - TP_HEAP_SMEM].f = f;
  }
  // std::__push_heap
  __device__ __forceinline__ void sift_up(int hole, int top, const HeapEnt& value) {
    int parent = (hole - 1) / 2;
    while (hole > top && key(parent) > value.f) {
      put(hole, get(parent));
      hole = parent;
      parent = (hole - 1) / 2;
    }
    put(hole, value);
  }
  __device__ __forceinline__ void push(uint32_t node, double f) {
    HeapEnt e;
    e.f = f; e.node = node; e.pad = 0;
    sift_up(size, 0, e);
    ++size;
  }
  // top() + std::pop_heap (= __adjust_heap on the hole at the root, then __push_heap) + pop_back
  __device__ __forceinline__ uint32_t pop() {
    const uint32_t top = get(0).node;
    if (size > 1) {
      const HeapEnt value = get(size - 1);
      const int len = size - 1;
      int hole = 0, second = 0;
      while (second < (len - 1) / 2) {
        second = 2 * (second + 1);
        if (key(second) > key(second - 1)) second--;
        put(hole, get(second));
        hole = second;
      }
      if ((len & 1) == 0 && second == (len - 2) / 2) {
        second = 2 * (second + 1);
        put(hole, get(second - 1));
        hole = second - 1;
      }
      sift_up(hole, 0, value);
    }
    --size;
    return top;
  }
};

// AStar::AstarSearch + getPath.  Returns the number of path points written to W.path (cell centres
// start -> goal) or -1.  All 32 lanes call it with identical arguments.
//
// Node storage: only the z layers whose cell centre can lie inside [min_height, max_height] are
// stored (pool_kl of them, starting at k_lo) — cells outside the band are rejected before any node
// state is consulted (astarOcc.cpp:202) — plus one spare slot for the start cell, which the
// reference expands from even when it lies outside the band.  Per expansion the warp makes ONE
// round of global loads (current node, <= 26 neighbour nodes as LDG.128, <= 26 map words, all in
// flight together); the open-set heap lives in shared memory.
__device__ int astar_search(const DevMap& map, const VigoConst& C, Worker& W, const D3& start_in, const D3& end_in,
                            int& expansions, int& err) {
  const int lane = W.lane;
  AStarFrame F;
  F.step = map.res;
  F.inv_step = 1 / map.res;
  F.center = (start_in + end_in) / 2;
  uint32_t round = 0;
  if (lane == 0) {
    round = *W.round_ptr + 1;
    *W.round_ptr = round;
  }
  round = __shfl_sync(0xffffffffu, round, 0);
  F.round = round;
  expansions = 0;
  {
    // smallest k with (k - CZ)*step + center.z >= min_height, found with the reference's expression
    int k = (int)floor((C.p.min_height - F.center.z) * F.inv_step) + C.pool[2] / 2 - 1;
    if (k < 0) k = 0;
    while (k < C.pool[2] && (double)(k - C.pool[2] / 2) * F.step + F.center.z < C.p.min_height) ++k;
    F.k_lo = k;
  }
  // ---- ConvertToIndexAndAdjustStartEndPoints (astarOcc.cpp:90-117), lane 0
  int si = 0, sj = 0, sk = 0, ei = 0, ej = 0, ek = 0, ok = 1;
  if (lane == 0) {
    D3 sp = start_in, ep = end_in;
    if (!as_coord2index(C, F, sp, si, sj, sk) || !as_coord2index(C, F, ep, ei, ej, ek)) ok = 0;
    if (ok && dm_inflated(map, as_index2coord(C, F, si, sj, sk))) {
      do {
        const D3 dd = sp - ep;
        sp = (dd / norm3(dd)) * F.step + sp;
        if (!as_coord2index(C, F, sp, si, sj, sk)) { ok = 0; break; }
      } while (dm_inflated(map, as_index2coord(C, F, si, sj, sk)));
    }
    if (ok && dm_inflated(map, as_index2coord(C, F, ei, ej, ek))) {
      do {
        const D3 dd = ep - sp;
        ep = (dd / norm3(dd)) * F.step + ep;
        if (!as_coord2index(C, F, ep, ei, ej, ek)) { ok = 0; break; }
      } while (dm_inflated(map, as_index2coord(C, F, ei, ej, ek)));
    }
  }
  ok = __shfl_sync(0xffffffffu, ok, 0);
  if (!ok) return -1;
  si = __shfl_sync(0xffffffffu, si, 0); sj = __shfl_sync(0xffffffffu, sj, 0); sk = __shfl_sync(0xffffffffu, sk, 0);
  ei = __shfl_sync(0xffffffffu, ei, 0); ej = __shfl_sync(0xffffffffu, ej, 0); ek = __shfl_sync(0xffffffffu, ek, 0);
  const int PY = C.pool[1], KL = C.pool_kl;
  const uint32_t spare = (uint32_t)((size_t)C.pool[0] * PY * KL);
  const int k_lo = F.k_lo;
  auto lin_of = [&](int i, int j, int k) -> uint32_t {
    if (i == si && j == sj && k == sk) return spare;
    return (uint32_t)(((size_t)i * PY + j) * KL + (k - k_lo));
  };
  auto ijk_of = [&](uint32_t lin, int& i, int& j, int& k) {
    if (lin == spare) { i = si; j = sj; k = sk; return; }
    k = (int)(lin % (uint32_t)KL) + k_lo;
    const uint32_t r = lin / (uint32_t)KL;
    j = (int)(r % (uint32_t)PY);
    i = (int)(r / (uint32_t)PY);
  };
  ANode* nodes = W.nodes;
  Heap H;
  H.sm = W.heap_sm;
  H.gl = W.heap_gl;
  H.nodes = nodes;
  H.size = 0;
  if (lane == 0) {
    ANode nd;
    nd.stamp_state = (round << 2) | ST_OPEN;
    nd.parent = NODE_NONE;
    nd.g = 0;
    nd.heap_pos = 0; nd.pad0 = 0; nd.pad1 = 0;
    nodes[spare] = nd;
    H.push(spare, as_heu(si, sj, sk, ei, ej, ek));
  }
  __syncwarp();
  int result = -1;
  int num_iter = 0;
  uint32_t goal_lin = NODE_NONE;
  for (;;) {
    // ---- pop (lane 0, shared memory)
    uint32_t cur = NODE_NONE;
    if (lane == 0 && H.size > 0) cur = H.pop();
    cur = __shfl_sync(0xffffffffu, cur, 0);
    if (cur == NODE_NONE) break;  // open set empty
    ++num_iter;
    int ci, cj, ck;
    ijk_of(cur, ci, cj, ck);
    if (ci == ei && cj == ej && ck == ek) {
      goal_lin = cur;
      result = 0;
      break;
    }
    const double gcur = nodes[cur].g;  // uniform address: one broadcast load, overlaps the loads below
    if (lane == 0) nodes[cur].stamp_state = (round << 2) | ST_CLOSED;
    // ---- lane-parallel neighbour evaluation (astarOcc.cpp:173-229); lane L <-> (dx,dy,dz) in the
    // reference's loop order
    int kind = 0;  // 0 skip, 1 push (new node), 2 in-place update
    uint32_t nl = NODE_NONE;
    double tentative = 0, fnew = 0;
    if (lane < 27 && lane != 13) {
      const int dx = lane / 9 - 1, dy = (lane / 3) % 3 - 1, dz = lane % 3 - 1;
      const int ni = ci + dx, nj = cj + dy, nk = ck + dz;
      const bool inb = !(ni < 1 || ni >= C.pool[0] - 1 || nj < 1 || nj >= PY - 1 || nk < 1 || nk >= C.pool[2] - 1);
      if (inb) {
        const D3 pc = as_index2coord(C, F, ni, nj, nk);
        const bool is_start = (ni == si && nj == sj && nk == sk);
        const bool band = !(pc.z > C.p.max_height || pc.z < C.p.min_height);
        const bool layer_ok = nk >= k_lo && nk < k_lo + KL;
        if (band && !layer_ok && !is_start) err |= ERR_BAND;  // cannot happen (pool_kl has slack)
        // cells outside the band are rejected by :202 whatever their node state says, except that a
        // CLOSED start node is skipped one line earlier — same outcome (skip) either way.
        if ((band && layer_ok) || is_start) {
          nl = lin_of(ni, nj, nk);
          const uint4 raw = *reinterpret_cast<const uint4*>(&nodes[nl]);  // stamp_state, parent, g
          const bool blocked_map = band ? dm_inflated(map, pc) : true;   // issued before `raw` is consumed
          const bool explored = (raw.x >> 2) == round;
          const uint32_t state = raw.x & 3u;
          if (!(explored && state == ST_CLOSED)) {
            if (blocked_map) {
              // blocked this round: remember it so the map is not queried again (observably the same
              // as the reference's stale-state handling: the cell is skipped on every visit)
              nodes[nl].stamp_state = (round << 2) | ST_CLOSED;
            } else {
              tentative = gcur + sqrt((double)(dx * dx + dy * dy + dz * dz));
              const double gold = __hiloint2double((int)raw.w, (int)raw.z);
              if (!explored) {
                kind = 1;
                fnew = tentative + as_heu(ni, nj, nk, ei, ej, ek);
                uint4 w;
                w.x = (round << 2) | ST_OPEN;
                w.y = cur;
                w.z = (uint32_t)__double2loint(tentative);
                w.w = (uint32_t)__double2hiint(tentative);
                *reinterpret_cast<uint4*>(&nodes[nl]) = w;
              } else if (tentative < gold) {
                kind = 2;
                fnew = tentative + as_heu(ni, nj, nk, ei, ej, ek);
              }
            }
          }
        }
      }
    }
    __syncwarp();
    // ---- serial pass in the reference's neighbour order: in-place key updates and pushes
    unsigned mask = __ballot_sync(0xffffffffu, kind != 0);
    int overflow = 0;
    while (mask) {
      const int L = __ffs(mask) - 1;
      mask &= mask - 1;
      const int kL = __shfl_sync(0xffffffffu, kind, L);
      const uint32_t nL = __shfl_sync(0xffffffffu, nl, L);
      const double fL = __shfl_sync(0xffffffffu, fnew, L);
      if (kL == 2) {
        const double tL = __shfl_sync(0xffffffffu, tentative, L);
        if (lane == 0) {
          nodes[nL].parent = cur;
          nodes[nL].g = tL;
          H.set_key((int)nodes[nL].heap_pos, fL);
        }
      } else if (lane == 0) {
        if (H.size >= C.heap_cap) overflow = 1;
        else H.push(nL, fL);
      }
      __syncwarp();
    }
    if (__any_sync(0xffffffffu, overflow != 0 || (err & ERR_BAND) != 0)) {
      err |= ERR_HEAP_OVERFLOW;
      break;
    }
    if (C.p.astar_max_expansions > 0 && num_iter >= C.p.astar_max_expansions) break;
  }
  expansions = num_iter;
  if (result < 0) return -1;
  // ---- retrievePath + getPath (astarOcc.cpp:77-88, 246-254), lane 0
  int len = 0;
  if (lane == 0) {
    uint32_t p = goal_lin;
    while (p != NODE_NONE) {
      ++len;
      p = nodes[p].parent;
    }
    if (len + 1 > C.path_cap) {
      err |= ERR_PATH_OVERFLOW;
      len = -1;
    } else {
      p = goal_lin;
      int w = len - 1;
      while (p != NODE_NONE) {
        int i, j, k;
        ijk_of(p, i, j, k);
        const D3 c = as_index2coord(C, F, i, j, k);
        W.path[3 * w] = c.x;
        W.path[3 * w + 1] = c.y;
        W.path[3 * w + 2] = c.z;
        --w;
        p = nodes[p].parent;
      }
    }
  }
  len = __shfl_sync(0xffffffffu, len, 0);
  __syncwarp();
  return len;
}

// bsplineTraj::checkCollisionLine (bsplineTraj.h:196-204): samples a*p1 + (1-a)*p2 for the serially
// accumulated a = 0, res, 2res, ... <= 1 (table a_line); lanes take samples.
__device__ __forceinline__ bool check_collision_line(const DevMap& map, const VigoConst& C, const BatchView& bv,
                                                     const D3& p1, const D3& p2, int lane) {
  bool any = false;
  for (int base = 0; base < C.n_a_line; base += 32) {
    const int s = base + lane;
    bool hit = false;
    if (s < C.n_a_line) {
      const double a = bv.a_line[s];
      const D3 pm = a * p1 + (1 - a) * p2;
      hit = dm_inflated(map, pm);
    }
    if (__any_sync(0xffffffffu, hit)) {
      any = true;
      break;
    }
  }
  return any;
}

__device__ __forceinline__ D3 ld3(const double* p, int i) { return d3(p[3 * i], p[3 * i + 1], p[3 * i + 2]); }
__device__ __forceinline__ void st3(double* p, int i, const D3& v) {
  p[3 * i] = v.x;
  p[3 * i + 1] = v.y;
  p[3 * i + 2] = v.z;
}

// bsplineTraj::shortcutPath (bsplineTraj.h:206-240): path (len points) -> sc; returns its length
__device__ int shortcut_path(const DevMap& map, const VigoConst& C, const BatchView& bv, const double* path, int len,
                             double* sc, int lane, int& err) {
  int n = 0;
  auto push = [&](const D3& p) {
    if (n < TP_SC_CAP) {
      if (lane == 0) st3(sc, n, p);
    } else
      err |= ERR_SC_OVERFLOW;
    ++n;
  };
  int ptr1 = 0, ptr2 = 2;
  push(ld3(path, 0));
  if (len == 1) return n;
  if (len == 2) {
    push(ld3(path, 1));
    return n;
  }
  for (;;) {
    if (ptr2 > len - 1) break;
    const D3 p1 = ld3(path, ptr1), p2 = ld3(path, ptr2);
    if (!check_collision_line(map, C, bv, p1, p2, lane)) {
      if (ptr2 >= len - 1) {
        push(p2);
        break;
      }
      ++ptr2;
    } else {
      push(ld3(path, ptr2 - 1));
      ptr1 = ptr2 - 1;
      ptr2 = ptr1 + 2;
    }
  }
  __syncwarp();
  return n > TP_SC_CAP ? TP_SC_CAP : n;
}

// utils.h:84-86
__device__ __forceinline__ double angle_between(const D3& a, const D3& b) { return tp_atan2(norm3(cross3(a, b)), dot3(a, b)); }

// bsplineTraj::findGuidePointSemiCircle (bsplineTraj.h:251-304), serial (lane 0)
__device__ bool find_guide_point(int cpIdx, int segFirst, int segSecond, const double* path, int plen, D3& guide) {
  const double PI_const = 3.1415926;  // utils.h:19 (sic)
  const double minAngle = PI_const * 0.0 / 4.0, maxAngle = PI_const * 4.0 / 4.0;
  const int numCp = segSecond - segFirst - 1;
  double targetAngle;
  D3 psudo;
  const D3 p0 = ld3(path, 0), pb = ld3(path, plen - 1);
  if (numCp != 0) {
    const int order = cpIdx - segFirst;
    targetAngle = (cpIdx - segFirst) * PI_const / (numCp + 2);
    targetAngle = fmin(fmax(minAngle, targetAngle), maxAngle);
    const double ratio = (double)order / (double)(numCp + 1.0);
    psudo = ratio * (pb - p0) + p0;
  } else {
    targetAngle = PI_const / 2.0;
    psudo = (p0 + pb) / 2.0;
  }
  const D3 direction = p0 - psudo;
  for (int i = 0; i + 1 < plen; ++i) {
    const D3 wc = ld3(path, i), wn = ld3(path, i + 1);
    const double angleCurr = angle_between(direction, wc - psudo);
    const double angleNext = angle_between(direction, wn - psudo);
    if (targetAngle >= angleCurr && targetAngle <= angleNext) {
      double prevDiff = 0.0;
      D3 prevTemp = d3(0, 0, 0);
      for (double a = 1.0; a >= 0.0; a -= 0.1) {
        const D3 temp = a * wc + (1 - a) * wn;
        const double tempAngle = angle_between(direction, temp - psudo);
        const double diff = tempAngle - targetAngle;
        if (diff == 0) {
          guide = temp;
          return true;
        }
        if (diff * prevDiff < 0) {
          const double total = fabs(diff) + fabs(prevDiff);
          guide = fabs(prevDiff) / total * (temp - prevTemp) + prevTemp;
          return true;
        }
        prevDiff = diff;
        prevTemp = temp;
      }
    }
  }
  return false;
}

// append one (guide point, direction) pair to control point c of trajectory `st` (lane 0)
__device__ void append_pair(const DevMap& map, const VigoConst& C, const BatchView& bv, int b, TrajState& st, int c,
                            const D3& gp, const D3& gv) {
  if (st.n_pairs >= C.gcap) {
    st.err |= ERR_PAIR_OVERFLOW;
    return;
  }
  const int gi = st.n_pairs++;
  GuidePair& pr = bv.pairs[(size_t)b * C.gcap + gi];
  pr.p[0] = gp.x; pr.p[1] = gp.y; pr.p[2] = gp.z;
  pr.v[0] = gv.x; pr.v[1] = gv.y; pr.v[2] = gv.z;
  pr.next = -1;
  pr.unknown = dm_unknown(map, gp) ? 1 : 0;
  const int ci = st.off + c;
  if (bv.cp_tail[ci] < 0) bv.cp_head[ci] = gi;
  else bv.pairs[(size_t)b * C.gcap + bv.cp_tail[ci]].next = gi;
  bv.cp_tail[ci] = gi;
}

// bsplineTraj::assignGuidePointsSemiCircle (bsplineTraj.cpp:517-571) over the shortcut paths kept
// in W.sc; serial, lane 0.  npaths may exceed nseg after the merge quirk of pathSearch.
__device__ void assign_guides(const DevMap& map, const VigoConst& C, const BatchView& bv, int b, TrajState& st,
                              const Worker& W, const int (*segs)[2], int nseg, int npaths) {
  const double* ctrl = bv.ctrl + 3 * (size_t)st.off;
  D3 guide = d3(0, 0, 0);  // the reference leaves it uninitialised when the search fails (H4)
  for (int i = 0; i < nseg && i < npaths; ++i) {
    const int s0 = segs[i][0], s1 = segs[i][1];
    const double* path = W.sc + (size_t)i * TP_SC_CAP * 3;
    const int plen = W.sc_len[i];
    for (int c = s0 + 1; c < s1; ++c) {
      find_guide_point(c, s0, s1, path, plen, guide);
      const D3 dd = guide - ld3(ctrl, c);
      append_pair(map, C, bv, b, st, c, guide, dd / norm3(dd));
    }
    if (s1 - s0 - 1 == 0) {
      find_guide_point(s0, s0, s1, path, plen, guide);
      const D3 mid = (ld3(ctrl, s0) + ld3(ctrl, s1)) / 2.0;
      const D3 dd = guide - mid;
      const D3 dir = dd / norm3(dd);
      for (int c = s0 - 1; c <= s1 + 1; ++c)
        if (c >= TP_DEGREE && c <= st.N - TP_DEGREE - 1) append_pair(map, C, bv, b, st, c, guide, dir);
    }
  }
}

// bsplineTraj::findCollisionSeg (bsplineTraj.cpp:403-445).  Lanes evaluate the per-point and
// per-line map queries (pure), lane 0 replays the serial scan.  hit/line: shared scratch bytes.
__device__ int find_collision_seg(const DevMap& map, const VigoConst& C, const BatchView& bv, const TrajState& st,
                                  uint8_t* hit, uint8_t* line, int (*out)[2], int lane, int& err) {
  const double* ctrl = bv.ctrl + 3 * (size_t)st.off;
  const int N = st.N;
  const int endIdx = (int)((N - TP_DEGREE - 1) - C.p.not_check_ratio * (N - 2 * TP_DEGREE));
  for (int i = TP_DEGREE + lane; i <= endIdx; i += 32) hit[i] = dm_inflated(map, ld3(ctrl, i)) ? 1 : 0;
  __syncwarp();
  for (int i = TP_DEGREE + 1 + lane; i <= endIdx; i += 32) {
    uint8_t l = 0;
    if (!hit[i - 1] && !hit[i]) l = dm_line(map, ld3(ctrl, i - 1), ld3(ctrl, i)) ? 1 : 0;
    line[i] = l;
  }
  __syncwarp();
  int n = 0;
  if (lane == 0) {
    bool prev = false;
    int s = TP_DEGREE, e = TP_DEGREE;
    auto push = [&](int a, int b2) {
      if (n < C.max_seg) {
        out[n][0] = a;
        out[n][1] = b2;
      } else
        err |= ERR_SEG_OVERFLOW;
      ++n;
    };
    for (int i = TP_DEGREE; i <= endIdx; ++i) {
      const bool h = hit[i] != 0;
      if (h != prev) {
        if (h) s = i - 1;
        else {
          e = i;
          push(s, e);
        }
      }
      if (h && i == endIdx - 1) {  // corner case fires one index early (reference quirk)
        e = N - 1;
        push(s, e);
      }
      if (i != TP_DEGREE && !prev && !h && line[i]) push(i - 1, i);
      prev = h;
    }
    if (n > C.max_seg) n = C.max_seg;
  }
  n = __shfl_sync(0xffffffffu, n, 0);
  __syncwarp();
  return n;
}

// bsplineTraj::pathSearch (bsplineTraj.cpp:447-514) + the shortcut of each found path (kept in
// W.sc[path index]).  segs/nseg are updated in place by the merge quirk (:496-511: after a merge
// only the merged segments survive).  Returns the number of paths, or -1 when the search fails.
__device__ int path_search(const DevMap& map, const VigoConst& C, const BatchView& bv, TrajState& st, Worker& W,
                           int (*segs)[2], int& nseg, int& err) {
  const double* ctrl = bv.ctrl + 3 * (size_t)st.off;
  const int lane = W.lane;
  int npaths = 0;
  int merged[TP_MAX_SEG_HARD / 2];
  int nmerged = 0;
  const int num = nseg;
  auto run = [&](int a, int b2) -> bool {
    const D3 ps = ld3(ctrl, a), pe = ld3(ctrl, b2);
    int ex = 0;
    int len = astar_search(map, C, W, ps, pe, ex, err);
    if (lane == 0) {
      st.astar_searches += 1;
      st.astar_expansions += ex;
      st.vclock += 30LL * ex;
    }
    if (len < 0) return false;
    // searchedPath[0] = pStart; push_back(pEnd)  (:457-458)
    if (lane == 0) {
      st3(W.path, 0, ps);
      st3(W.path, len, pe);
    }
    __syncwarp();
    if (npaths < C.max_seg) {
      const int sl = shortcut_path(map, C, bv, W.path, len + 1, W.sc + (size_t)npaths * TP_SC_CAP * 3, lane, err);
      if (lane == 0) W.sc_len[npaths] = sl;
    } else
      err |= ERR_SEG_OVERFLOW;
    ++npaths;
    __syncwarp();
    return true;
  };
  for (int i = 0; i < num; ++i) {
    if (run(segs[i][0], segs[i][1])) continue;
    bool recovered = false;
    if (i + 1 < num) {
      if (segs[i + 1][0] - segs[i][1] <= 2) {
        if (run(segs[i][0], segs[i + 1][1])) {
          if (nmerged < TP_MAX_SEG_HARD / 2) merged[nmerged++] = i;
          ++i;
          recovered = true;
        }
      }
    }
    if (!recovered) return -1;
  }
  if (nmerged != 0) {
    // only the merged segments survive (reference pushes the others into the wrong vector, :507)
    // (merged[q] >= q, so the in-place forward rewrite never reads an overwritten entry)
    __syncwarp();
    if (lane == 0)
      for (int q = 0; q < nmerged; ++q) {
        const int a = segs[merged[q]][0], b2 = segs[merged[q] + 1][1];
        segs[q][0] = a;
        segs[q][1] = b2;
      }
    __syncwarp();
    nseg = nmerged;
  }
  return npaths;
}

__device__ __forceinline__ bool index_in_seg(const int (*segs)[2], int n, int idx) {
  for (int i = 0; i < n; ++i)
    if (idx >= segs[i][0] && idx <= segs[i][1]) return true;
  return false;
}
__device__ __forceinline__ int find_seg_index(const int (*segs)[2], int n, int idx) {
  for (int i = 0; i < n; ++i)
    if (idx >= segs[i][0] && idx <= segs[i][1]) return i;
  return -1;
}
// bsplineTraj::isControlPointRequireNewGuide (bsplineTraj.h:417-429)
__device__ bool cp_requires_new_guide(const VigoConst& C, const BatchView& bv, int b, const TrajState& st, int c) {
  const D3 cp = ld3(bv.ctrl + 3 * (size_t)st.off, c);
  for (int gi = bv.cp_head[st.off + c]; gi >= 0;) {
    const GuidePair& pr = bv.pairs[(size_t)b * C.gcap + gi];
    const double dist = dot3(cp - d3(pr.p[0], pr.p[1], pr.p[2]), d3(pr.v[0], pr.v[1], pr.v[2]));
    if (C.p.dthresh - dist > 0) return false;
    gi = pr.next;
  }
  return true;
}
