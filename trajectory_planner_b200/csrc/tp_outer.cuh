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
  AStarSmem* sm;       // shared: heap arrays, per-axis tables, staging
  double* heap_k_gl;   // HBM spill for heap entries >= TP_HEAP_SMEM
  uint32_t* heap_n_gl;
  double* path;        // xyz, path_cap points
  double* sc;          // max_seg x TP_SC_CAP x 3
  int* sc_len;         // max_seg
  uint32_t* round_ptr;
  int lane;
  int flood_trigger;      // expansions after which astar_search checks reachability (adaptive per trajectory)
  int goal_unreachable;   // set by astar_search when the flood fill proved the goal unreachable
  int flood_wasted;       // consecutive floods of this trajectory that found the goal reachable (nothing gained)
  int* flood_stats;       // engine-wide outcome counts of FIRST reachability checks (AStarPools::flood_stats)
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

// ---- shared-memory accessors on 32-bit shared-window addresses (keeps the sift loops free of
// generic-pointer arithmetic)
__device__ __forceinline__ double lds_f64(uint32_t a) {
  double v;
  asm volatile("ld.shared.f64 %0, [%1];" : "=d"(v) : "r"(a));
  return v;
}
// key as its bit pattern: fScore >= 0 always, and for non-negative doubles the IEEE order is the order of the
// 64-bit patterns — the sift comparisons run on the integer pipe
__device__ __forceinline__ unsigned long long lds_k64(uint32_t a) {
  unsigned long long v;
  asm volatile("ld.shared.u64 %0, [%1];" : "=l"(v) : "r"(a));
  return v;
}
// six keys at once (children + grandchildren of a heap node): ONE asm statement, so that all six loads are issued
// before the first comparison consumes any of them
__device__ __forceinline__ void lds_k64x6(uint32_t a0, uint32_t a1, uint32_t a2, uint32_t a3, uint32_t a4, uint32_t a5,
                                          unsigned long long& v0, unsigned long long& v1, unsigned long long& v2,
                                          unsigned long long& v3, unsigned long long& v4, unsigned long long& v5) {
  asm volatile(
      "ld.shared.u64 %0, [%6];\n\tld.shared.u64 %1, [%7];\n\tld.shared.u64 %2, [%8];\n\t"
      "ld.shared.u64 %3, [%9];\n\tld.shared.u64 %4, [%10];\n\tld.shared.u64 %5, [%11];"
      : "=l"(v0), "=l"(v1), "=l"(v2), "=l"(v3), "=l"(v4), "=l"(v5)
      : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(a4), "r"(a5));
}
__device__ __forceinline__ uint32_t lds_u32(uint32_t a) {
  uint32_t v;
  asm volatile("ld.shared.u32 %0, [%1];" : "=r"(v) : "r"(a));
  return v;
}
__device__ __forceinline__ void sts_f64(uint32_t a, double v) { asm volatile("st.shared.f64 [%0], %1;" ::"r"(a), "d"(v) : "memory"); }
__device__ __forceinline__ void sts_u32(uint32_t a, uint32_t v) { asm volatile("st.shared.u32 [%0], %1;" ::"r"(a), "r"(v) : "memory"); }

// packed node id: (i << 16) | (j << 8) | kk, kk = k - k_lo (kk == 255: the spare slot of the start cell)
#define AS_SPARE_KK 255u

// ---- libstdc++ heap mechanics (std::priority_queue without decrease-key, astarOcc.cpp:150-228).
// comp(a,b) = a->fScore > b->fScore with keys read at comparison time; here the key is cached beside
// the node id and patched whenever the reference mutates fScore in place, so every comparison sees
// exactly the value the reference would read through the node pointer.
struct Heap {
  uint32_t sk, sn;     // shared-window addresses of the key / id arrays
  double* gk;          // HBM spill (entries >= TP_HEAP_SMEM)
  uint32_t* gn;
  ANode* nodes;
  uint32_t PY, KL, spare;
  int size;
  __device__ __forceinline__ uint32_t lin_of(uint32_t id) const {
    const uint32_t kk = id & 255u;
    return kk == AS_SPARE_KK ? spare : ((id >> 16) * PY + ((id >> 8) & 255u)) * KL + kk;
  }
  __device__ __forceinline__ double key(int i) const { return i < TP_HEAP_SMEM ? lds_f64(sk + 8u * i) : gk[i - TP_HEAP_SMEM]; }
  __device__ __forceinline__ uint32_t node(int i) const { return i < TP_HEAP_SMEM ? lds_u32(sn + 4u * i) : gn[i - TP_HEAP_SMEM]; }
  __device__ __forceinline__ void set(int i, double k, uint32_t n) {
    if (i < TP_HEAP_SMEM) { sts_f64(sk + 8u * i, k); sts_u32(sn + 4u * i, n); }
    else { gk[i - TP_HEAP_SMEM] = k; gn[i - TP_HEAP_SMEM] = n; }
    nodes[lin_of(n)].heap_pos = (uint32_t)i;
  }
  __device__ __forceinline__ void set_key(int i, double k) {
    if (i < TP_HEAP_SMEM) sts_f64(sk + 8u * i, k); else gk[i - TP_HEAP_SMEM] = k;
  }
  // all-shared fast variant (valid while every index touched is < TP_HEAP_SMEM)
  __device__ __forceinline__ void set_s(int i, double k, uint32_t n) {
    sts_f64(sk + 8u * i, k);
    sts_u32(sn + 4u * i, n);
    nodes[lin_of(n)].heap_pos = (uint32_t)i;
  }
  // std::__push_heap
  __device__ __forceinline__ void sift_up(int hole, double vk, uint32_t vn) {
    if (hole < TP_HEAP_SMEM) {
      const unsigned long long vki = (unsigned long long)__double_as_longlong(vk);   // fScore >= 0: integer order
      while (hole > 0) {
        const int parent = (hole - 1) >> 1;
        const unsigned long long pki = lds_k64(sk + 8u * parent);
        const uint32_t pn = lds_u32(sn + 4u * parent);
        if (!(pki > vki)) break;
        set_s(hole, __longlong_as_double((long long)pki), pn);
        hole = parent;
      }
      set_s(hole, vk, vn);
      return;
    }
    while (hole > 0) {
      const int parent = (hole - 1) >> 1;
      const double pk = key(parent);
      if (!(pk > vk)) break;
      set(hole, pk, node(parent));
      hole = parent;
    }
    set(hole, vk, vn);
  }
  __device__ __forceinline__ void push(uint32_t n, double f) {
    sift_up(size, f, n);
    ++size;
  }
  // top() + std::pop_heap (= __adjust_heap on the hole at the root, then __push_heap) + pop_back
  __device__ __forceinline__ uint32_t pop() {
    const uint32_t top = node(0);
    if (size > 1) {
      const int len = size - 1;
      const double vk = key(len);
      const uint32_t vn = node(len);
      int hole = 0, second = 0;
      const int lim = (len - 1) / 2;
      if (size <= TP_HEAP_SMEM) {
        // two levels per step: the keys / ids of both children AND of all four grandchildren are loaded together
        // (12 independent LDS), then the two comparisons of std::__adjust_heap are resolved from registers
        while (second < lim) {
          const int r1 = 2 * (second + 1), l1 = r1 - 1;
          const int rr = 2 * (r1 + 1), lr = 2 * r1;             // right children of r1 / l1 (left ones: - 1)
          const bool spec = rr < TP_HEAP_SMEM;
          const double kr1 = lds_f64(sk + 8u * r1), kl1 = lds_f64(sk + 8u * l1);
          const uint32_t nr1 = lds_u32(sn + 4u * r1), nl1 = lds_u32(sn + 4u * l1);
          double krr = 0, krl = 0, klr = 0, kll = 0;
          uint32_t nrr = 0, nrl = 0, nlr = 0, nll = 0;
          if (spec) {
            krr = lds_f64(sk + 8u * rr); krl = lds_f64(sk + 8u * rr - 8u);
            klr = lds_f64(sk + 8u * lr); kll = lds_f64(sk + 8u * lr - 8u);
            nrr = lds_u32(sn + 4u * rr); nrl = lds_u32(sn + 4u * rr - 4u);
            nlr = lds_u32(sn + 4u * lr); nll = lds_u32(sn + 4u * lr - 4u);
          }
          const bool left1 = kr1 > kl1;
          second = left1 ? l1 : r1;
          set_s(hole, left1 ? kl1 : kr1, left1 ? nl1 : nr1);
          hole = second;
          if (spec && second < lim) {
            const int r2 = left1 ? lr : rr;
            const double kr2 = left1 ? klr : krr, kl2 = left1 ? kll : krl;
            const uint32_t nr2 = left1 ? nlr : nrr, nl2 = left1 ? nll : nrl;
            const bool left2 = kr2 > kl2;
            second = left2 ? r2 - 1 : r2;
            set_s(hole, left2 ? kl2 : kr2, left2 ? nl2 : nr2);
            hole = second;
          }
        }
      } else {
        while (second < lim) {
          second = 2 * (second + 1);
          double kr = key(second);
          const double kl = key(second - 1);
          if (kr > kl) { --second; kr = kl; }
          set(hole, kr, node(second));
          hole = second;
        }
      }
      if ((len & 1) == 0 && second == (len - 2) / 2) {
        second = 2 * (second + 1);
        set(hole, key(second - 1), node(second - 1));
        hole = second - 1;
      }
      sift_up(hole, vk, vn);
    }
    --size;
    return top;
  }
  // Warp-cooperative pop_heap (all 32 lanes call it; the caller has read the top already).  std::__adjust_heap's
  // DECISIONS — which child moves up at each level — depend only on keys the sift-down never rewrites, so every lane
  // walks the root-to-leaf path redundantly (two levels per step from one batch of speculative loads) and lane l keeps
  // the move of level l; the moves themselves (key, id, the node record's heap_pos: address arithmetic and a global
  // store each) are then applied by their lanes in parallel instead of one after the other on lane 0.
  // Same final heap as pop(): bit-identical A* behaviour.
  __device__ __forceinline__ void pop_warp(int lane) {
    const int sz = __shfl_sync(0xffffffffu, size, 0);
    if (sz > TP_HEAP_SMEM) {   // spilled heap: sequential path (warp-uniform branch)
      if (lane == 0) pop();
      return;
    }
    if (sz > 1) {
      const int len = sz - 1;
      const double vk = lds_f64(sk + 8u * len);
      const uint32_t vn = lds_u32(sn + 4u * len);
      int hole = 0, second = 0, level = 0, my_from = -1, my_to = -1;
      const int lim = (len - 1) / 2;
      while (second < lim) {
        const int r1 = 2 * (second + 1), l1 = r1 - 1;
        const int rr = 2 * (r1 + 1), lr = 2 * r1;
        const bool spec = rr < TP_HEAP_SMEM;
        unsigned long long kr1, kl1, krr = 0, krl = 0, klr = 0, kll = 0;
        if (spec) {
          lds_k64x6(sk + 8u * r1, sk + 8u * l1, sk + 8u * rr, sk + 8u * rr - 8u, sk + 8u * lr, sk + 8u * lr - 8u, kr1, kl1, krr,
                    krl, klr, kll);
        } else {
          kr1 = lds_k64(sk + 8u * r1);
          kl1 = lds_k64(sk + 8u * l1);
        }
        const bool left1 = kr1 > kl1;
        second = left1 ? l1 : r1;
        if (lane == level) { my_from = second; my_to = hole; }
        hole = second;
        ++level;
        if (spec && second < lim) {
          const int r2 = left1 ? lr : rr;
          const bool left2 = (left1 ? klr : krr) > (left1 ? kll : krl);
          second = left2 ? r2 - 1 : r2;
          if (lane == level) { my_from = second; my_to = hole; }
          hole = second;
          ++level;
        }
      }
      if ((len & 1) == 0 && second == (len - 2) / 2) {
        second = 2 * (second + 1);
        if (lane == level) { my_from = second - 1; my_to = hole; }
        hole = second - 1;
        ++level;
      }
      double mk = 0;
      uint32_t mn = 0;
      if (my_from >= 0) { mk = lds_f64(sk + 8u * my_from); mn = lds_u32(sn + 4u * my_from); }
      __syncwarp();
      if (my_from >= 0) set_s(my_to, mk, mn);
      __syncwarp();
      if (lane == 0) sift_up(hole, vk, vn);
    }
    if (lane == 0) --size;
  }
};

// Reachability shortcut for searches that are about to exhaust their pool.  AstarSearch pops every
// cell reachable from the start exactly once before it gives up (nodes are pushed once and never
// re-opened, astarOcc.cpp:209-228), so when the goal is NOT in the start's connected component the
// search's outcome (failure) and its expansion count (= component size) are known without running
// it.  The component is flooded with a lane-per-cell worklist over two bitmaps that temporarily
// take over the heap's shared memory (the heap is parked in the tail of its HBM spill area).
// Returns -1 when the goal is reachable (the search resumes), else the component size.
#ifndef TP_FLOOD_TRIGGER
#define TP_FLOOD_TRIGGER 2048       // expansions after which a search checks reachability (a flood costs about as much as 500)
#endif
#define TP_FLOOD_TRIGGER_AGAIN 192  // ... once this trajectory has already had an unreachable goal
#ifndef TP_FLOOD_TRIGGER_EARLY
#define TP_FLOOD_TRIGGER_EARLY 512  // first check of a search on a map where such checks mostly prove the goal unreachable
#endif
// A flood costs about as much as 250 expansions and, when it proves the goal unreachable after 512 instead of 2 048
// expansions, saves 1 536 of them: checking early pays once one early check in seven succeeds.  The engine counts the
// outcomes of first checks (maze.bt: most goals are unreachable -> 512; square_static_map: most long searches arrive -> 2 048).
// Neither the outcome nor the expansion count of a search depends on when it checks.
__device__ __forceinline__ int flood_first_trigger(const int* stats) {
  const int unreach = *((volatile const int*)&stats[0]), reach = *((volatile const int*)&stats[1]);
  return 7 * unreach >= reach ? TP_FLOOD_TRIGGER_EARLY : TP_FLOOD_TRIGGER;
}
#define TP_FLOOD_ROUND (32 * 26)    // cells one round can add to the worklist
// The flood runs on ONE shared-memory bitmap F over the pool's stored cells (bit (i PY + j) KL + kk): first every
// enterable free cell is marked (one coalesced pass over the map words of the window, lane per (i, j) column — no map
// gather inside the flood itself), then a lane-per-cell worklist claims neighbours by clearing their bits
// (atomicAnd: a cell is enqueued exactly once).  F, the counters and the worklist ring take over the heap's shared
// memory; the heap is parked in the tail of its HBM spill area meanwhile.
__device__ __noinline__ int flood_component(const DevMap& map, const VigoConst& C, Worker& W, int heap_size, int si, int sj, int sk,
                               int ei, int ej, int ek, int k_lo) {
  AStarSmem& S = *W.sm;
  const int lane = W.lane;
  const int PX = C.pool[0], PY = C.pool[1], PZ = C.pool[2], KL = C.pool_kl;
  const int ncell = PX * PY * KL;
  const int nw = (ncell + 31) >> 5;
  uint32_t* F = reinterpret_cast<uint32_t*>(S.hk);   // [nw + 1] (one spare word: two-word window reads)
  uint32_t* Q = F + nw + 1 + 4;                       // worklist ring of packed cells (i << 16 | j << 8 | kk)
  const int qcap = (int)((sizeof(S.hk) + sizeof(S.hn)) / 4) - nw - 5;
  if (qcap < 2 * TP_FLOOD_ROUND || KL > 30 || C.heap_cap < 3 * TP_HEAP_SMEM) return -1;
  // park the heap
  const int keep = heap_size < TP_HEAP_SMEM ? heap_size : TP_HEAP_SMEM;
  double* pk = W.heap_k_gl + (C.heap_cap - TP_HEAP_SMEM);
  uint32_t* pn = W.heap_n_gl + (C.heap_cap - TP_HEAP_SMEM);
  for (int i = lane; i < keep; i += 32) { pk[i] = S.hk[i]; pn[i] = S.hn[i]; }
  __syncwarp();
  for (int i = lane; i < nw + 5; i += 32) F[i] = 0u;
  __syncwarp();
  const int klo_c = k_lo > 1 ? k_lo : 1;                                  // neighbour layers that can be entered
  const int khi_c = (k_lo + KL - 1) < (PZ - 2) ? (k_lo + KL - 1) : (PZ - 2);
#ifdef TP_ASTAR_TIMING
  const long long tz0 = clock64();
#endif
  // ---- F = enterable (astarOcc.cpp:181-202: inside the pool's interior and the height band) and free in the map,
  // the start cell excluded (it is never re-opened).  Per layer kk: the map's z index and whether the layer can be
  // entered at all (lane-invariant, so evaluated once); fast path when the layers map to consecutive bits of one word.
  uint32_t layer_ok = 0;
  int iz0 = -1;
  bool affine = true;
  for (int kk = 0; kk < KL; ++kk) {
    const int nk = kk + k_lo;
    if (nk < klo_c || nk > khi_c || !S.band[nk] || S.tz[nk] < 0) continue;
    layer_ok |= 1u << kk;
    const int iz = S.tz[nk];
    if (iz0 < 0) iz0 = iz - kk;
    if (iz - kk != iz0) affine = false;
  }
  if (iz0 < 0 || (iz0 >> 5) != ((iz0 + KL - 1) >> 5) || iz0 < 0) affine = false;
  const int ncol = PX * PY;
  int ci0 = lane / PY, cj0 = lane - ci0 * PY;   // (i, j) of column c0, advanced without divisions
  for (int c0 = lane; c0 < ncol; c0 += 128) {
    uint32_t wv[4], bits[4];
    int cols[4];
    int ii = ci0, jj = cj0;
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const int col = c0 + 32 * u;
      cols[u] = -1;
      wv[u] = 0xFFFFFFFFu;
      bits[u] = 0u;
      const int i = ii, j = jj;
      jj += 32;
      while (jj >= PY) { jj -= PY; ++ii; }
      if (col >= ncol || i < 1 || i >= PX - 1 || j < 1 || j >= PY - 1) continue;
      const int ix = S.tx[i], iy = S.ty[j];
      if (ix < 0 || iy < 0) continue;
      cols[u] = col;
      if (affine) {
        wv[u] = __ldg(map.inflated + ((size_t)ix * map.dim[1] + iy) * map.wz + (iz0 >> 5));
      } else {
        const uint32_t* colw = map.inflated + ((size_t)ix * map.dim[1] + iy) * map.wz;
        for (int kk = 0; kk < KL; ++kk) {
          if (!((layer_ok >> kk) & 1u)) continue;
          const int iz = S.tz[kk + k_lo];
          if (!((__ldg(colw + (iz >> 5)) >> (iz & 31)) & 1u)) bits[u] |= 1u << kk;
        }
      }
      if (i == si && j == sj && sk - k_lo >= 0 && sk - k_lo < KL) cols[u] |= 0x40000000;   // the start's column
    }
    ci0 = ii; cj0 = jj;
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      // unconditional (possibly empty) ORs: eight shared-memory atomics in flight instead of eight divergent regions
      uint32_t bt = 0u;
      int b0 = 0;
      if (cols[u] >= 0) {
        bt = affine ? ((~wv[u] >> (iz0 & 31)) & layer_ok) : bits[u];
        if (cols[u] & 0x40000000) bt &= ~(1u << (sk - k_lo));
        b0 = (cols[u] & 0x3FFFFFFF) * KL;
      }
      const int sh = b0 & 31;
      atomicOr(&F[b0 >> 5], bt << sh);
      atomicOr(&F[(b0 >> 5) + 1], sh ? bt >> (32 - sh) : 0u);
    }
  }
  __syncwarp();
#ifdef TP_ASTAR_TIMING
  const long long tb0 = clock64();
  int n_rounds = 0;
#endif
  // ring index pos % qcap without an integer division (pos < 2^24: the float quotient is off by at most one)
  const float inv_qcap = 1.0f / (float)qcap;
  auto ring = [&](uint32_t pos) {
    const uint32_t q = __float2uint_rz(__uint2float_rn(pos) * inv_qcap);
    int r = (int)pos - (int)(q * (uint32_t)qcap);
    if (r >= qcap) r -= qcap;
    if (r < 0) r += qcap;
    return (uint32_t)r;
  };
  // One round = up to 32 worklist cells, lane per cell: read the <= 3 candidate bits of each of the 9 neighbour
  // columns through a two-word window, claim them with one atomicAnd per word (a cell is claimed by exactly one lane),
  // then append the claimed cells to the ring at positions from a warp prefix sum (the worklist tail is warp-private).
  uint32_t head = 0, tail = 0, claimed = 0;
  auto round = [&](int ci, int cj, int ck, bool valid) {
    uint32_t got[9];
    int lo = 0, mine = 0;
#pragma unroll
    for (int q = 0; q < 9; ++q) got[q] = 0u;
    if (valid) {
      const int kc = ck - k_lo;
      lo = kc - 1 > 0 ? kc - 1 : 0;
      const int hi = kc + 1 < KL - 1 ? kc + 1 : KL - 1;
      if (lo <= hi) {
        const uint32_t wmask = (1u << (hi - lo + 1)) - 1u;
        uint32_t v[9];
        int bp[9];
#pragma unroll
        for (int q = 0; q < 9; ++q) {
          const int ni = ci + q / 3 - 1, nj = cj + q % 3 - 1;
          v[q] = 0u;
          bp[q] = 0;
          if (ni < 0 || ni >= PX || nj < 0 || nj >= PY) continue;
          bp[q] = (ni * PY + nj) * KL + lo;
          v[q] = __funnelshift_r(F[bp[q] >> 5], F[(bp[q] >> 5) + 1], bp[q] & 31) & wmask;
        }
        // claims: one (possibly empty) atomicAnd per word touched, all 18 in flight before any result is used
        uint32_t o0[9], o1[9];
#pragma unroll
        for (int q = 0; q < 9; ++q) {
          const int sh = bp[q] & 31;
          o0[q] = atomicAnd(&F[bp[q] >> 5], ~(v[q] << sh));
          o1[q] = atomicAnd(&F[(bp[q] >> 5) + 1], ~(sh ? v[q] >> (32 - sh) : 0u));
        }
#pragma unroll
        for (int q = 0; q < 9; ++q) {
          const int sh = bp[q] & 31;
          const uint32_t m0 = v[q] << sh, m1 = sh ? v[q] >> (32 - sh) : 0u;
          const uint32_t c = ((o0[q] & m0) >> sh) | (sh ? (o1[q] & m1) << (32 - sh) : 0u);
          got[q] = c;
          mine += __popc(c);
        }
      }
    }
    // exclusive prefix sum of `mine` over the warp
    int incl = mine;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const int up = __shfl_up_sync(0xffffffffu, incl, o);
      if (lane >= o) incl += up;
    }
    uint32_t pos = ring(tail + (uint32_t)(incl - mine));
    const uint32_t got_all = (uint32_t)__shfl_sync(0xffffffffu, incl, 31);
    tail += got_all;
    claimed += got_all;
    // one loop over all claimed bits of the lane (bit 3 q + r = neighbour column q, layer lo + r)
    uint32_t all = 0u;
#pragma unroll
    for (int q = 0; q < 9; ++q) all |= got[q] << (3 * q);
    while (all) {
      const int t = __ffs(all) - 1;
      all &= all - 1;
      const int q = (t * 11) >> 5, r = t - 3 * q;       // t / 3, t % 3 for t < 27
      const int qx = (q * 11) >> 5, qy = q - 3 * qx;
      Q[pos] = ((uint32_t)(ci + qx - 1) << 16) | ((uint32_t)(cj + qy - 1) << 8) | (uint32_t)(lo + r);
      pos = pos + 1 == (uint32_t)qcap ? 0u : pos + 1;
    }
    __syncwarp();
  };
  round(si, sj, sk, lane == 0);
  bool bail = false;
  // The ring holds ~3 300 cells; a wide front (open rooms) outgrows it.  Order is irrelevant for a flood, so the oldest
  // half of a full ring is moved to an overflow STACK in the heap's HBM spill area (below the parked heap) and fetched
  // back when the ring runs dry.
  // (above the heap entries that already live in the spill area: those beyond TP_HEAP_SMEM)
  const int spilled = heap_size > TP_HEAP_SMEM ? heap_size - TP_HEAP_SMEM : 0;
  uint32_t* ovf = reinterpret_cast<uint32_t*>(W.heap_k_gl + spilled);
  const int ovf_cap = (C.heap_cap - TP_HEAP_SMEM - spilled) * 2;
  int ovf_n = 0, n_spills = 0;
  for (;;) {
    int n = (int)(tail - head);
    if (n == 0) {
      if (ovf_n == 0) break;
      const int m = ovf_n < qcap / 2 ? ovf_n : qcap / 2;
      for (int i = lane; i < m; i += 32) Q[ring(tail + (uint32_t)i)] = ovf[ovf_n - m + i];
      tail += (uint32_t)m;
      ovf_n -= m;
      __syncwarp();
      continue;
    }
    if (n > qcap - TP_FLOOD_ROUND) {   // the ring could overflow during this round
      const int m = n / 2;
      if (ovf_n + m > ovf_cap) { bail = true; break; }
      for (int i = lane; i < m; i += 32) ovf[ovf_n + i] = Q[ring(head + (uint32_t)i)];
      head += (uint32_t)m;
      ovf_n += m;
      ++n_spills;
      __syncwarp();
      n -= m;
    }
    if (n > 32) n = 32;
#ifdef TP_ASTAR_TIMING
    ++n_rounds;
#endif
    uint32_t cell = 0xFFFFFFFFu;
    if (lane < n) cell = Q[ring(head + (uint32_t)lane)];
    head += (uint32_t)n;
    round((int)(cell >> 16), (int)((cell >> 8) & 255u), (int)(cell & 255u) + k_lo, cell != 0xFFFFFFFFu);
  }
#ifdef TP_ASTAR_TIMING
  if (lane == 0) printf("[flood-detail] build %lld bfs %lld rounds %d spills %d bail %d cells %u\n", tb0 - tz0, clock64() - tb0, n_rounds, n_spills, (int)bail, claimed);
#endif
  (void)n_spills;
  const int count = (int)claimed + 1;  // every free cell was claimed exactly once, plus the start cell
  // is the goal in the component?  (claimed cells have their bit cleared: ask the map whether it was free at all)
  bool reach = false;
  if (!bail) {
    if (ei == si && ej == sj && ek == sk) reach = true;
    else if (ek >= klo_c && ek <= khi_c && ei >= 1 && ei < PX - 1 && ej >= 1 && ej < PY - 1 && S.band[ek]) {
      const int b = (ei * PY + ej) * KL + (ek - k_lo);
      if (!((F[b >> 5] >> (b & 31)) & 1u)) {   // not (or no longer) marked: blocked, or free and reached
        const int ix = S.tx[ei], iy = S.ty[ej], iz = S.tz[ek];
        bool blocked = true;
        if (ix >= 0 && iy >= 0 && iz >= 0) {
          const uint32_t w = __ldg(&map.inflated[((size_t)ix * map.dim[1] + iy) * map.wz + (iz >> 5)]);
          blocked = (w >> (iz & 31)) & 1u;
        }
        reach = !blocked;
      }
    }
  }
  __syncwarp();
  // restore the heap
  for (int i = lane; i < keep; i += 32) { S.hk[i] = pk[i]; S.hn[i] = pn[i]; }
  __syncwarp();
  return (bail || reach) ? -1 : count;
}

// AStar::AstarSearch + getPath.  Returns the number of path points written to W.path (cell centres
// start -> goal) or -1.  All 32 lanes call it with identical arguments.
//
// Node storage: only the z layers whose cell centre can lie inside [min_height, max_height] are
// stored (pool_kl of them, starting at k_lo) — cells outside the band are rejected before any node
// state is consulted (astarOcc.cpp:202) — plus one spare slot for the start cell, which the
// reference expands from even when it lies outside the band.  Per expansion the warp makes ONE
// round of global loads (current node, <= 26 neighbour nodes as LDG.128, <= 26 map words, all in
// flight together); the open-set heap, the per-axis map-index tables and the serial pass's staging
// live in shared memory; no FP64 division or integer division is on the per-expansion path.
__device__ __noinline__ int astar_search(const DevMap& map, const VigoConst& C, Worker& W, const D3& start_in, const D3& end_in,
                            int& expansions, int& err) {
  const int lane = W.lane;
  AStarSmem& S = *W.sm;
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
  // this search's reachability check: the trajectory's own setting when it has history, else the engine-wide choice
  const bool first_check = W.flood_trigger == TP_FLOOD_TRIGGER;
  const int flood_at = first_check ? flood_first_trigger(W.flood_stats) : W.flood_trigger;
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
  const int PX = C.pool[0], PY = C.pool[1], PZ = C.pool[2], KL = C.pool_kl;
  const int k_lo = F.k_lo;
  // ---- per-axis tables: map cell of pool index (Index2Coord -> posToIndex, evaluated once per axis)
  for (int t = lane; t < PX; t += 32) {
    const double f = floor(((double)(t - PX / 2) * F.step + F.center.x - map.mn[0]) / map.res);
    S.tx[t] = (f >= 0.0 && f < (double)map.dim[0]) ? (short)(int)f : (short)-1;
  }
  for (int t = lane; t < PY; t += 32) {
    const double f = floor(((double)(t - PY / 2) * F.step + F.center.y - map.mn[1]) / map.res);
    S.ty[t] = (f >= 0.0 && f < (double)map.dim[1]) ? (short)(int)f : (short)-1;
  }
  for (int t = lane; t < PZ; t += 32) {
    const double z = (double)(t - PZ / 2) * F.step + F.center.z;
    const double f = floor((z - map.mn[2]) / map.res);
    S.tz[t] = (f >= 0.0 && f < (double)map.dim[2]) ? (short)(int)f : (short)-1;
    S.band[t] = !(z > C.p.max_height || z < C.p.min_height) ? 1 : 0;
  }
  const uint32_t spare = (uint32_t)((size_t)PX * PY * KL);
  const uint32_t start_id = ((uint32_t)si << 16) | ((uint32_t)sj << 8) | AS_SPARE_KK;
  ANode* nodes = W.nodes;
  Heap H;
  H.sk = (uint32_t)__cvta_generic_to_shared(S.hk);
  H.sn = (uint32_t)__cvta_generic_to_shared(S.hn);
  H.gk = W.heap_k_gl;
  H.gn = W.heap_n_gl;
  H.nodes = nodes;
  H.PY = (uint32_t)PY;
  H.KL = (uint32_t)KL;
  H.spare = spare;
  H.size = 0;
  if (lane == 0) {
    ANode nd;
    nd.stamp_state = (round << 2) | ST_OPEN;
    nd.parent = NODE_NONE;
    nd.g = 0;
    nd.heap_pos = 0; nd.pad0 = 0; nd.pad1 = 0;
    nodes[spare] = nd;
    H.push(start_id, as_heu(si, sj, sk, ei, ej, ek));
  }
  __syncwarp();
  const double gstep1 = 1.0, gstep2 = sqrt(2.0), gstep3 = sqrt(3.0);
  int result = -1;
  int num_iter = 0;
  uint32_t goal_id = NODE_NONE;
#ifdef TP_ASTAR_TIMING
  long long tA = 0, tB = 0, tC = 0, tPop = 0, t0 = clock64(), t1;
#define TICK(acc) { t1 = clock64(); acc += t1 - t0; t0 = t1; }
#else
#define TICK(acc)
#endif
  for (;;) {
    // ---- next node = top of the open set.  Its neighbours' node records and map words are requested FIRST (one
    // round of global loads in flight), then lane 0 runs pop_heap's sift-down in shared memory under that latency.
    // (The sift-down moves heap entries, i.e. rewrites heap_pos fields: in-place updates below re-read theirs.)
    uint32_t cur = NODE_NONE;
    if (lane == 0 && H.size > 0) cur = H.node(0);
    cur = __shfl_sync(0xffffffffu, cur, 0);
    if (cur == NODE_NONE) break;  // open set empty
    ++num_iter;
    const int ci = (int)(cur >> 16), cj = (int)((cur >> 8) & 255u);
    const int ck = (cur & 255u) == AS_SPARE_KK ? sk : (int)(cur & 255u) + k_lo;
    if (ci == ei && cj == ej && ck == ek) {
      if (lane == 0) H.pop();
      goal_id = cur;
      result = 0;
      break;
    }
    const uint32_t cur_lin = H.lin_of(cur);
    const double gcur = nodes[cur_lin].g;  // uniform address: one broadcast load, overlaps the loads below
    // ---- lane-parallel neighbour loads (astarOcc.cpp:173-229); lane L <-> (dx,dy,dz) in the reference's loop order
    int kind = 0;  // 0 skip, 1 push (new node), 2 in-place update
    bool cand = false, is_start = false, band = false;
    int ni = 0, nj = 0, nk = 0, d2 = 0;
    uint32_t nl = 0, nid = 0;
    uint4 raw = make_uint4(0u, 0u, 0u, 0u);
    uint32_t mapw = 0xFFFFFFFFu;
    int mapbit = 0;
    if (lane < 27 && lane != 13) {
      const int dx = lane / 9 - 1, dy = (lane / 3) % 3 - 1, dz = lane % 3 - 1;
      ni = ci + dx; nj = cj + dy; nk = ck + dz;
      d2 = dx * dx + dy * dy + dz * dz;
      const bool inb = !(ni < 1 || ni >= PX - 1 || nj < 1 || nj >= PY - 1 || nk < 1 || nk >= PZ - 1);
      if (inb) {
        is_start = (ni == si && nj == sj && nk == sk);
        band = S.band[nk] != 0;
        const bool layer_ok = nk >= k_lo && nk < k_lo + KL;
        if (band && !layer_ok && !is_start) err |= ERR_BAND;  // cannot happen (pool_kl has slack)
        // cells outside the band are rejected by :202 whatever their node state says, except that a
        // CLOSED start node is skipped one line earlier — same outcome (skip) either way.
        if ((band && layer_ok) || is_start) {
          cand = true;
          nid = is_start ? start_id : (((uint32_t)ni << 16) | ((uint32_t)nj << 8) | (uint32_t)(nk - k_lo));
          nl = is_start ? spare : (uint32_t)((ni * PY + nj) * KL + (nk - k_lo));
          raw = *reinterpret_cast<const uint4*>(&nodes[nl]);  // stamp_state, parent, g
          if (band) {
            const int ix = S.tx[ni], iy = S.ty[nj], iz = S.tz[nk];
            if (ix >= 0 && iy >= 0 && iz >= 0) {
              mapw = __ldg(&map.inflated[((size_t)ix * map.dim[1] + iy) * map.wz + (iz >> 5)]);
              mapbit = iz & 31;
            }
          }
        }
      }
    }
    // ---- pop (lane 0, shared memory) while the loads are in flight
#ifdef TP_ASTAR_TIMING
    const long long tp0 = clock64();
#endif
    H.pop_warp(lane);
    if (lane == 0) nodes[cur_lin].stamp_state = (round << 2) | ST_CLOSED;
    __syncwarp();
#ifdef TP_ASTAR_TIMING
    tPop += clock64() - tp0;
    if ((raw.x ^ mapw) == 0x12345679u && gcur == 1.2345) err |= ERR_BAND;   // forces the loads to have landed
    __syncwarp();
#endif
    TICK(tA)
    // ---- classification
    if (cand) {
      const bool blocked_map = (mapw >> mapbit) & 1u;
      const bool explored = (raw.x >> 2) == round;
      const uint32_t state = raw.x & 3u;
      if (!(explored && state == ST_CLOSED)) {
        if (blocked_map) {
          // blocked this round: remember it so the map is not queried again (observably the same
          // as the reference's stale-state handling: the cell is skipped on every visit)
          nodes[nl].stamp_state = (round << 2) | ST_CLOSED;
        } else {
          const double tentative = gcur + (d2 == 1 ? gstep1 : (d2 == 2 ? gstep2 : gstep3));
          const double gold = __hiloint2double((int)raw.w, (int)raw.z);
          if (!explored) {
            kind = 1;
            uint4 w;
            w.x = (round << 2) | ST_OPEN;
            w.y = cur;
            w.z = (uint32_t)__double2loint(tentative);
            w.w = (uint32_t)__double2hiint(tentative);
            *reinterpret_cast<uint4*>(&nodes[nl]) = w;
          } else if (tentative < gold) {
            kind = 2;
          }
          if (kind) {
            S.st_id[lane] = nid;
            S.st_g[lane] = tentative;
            S.st_f[lane] = tentative + as_heu(ni, nj, nk, ei, ej, ek);
          }
        }
      }
    }
    __syncwarp();
    // ---- serial pass in the reference's neighbour order: in-place key updates and pushes
    unsigned mask_new = __ballot_sync(0xffffffffu, kind == 1);
    unsigned mask_upd = __ballot_sync(0xffffffffu, kind == 2);
    TICK(tB)
    int overflow = 0;
    if (lane == 0) {
      // in-place updates read the entry's CURRENT heap position from the node record: the pop and this expansion's
      // earlier pushes may have moved it
      unsigned m = mask_new | mask_upd;
      while (m) {
        const int L = __ffs(m) - 1;
        m &= m - 1;
        const uint32_t nid = S.st_id[L];
        if ((mask_upd >> L) & 1u) {
          const uint32_t nl = H.lin_of(nid);
          nodes[nl].parent = cur;
          nodes[nl].g = S.st_g[L];
          const uint32_t pos = nodes[nl].heap_pos;   // current position (the pop and earlier pushes may have moved it)
          H.set_key((int)pos, S.st_f[L]);
        } else {
          if (H.size >= C.heap_cap) { overflow = 1; break; }
          H.push(nid, S.st_f[L]);
        }
      }
    }
    __syncwarp();
    TICK(tC)
    if (__any_sync(0xffffffffu, overflow != 0 || (err & ERR_BAND) != 0)) {
      err |= ERR_HEAP_OVERFLOW;
      break;
    }
    if (C.p.astar_max_expansions > 0 && num_iter >= C.p.astar_max_expansions) break;
    if (num_iter == flood_at) {
      const int hs = __shfl_sync(0xffffffffu, H.size, 0);
#ifdef TP_ASTAR_TIMING
      const long long tf0 = clock64();
#endif
      const int comp = flood_component(map, C, W, hs, si, sj, sk, ei, ej, ek, k_lo);
#ifdef TP_ASTAR_TIMING
      if (lane == 0) printf("[flood] result %d cycles %lld\n", comp, clock64() - tf0);
      t0 = clock64();
#endif
      W.flood_wasted = comp >= 0 ? 0 : W.flood_wasted + 1;
      if (first_check && lane == 0) {
        const int idx = comp >= 0 ? 0 : 1;
        if (atomicAdd(&W.flood_stats[idx], 1) > (1 << 20)) { atomicSub(&W.flood_stats[0], W.flood_stats[0] / 2); atomicSub(&W.flood_stats[1], W.flood_stats[1] / 2); }
      }
      if (comp >= 0) {  // goal unreachable: the search would pop the whole component and fail
        W.goal_unreachable = 1;
        num_iter = (C.p.astar_max_expansions > 0 && comp > C.p.astar_max_expansions) ? C.p.astar_max_expansions : comp;
        break;
      }
    }
  }
  expansions = num_iter;
#ifdef TP_ASTAR_TIMING
  if (lane == 0)
    printf("[astar] exp %d heap %d cycles/exp: sift %.0f pop+loadwait %.0f classify %.0f push %.0f\n", num_iter, H.size, (double)tPop / num_iter, (double)tA / num_iter,
           (double)tB / num_iter, (double)tC / num_iter);
#endif
  if (result < 0) return -1;
  // ---- retrievePath + getPath (astarOcc.cpp:77-88, 246-254), lane 0
  int len = 0;
  if (lane == 0) {
    uint32_t p = goal_id;
    while (p != NODE_NONE) {
      ++len;
      p = nodes[H.lin_of(p)].parent;
    }
    if (len + 1 > C.path_cap) {
      err |= ERR_PATH_OVERFLOW;
      len = -1;
    } else {
      p = goal_id;
      int w = len - 1;
      while (p != NODE_NONE) {
        const int i = (int)(p >> 16), j = (int)((p >> 8) & 255u);
        const int k = (p & 255u) == AS_SPARE_KK ? sk : (int)(p & 255u) + k_lo;
        const D3 c = as_index2coord(C, F, i, j, k);
        W.path[3 * w] = c.x;
        W.path[3 * w + 1] = c.y;
        W.path[3 * w + 2] = c.z;
        --w;
        p = nodes[H.lin_of(p)].parent;
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
__device__ __noinline__ int shortcut_path(const DevMap& map, const VigoConst& C, const BatchView& bv, const double* path, int len,
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
__device__ __noinline__ bool find_guide_point(int cpIdx, int segFirst, int segSecond, const double* path, int plen, D3& guide) {
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
__device__ __noinline__ void append_pair(const DevMap& map, const VigoConst& C, const BatchView& bv, int b, TrajState& st, int c,
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
__device__ __noinline__ void assign_guides(const DevMap& map, const VigoConst& C, const BatchView& bv, int b, TrajState& st,
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
__device__ __noinline__ int find_collision_seg(const DevMap& map, const VigoConst& C, const BatchView& bv, const TrajState& st,
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
__device__ __noinline__ int path_search(const DevMap& map, const VigoConst& C, const BatchView& bv, TrajState& st, Worker& W,
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
    // this trajectory's next searches check early — unless the early checks keep finding the goal reachable
    if (W.goal_unreachable) W.flood_trigger = W.flood_wasted >= 2 ? TP_FLOOD_TRIGGER : TP_FLOOD_TRIGGER_AGAIN;
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
__device__ __noinline__ bool cp_requires_new_guide(const VigoConst& C, const BatchView& bv, int b, const TrajState& st, int c) {
  const D3 cp = ld3(bv.ctrl + 3 * (size_t)st.off, c);
  for (int gi = bv.cp_head[st.off + c]; gi >= 0;) {
    const GuidePair& pr = bv.pairs[(size_t)b * C.gcap + gi];
    const double dist = dot3(cp - d3(pr.p[0], pr.p[1], pr.p[2]), d3(pr.v[0], pr.v[1], pr.v[2]));
    if (C.p.dthresh - dist > 0) return false;
    gi = pr.next;
  }
  return true;
}
