// ORACLE — TEST INFRASTRUCTURE ONLY.  Nothing in the product path may include or link this file.
//
// CPU restatement of the WARP FORM of the fused ViGO cost + L-BFGS solve — the arithmetic the product's default
// (benchmarked) kernel runs: trajectory_planner_b200/csrc/tp_lbfgs_warp.cuh.  It emulates the 32 lanes of the warp
// of the team of NW warps (1 or 4; the product's default is 4) that owns a trajectory with plain loops: optimised
// element i belongs to thread i % (32 NW), a thread visits its elements in increasing order with private running sums,
// sums over a warp's lanes follow the kernel's shuffle trees (sum4 / butterfly), the warps' results are combined in
// warp order, the Gram dot products are serial per history row over each warp's contiguous chunk with four interleaved
// partial sums, and fused multiply-adds happen exactly where std::fma is written (build with -ffp-contract=off).  tests/ assert that the CUDA kernel and this file agree BIT FOR BIT on
// return code, iteration / evaluation counts and control points — that is the pin of the benchmarked mode.
// What is restated is every VALUE's sequence of operations, not the kernel's schedule or storage: the kernel keeps the
// history element-major and runs the two triangular recurrences four steps per round of shuffles (each row still gets
// its 16 FMAs in step order with the same operands), opens the next iteration inside its direction pass (xp = x, gp = g,
// x = xp + 1.0 d) — this file keeps row-major vectors and the step-by-step loops, and the results must still be equal.
//
// What the algorithm is (and how it relates to the reference): the same L-BFGS as solver/lbfgs.hpp:1024-1349 with
// m = 16 — identical More-Thuente line search (update_trial etc. from lbfgs_port.hpp), identical convergence and
// iteration tests — but (a) the cost terms of bsplineTraj.cpp:802-1064 are evaluated in gather form with fused
// multiply-adds and difference stencils, and (b) the two-loop recursion (lbfgs.hpp:1293-1316) is evaluated in
// coefficient space on Gram blocks.  Both change roundings, not mathematics; tests/test_oracle_cpu.py quantifies
// the effect against the reference-order oracle (per-evaluation relative difference, per-iteration direction
// difference, solve-level statistics on thousands of problems).
#pragma once
#include <cmath>
#include <cstring>
#include <vector>

#include "lbfgs_port.hpp"

namespace orc {
namespace wform {

static const int M = 16, GS = 18, PAIRS_SM = 64;

struct Problem {
  int N = 0;
  double* cp = nullptr;          // 3N control points, in/out (ends as the LAST EVALUATED point, bsplineTraj.cpp:803)
  std::vector<int> pstart;       // CSR by control point (N + 1), pairs in append order
  std::vector<double> pairs;     // 7 per pair: p xyz, v xyz, unknown flag (1.0 / 0.0)
  // parameters / constants (derived exactly as the product's make_const does)
  double ctrl_pt_ts = 0.2, ts = 0.1, dthresh = 0.5, dthresh_dyn = 0.5;
  double w_dist = 1, w_smooth = 1, w_feas = 1, w_dyn = 1, uncertain_factor = 1;
  double min_height = 0, max_height = 0;
  int plan_in_z = 0, pred_num = 0;
  double dist_a = 0, dist_b = 0, dist_c = 0, h_a = 0, h_b = 0, h_c = 0, dyn_a = 0, dyn_b = 0, dyn_c = 0, ts_inv_sqr = 0;
  std::vector<double> dyn_pos, dyn_vel, dyn_size;   // 3 per obstacle
  // solver parameters
  double g_eps = 0.01;
  int max_iter = 200, max_linesearch = 40;
};

struct Trace {   // optional per-iteration record (direction check in tests)
  std::vector<std::vector<double>> d, g, S, Y;   // direction, gradient, history (slot-major 16 x n) after each update
  std::vector<int> end, bound;
};

inline double cube_cr(double x) {   // tp_device.cuh cube_cr
  const double hi = x * x;
  const double lo = std::fma(x, x, -hi);
  const double h2 = hi * x;
  const double e2 = std::fma(hi, x, -h2);
  const double l2 = lo * x + e2;
  return h2 + l2;
}
inline double excess(double v) {   // wf_excess
  const double e = std::fabs(v) - 1.0;
  return e > 0.0 ? std::copysign(e, v) : 0.0;
}

// vf_pair_term (tp_lbfgs_fast.cuh)
inline void pair_term(const Problem& P, double cx, double cy, double cz, const double* pr, int a, double& grad, double& cost) {
  const double px = pr[0], py = pr[1], pz = pr[2], vx = pr[3], vy = pr[4], vz = pr[5];
  bool unk = pr[6] != 0.0;
  const double dth = P.dthresh;
  const double dist = std::fma(cx - px, vx, std::fma(cy - py, vy, (cz - pz) * vz));
  const double e = dth - dist;
  double va = a == 0 ? vx : (a == 1 ? vy : vz);
  if (!P.plan_in_z && a == 2) va = 0.0;
  double costTemp, gt;
  if (e <= -dth) {
    costTemp = -(e * e) * e;
    gt = 3.0 * (e * e);
    unk = false;
  } else if (e > 0 && e <= dth) {
    costTemp = (e * e) * e;
    gt = -3.0 * (e * e);
  } else if (e >= dth) {
    costTemp = std::fma(std::fma(P.dist_a, e, P.dist_b), e, P.dist_c);
    gt = -std::fma(2.0 * P.dist_a, e, P.dist_b);
  } else {
    return;
  }
  if (unk) { costTemp *= P.uncertain_factor; gt *= P.uncertain_factor; }
  grad = std::fma(gt, va, grad);
  cost += costTemp;
}
// vf_height_term (tp_lbfgs_fast.cuh)
inline void height_term(const Problem& P, double cz, int a, double& grad, double& cost) {
  const double hth = 0.2;
  const double hmin = cz - P.min_height, hmax = cz - P.max_height;
  const double ua = a == 0 ? 1.0 : 0.0;
  if (hmin < 0) {
    const double e = hth - hmin;
    cost += P.h_a * (e * e) + P.h_b * e + P.h_c;
    grad += (-(2 * P.h_a * e + P.h_b)) * -ua;
  } else if (hmin >= 0 && hmax < hth) {
    const double e = hth - hmin;
    cost += cube_cr(e);
    grad += (-3.0 * (e * e)) * -ua;
  }
  if (hmax > 0) {
    const double e = hth + hmax;
    cost += P.h_a * (e * e) + P.h_b * e + P.h_c;
    grad += (-(2 * P.h_a * e + P.h_b)) * ua;
  } else if (hmax <= 0 && hmax >= -hth) {
    const double e = hth + hmax;
    cost += cube_cr(e);
    grad += (-3.0 * (e * e)) * ua;
  }
}
// dynamic_terms (tp_lbfgs.cuh)
inline void dynamic_terms(const Problem& P, const double* cp, int c, int a, double& grad, double& cost) {
  const double dd = P.dthresh_dyn;
  const double cx = cp[3 * c], cy = cp[3 * c + 1];
  const int M_ = (int)(P.dyn_pos.size() / 3);
  for (int j = 0; j < M_; ++j) {
    const double hx = P.dyn_size[3 * j] / 2, hy = P.dyn_size[3 * j + 1] / 2;
    const double size = std::sqrt(hx * hx + hy * hy);
    for (int n = 0; n <= P.pred_num; n += 2) {
      const double tt = (double)n * P.ts;
      const double ox = P.dyn_pos[3 * j] + tt * P.dyn_vel[3 * j];
      const double oy = P.dyn_pos[3 * j + 1] + tt * P.dyn_vel[3 * j + 1];
      const double distThresh = (1 - (double)(P.pred_num != 0 ? n / P.pred_num : 0) * 0.2) * dd;
      const double dx = cx - ox, dy = cy - oy;
      const double dn = std::sqrt((dx * dx + dy * dy) + 0.0 * 0.0);
      const double dist = dn - size;
      const double distErr = distThresh - dist;
      const double ga = a == 0 ? dx / dn : (a == 1 ? dy / dn : 0.0 / dn);
      if (distErr <= 0) {
      } else if (distErr > 0 && distErr <= distThresh) {
        cost += cube_cr(distErr);
        grad += (-3.0 * (distErr * distErr)) * ga;
      } else if (distErr >= distThresh) {
        cost += (P.dyn_a * (distErr * distErr) + P.dyn_b * distErr + P.dyn_c);
        grad += (-(2 * P.dyn_a * distErr + P.dyn_b)) * ga;
      }
    }
  }
}

// wf_sum4: sums of four per-lane values, recursive halving on xor 16 / xor 8, butterfly 4, 2, 1 inside each group of
// eight lanes (group k holds value k), broadcast from lanes 0, 8, 16, 24
inline void sum4(const double (&v)[32][4], double (&out)[4]) {
  double a0[32], a1[32], b[32], w[32];
  for (int l = 0; l < 32; ++l) {
    const bool up = (l & 16) != 0;
    const int p = l ^ 16;
    a0[l] = (up ? v[l][2] : v[l][0]) + (up ? v[p][2] : v[p][0]);   // the partner sends what this lane keeps
    a1[l] = (up ? v[l][3] : v[l][1]) + (up ? v[p][3] : v[p][1]);
  }
  for (int l = 0; l < 32; ++l) {
    const bool up = (l & 8) != 0;
    const int p = l ^ 8;
    b[l] = (up ? a1[l] : a0[l]) + (up ? a1[p] : a0[p]);
  }
  for (int o = 4; o > 0; o >>= 1) {
    for (int l = 0; l < 32; ++l) w[l] = b[l] + b[l ^ o];
    std::memcpy(b, w, sizeof(w));
  }
  out[0] = b[0]; out[1] = b[8]; out[2] = b[16]; out[3] = b[24];
}

struct Solver {
  const Problem& P;
  int NW = 1;   // warps per team (1 or 4)
  int N, n, ns;
  double* cp;
  double* x;
  std::vector<double> g, xp, gp, d, S, Y, G, ca, cb;
  double sc[2] = {0, 0};
  double f_const = 0.0;
  int evals = 0;
  Trace* trace = nullptr;
  explicit Solver(const Problem& p, int nw = 1) : P(p), NW(nw) {
    N = P.N;
    n = 3 * (N - 6);
    const int ne = n > 0 ? (n + 3) & ~3 : 0;
    ns = ne + 1;
    cp = P.cp;
    x = cp + 9;
    g.assign(ne, 0.0); xp.assign(ne, 0.0); gp.assign(ne, 0.0); d.assign(ne, 0.0);
    S.assign((size_t)M * ns, 0.0); Y.assign((size_t)M * ns, 0.0);
    G.assign(3 * M * GS + 4 * M + 4, 0.0);
    ca.assign(M, 0.0); cb.assign(M, 0.0);
  }
  // A[s][c] = (s_s.y_c)/ys_s and Bt[c][s] = (s_s.y_c)/ys_c if pair s is OLDER than pair c, else 0; YY[s][c] = y_s.y_c
  double* A() { return G.data(); }
  double* Bt() { return G.data() + M * GS; }
  double* YY() { return G.data() + 2 * M * GS; }
  double* Sg() { return G.data() + 3 * M * GS; }
  double* Yg() { return G.data() + 3 * M * GS + M; }
  double* INV() { return G.data() + 3 * M * GS + 2 * M; }
  double* YS() { return G.data() + 3 * M * GS + 4 * M; }   // ys, yy, 1/yy of the newest pair

  // wf_const_terms: feasibility terms made of fixed control points only, times w_feas
  double const_terms() const {
    const double icts = 1.0 / P.ctrl_pt_ts, k2 = P.ts_inv_sqr;
    double t[3];
    for (int a = 0; a < 3; ++a) {
      double acc = 0.0;
      const int vi[4] = {0, 1, N - 3, N - 2};
      for (int q = 0; q < 4; ++q) {
        const int e = 3 * vi[q] + a;
        const double ev = excess((cp[e + 3] - cp[e]) * icts);
        acc = std::fma(ev * ev, k2, acc);
      }
      const int ai[2] = {0, N - 3};
      for (int q = 0; q < 2; ++q) {
        const int e = 3 * ai[q] + a;
        const double ea = excess((cp[e + 6] - 2 * cp[e + 3] + cp[e]) * k2);
        acc = std::fma(ea, ea, acc);
      }
      t[a] = acc;
    }
    return P.w_feas * ((t[0] + t[1]) + t[2]);
  }

  // wf_eval: writes g, returns f, g.d, g.g, x.x
  void eval(bool with_d, double& f, double& dgo, double& ggo, double& xxo) {
    const double icts = 1.0 / P.ctrl_pt_ts;
    const double k2 = P.ts_inv_sqr;
    const double gv_c = 2.0 * icts * k2;
    const double ga_c = 2.0 * k2;
    const int PT = 32 * NW;
    double v[4][32][4];
    for (int tid = 0; tid < PT; ++tid) {
      const int lane = tid & 31, warp = tid >> 5;
      double sD = 0, sS = 0, sF = 0, sO = 0, dg = 0, gg = 0, xx = 0;
      for (int i = tid; i < n; i += PT) {
        const int e = i + 9;
        const int c = e / 3, a = e - 3 * c;
        const double pm3 = cp[e - 9], pm2 = cp[e - 6], pm1 = cp[e - 3], p0 = cp[e], p1 = cp[e + 3], p2 = cp[e + 6], p3 = cp[e + 9];
        const double dm3 = pm2 - pm3, dm2 = pm1 - pm2, dm1 = p0 - pm1, d0 = p1 - p0, d1 = p2 - p1, d2 = p3 - p2;
        const double am3 = dm2 - dm3, am2 = dm1 - dm2, am1 = d0 - dm1, a0 = d1 - d0, a1 = d2 - d1;
        const double jm3 = am2 - am3, jm2 = am1 - am2, jm1 = a0 - am1, j0 = a1 - a0;
        const double gs = 2.0 * ((jm3 - j0) + 3.0 * (jm1 - jm2));
        const double evm = excess(dm1 * icts), ev0 = excess(d0 * icts);
        const double eam2 = excess(am2 * k2), eam1 = excess(am1 * k2), ea0 = excess(a0 * k2);
        const double gf = std::fma(gv_c, evm - ev0, ga_c * ((eam2 + ea0) - 2.0 * eam1));
        sS = std::fma(j0, j0, sS);
        sF = std::fma(ev0 * ev0, k2, sF);
        sF = std::fma(ea0, ea0, sF);
        if (c == 3) {
          sS = std::fma(jm3, jm3, sS);
          sS = std::fma(jm2, jm2, sS);
          sS = std::fma(jm1, jm1, sS);
          sF = std::fma(evm * evm, k2, sF);
          sF = std::fma(eam2, eam2, sF);
          sF = std::fma(eam1, eam1, sF);
        }
        double gd = 0.0, cD = 0.0, go = 0.0, cO = 0.0;
        const double cx = cp[3 * c], cy = cp[3 * c + 1], cz = cp[3 * c + 2];
        const bool in_sm = P.pstart[N] <= PAIRS_SM;   // the kernel stages up to 64 pairs in shared memory
        if (in_sm)
          for (int q = P.pstart[c]; q < P.pstart[c + 1]; ++q) pair_term(P, cx, cy, cz, &P.pairs[7 * (size_t)q], a, gd, cD);
        if (!in_sm || P.plan_in_z || !P.dyn_pos.empty()) {
          // wf_rare_terms: accumulated from zero, then ADDED
          double r0 = 0.0, r1 = 0.0, r2 = 0.0, r3 = 0.0;
          if (!in_sm)
            for (int q = P.pstart[c]; q < P.pstart[c + 1]; ++q) pair_term(P, cx, cy, cz, &P.pairs[7 * (size_t)q], a, r0, r1);
          if (P.plan_in_z) height_term(P, cz, a, r0, r1);
          if (!P.dyn_pos.empty()) dynamic_terms(P, cp, c, a, r2, r3);
          gd += r0; cD += r1; go += r2; cO += r3;
        }
        if (a == 0) { sD += cD; sO += cO; }
        const double gv = std::fma(P.w_dist, gd, std::fma(P.w_smooth, gs, std::fma(P.w_feas, gf, P.w_dyn * go)));
        g[i] = gv;
        gg = std::fma(gv, gv, gg);
        xx = std::fma(p0, p0, xx);
        if (with_d) dg = std::fma(gv, d[i], dg);
      }
      v[warp][lane][0] = P.w_dist * sD + P.w_smooth * sS + P.w_feas * sF + P.w_dyn * sO;
      v[warp][lane][1] = dg;
      v[warp][lane][2] = gg;
      v[warp][lane][3] = xx;
    }
    double out[4];
    sum4(v[0], out);
    for (int w = 1; w < NW; ++w) {   // the warps' results, combined in warp order
      double ow[4];
      sum4(v[w], ow);
      for (int k = 0; k < 4; ++k) out[k] = out[k] + ow[k];
    }
    f = out[0] + f_const;
    dgo = out[1];
    ggo = out[2];
    xxo = out[3];
    ++evals;
  }

  // wf_gram_update: new pair into `slot` (y also into gp), then lane l = history row l (s_0..15, y_0..15) runs its two
  // dot products with y_new and g serially over the zero-padded elements, four interleaved partial sums each
  void gram_update(int slot) {
    for (int i = 0; i < n; ++i) {
      const double yi = g[i] - gp[i];
      S[(size_t)slot * ns + i] = x[i] - xp[i];
      Y[(size_t)slot * ns + i] = yi;
      gp[i] = yi;
    }
    const int n4 = (n + 3) & ~3;
    const int chunk = ((n4 / 4 + NW - 1) / NW) * 4;   // each warp covers a contiguous chunk of the elements
    double dy[32], dgv[32];
    for (int l = 0; l < 32; ++l) {
      const double* row = l < M ? &S[(size_t)l * ns] : &Y[(size_t)(l - M) * ns];
      for (int w = 0; w < NW; ++w) {
        const int i0 = w * chunk, i1 = i0 + chunk < n4 ? i0 + chunk : n4;
        double a[4] = {0, 0, 0, 0}, b[4] = {0, 0, 0, 0};
        for (int i = i0; i < i1; i += 4)
          for (int k = 0; k < 4; ++k) {
            a[k] = std::fma(row[i + k], gp[i + k], a[k]);
            b[k] = std::fma(row[i + k], g[i + k], b[k]);
          }
        const double py = (a[0] + a[1]) + (a[2] + a[3]), pg = (b[0] + b[1]) + (b[2] + b[3]);
        dy[l] = w == 0 ? py : dy[l] + py;      // combined in warp order
        dgv[l] = w == 0 ? pg : dgv[l] + pg;
      }
    }
    const double inv_new = 1.0 / dy[slot];         // 1 / ys_new
    const double inv_yy = 1.0 / dy[M + slot];      // 1 / yy_new
    for (int l = 0; l < M; ++l) {
      const bool self = l == slot;
      const double inv_l = self ? inv_new : INV()[l];
      A()[slot * GS + l] = 0.0;
      Bt()[l * GS + slot] = 0.0;
      A()[l * GS + slot] = self ? 0.0 : dy[l] * inv_l;
      Bt()[slot * GS + l] = self ? 0.0 : dy[l] * inv_new;
      Sg()[l] = dgv[l];
    }
    INV()[slot] = inv_new;
    YS()[0] = dy[slot];
    for (int j = 0; j < M; ++j) {
      YY()[j * GS + slot] = dy[M + j];
      YY()[slot * GS + j] = dy[M + j];
      Yg()[j] = dgv[M + j];
    }
    YS()[1] = dy[M + slot];
    YS()[2] = inv_yy;
  }

  // wf_coeffs: lane s <-> slot s (lanes 16..31 compute the same values redundantly on the device)
  void coeffs(int newest, double gg) {
    const double gamma = YS()[0] * YS()[2];
    double rr[M], al[M], bacc[M];
    for (int s_ = 0; s_ < M; ++s_) rr[s_] = -Sg()[s_] * INV()[s_];
    for (int t = 0; t < M; ++t) {
      const int st = (newest - t) & (M - 1);
      const double ala = rr[st];   // shuffle from lane st (its value BEFORE this step's update)
      for (int s_ = 0; s_ < M; ++s_) rr[s_] = std::fma(-ala, A()[s_ * GS + st], rr[s_]);
    }
    for (int s_ = 0; s_ < M; ++s_) al[s_] = rr[s_];
    for (int s_ = 0; s_ < M; ++s_) {
      double t0 = Yg()[s_], t1 = 0.0, t2 = 0.0, t3 = 0.0;
      for (int c = 0; c < M; c += 4) {
        t0 = std::fma(al[c], YY()[s_ * GS + c], t0);
        t1 = std::fma(al[c + 1], YY()[s_ * GS + c + 1], t1);
        t2 = std::fma(al[c + 2], YY()[s_ * GS + c + 2], t2);
        t3 = std::fma(al[c + 3], YY()[s_ * GS + c + 3], t3);
      }
      bacc[s_] = (-gamma * ((t0 + t1) + (t2 + t3))) * INV()[s_];
    }
    for (int t = M - 1; t >= 0; --t) {
      const int st = (newest - t) & (M - 1);
      const double cst = al[st] - bacc[st];   // shuffle from lane st (value BEFORE this step's update)
      for (int s_ = 0; s_ < M; ++s_) bacc[s_] = std::fma(cst, Bt()[s_ * GS + st], bacc[s_]);
    }
    double part[M];
    for (int s_ = 0; s_ < M; ++s_) {
      const double aa = al[s_] - bacc[s_];
      const double bb = -gamma * al[s_];
      ca[s_] = aa;
      cb[s_] = bb;
      part[s_] = std::fma(aa, Sg()[s_], bb * Yg()[s_]);
    }
    for (int o = 8; o > 0; o >>= 1) {
      double w[M];
      for (int i = 0; i < M; ++i) w[i] = part[i] + part[i ^ o];
      std::memcpy(part, w, sizeof(w));
    }
    sc[0] = -gamma;
    sc[1] = part[0] - gamma * gg;
  }

  // wf_coeffs AS THE KERNEL SCHEDULES IT (tp_lbfgs_warp.cuh): four steps per round of shuffles — the four rows whose
  // values become final in steps 4k..4k+3 are broadcast at once, every lane redoes their in-block updates, then each
  // lane applies the four steps to its own row.  Must equal coeffs() bit for bit (tests/test_oracle_cpu.py).
  void coeffs_blocked(int newest, double gg) {
    const double gamma = YS()[0] * YS()[2];
    double rr[M], al[M], bacc[M];
    for (int s_ = 0; s_ < M; ++s_) rr[s_] = -Sg()[s_] * INV()[s_];
    for (int k = 0; k < M / 4; ++k) {
      int sl[4];
      for (int j = 0; j < 4; ++j) sl[j] = (newest - 4 * k - j) & (M - 1);
      double a[4];
      for (int j = 0; j < 4; ++j) a[j] = rr[sl[j]];                                  // the round of shuffles
      for (int i = 0; i < 4; ++i)                                                      // in-block updates, every lane
        for (int j = i + 1; j < 4; ++j) a[j] = std::fma(-a[i], A()[sl[j] * GS + sl[i]], a[j]);
      for (int s_ = 0; s_ < M; ++s_)                                                   // own row, four steps in order
        for (int i = 0; i < 4; ++i) rr[s_] = std::fma(-a[i], A()[s_ * GS + sl[i]], rr[s_]);
    }
    for (int s_ = 0; s_ < M; ++s_) al[s_] = rr[s_];
    for (int s_ = 0; s_ < M; ++s_) {
      double t0 = Yg()[s_], t1 = 0.0, t2 = 0.0, t3 = 0.0;
      for (int c = 0; c < M; c += 4) {
        t0 = std::fma(al[c], YY()[s_ * GS + c], t0);
        t1 = std::fma(al[c + 1], YY()[s_ * GS + c + 1], t1);
        t2 = std::fma(al[c + 2], YY()[s_ * GS + c + 2], t2);
        t3 = std::fma(al[c + 3], YY()[s_ * GS + c + 3], t3);
      }
      bacc[s_] = (-gamma * ((t0 + t1) + (t2 + t3))) * INV()[s_];
    }
    for (int k = 0; k < M / 4; ++k) {
      int sl[4];
      for (int j = 0; j < 4; ++j) sl[j] = (newest - (M - 1) + 4 * k + j) & (M - 1);
      double b[4], c[4];
      for (int j = 0; j < 4; ++j) b[j] = bacc[sl[j]];
      for (int i = 0; i < 4; ++i) {
        c[i] = al[sl[i]] - b[i];
        for (int j = i + 1; j < 4; ++j) b[j] = std::fma(c[i], Bt()[sl[j] * GS + sl[i]], b[j]);
      }
      for (int s_ = 0; s_ < M; ++s_)
        for (int i = 0; i < 4; ++i) bacc[s_] = std::fma(c[i], Bt()[s_ * GS + sl[i]], bacc[s_]);
    }
    double part[M];
    for (int s_ = 0; s_ < M; ++s_) {
      const double aa = al[s_] - bacc[s_];
      const double bb = -gamma * al[s_];
      ca[s_] = aa;
      cb[s_] = bb;
      part[s_] = std::fma(aa, Sg()[s_], bb * Yg()[s_]);
    }
    for (int o = 8; o > 0; o >>= 1) {
      double w[M];
      for (int i = 0; i < M; ++i) w[i] = part[i] + part[i ^ o];
      std::memcpy(part, w, sizeof(w));
    }
    sc[0] = -gamma;
    sc[1] = part[0] - gamma * gg;
  }

  // wf_direction
  void direction() {
    const double cg = sc[0];
    for (int i = 0; i < n; ++i) {
      double a0 = cg * g[i], a1 = 0.0, a2 = 0.0, a3 = 0.0;
      for (int j = 0; j < M / 2; ++j) {
        a0 = std::fma(ca[j], S[(size_t)j * ns + i], a0);
        a1 = std::fma(ca[j + 8], S[(size_t)(j + 8) * ns + i], a1);
        a2 = std::fma(cb[j], Y[(size_t)j * ns + i], a2);
        a3 = std::fma(cb[j + 8], Y[(size_t)(j + 8) * ns + i], a3);
      }
      d[i] = (a0 + a1) + (a2 + a3);
    }
  }

  // lbfgs_run_warp.  xfinal (n, may be null) receives the solver's own x.
  LbfgsStats run(double* xfinal) {
    using namespace detail;
    LbfgsStats st;
    if (n <= 0) { st.ret = LBERR_INVALID_N; return st; }
    const double min_step = 1e-20, max_step = 1e20, ftol = 1e-4, gtol = 0.9, xtol = 1e-16;
    const int max_ls = P.max_linesearch;
    const double geps2 = P.g_eps * P.g_eps;
    int k = 0, ret;
    double fx, dgd, gg, xx;
    f_const = const_terms();
    eval(false, fx, dgd, gg, xx);
    for (int i = 0; i < n; ++i) d[i] = -g[i];
    double xnorm = std::sqrt(xx), gnorm = std::sqrt(gg);
    if (xnorm < 1.0) xnorm = 1.0;
    bool reverted = false;
    if (gnorm / xnorm <= P.g_eps) {
      ret = LB_ALREADY_MINIMIZED;
    } else {
      double step = 1.0 / std::sqrt(gg);
      double dginit_next = -gg;
      int end = 0;
      k = 1;
      for (;;) {
        for (int i = 0; i < n; ++i) { xp[i] = x[i]; gp[i] = g[i]; }
        int ls;
        {
          int count = 0, brackt = 0, stage1 = 1, uinfo = 0;
          double dg, stx, fxx, dgx, sty, fy, dgy, finit, ftest1, dginit, dgtest, width, prev_width, stmin = 0, stmax = 0;
          double stp = step;
          if (stp <= 0.) {
            ls = LBERR_INVALIDPARAMS;
          } else {
            dginit = dginit_next;
            if (0 < dginit) {
              ls = LBERR_INCREASEGRADIENT;
            } else {
              finit = fx;
              dgtest = ftol * dginit;
              width = max_step - min_step;
              prev_width = 2.0 * width;
              stx = sty = 0.;
              fxx = fy = finit;
              dgx = dgy = dginit;
              for (;;) {
                if (brackt) {
                  stmin = stx <= sty ? stx : sty;
                  stmax = stx >= sty ? stx : sty;
                } else {
                  stmin = stx;
                  stmax = stp + 4.0 * (stp - stx);
                }
                if (stp < min_step) stp = min_step;
                if (max_step < stp) stp = max_step;
                if ((brackt && ((stp <= stmin || stmax <= stp) || max_ls <= count + 1 || uinfo != 0)) ||
                    (brackt && (stmax - stmin <= xtol * stmax)))
                  stp = stx;
                for (int i = 0; i < n; ++i) x[i] = xp[i] + stp * d[i];
                eval(true, fx, dg, gg, xx);
                ftest1 = finit + stp * dgtest;
                ++count;
                if (brackt && ((stp <= stmin || stmax <= stp) || uinfo != 0)) { ls = LBERR_ROUNDING; break; }
                if (stp == max_step && fx <= ftest1 && dg <= dgtest) { ls = LBERR_MAXSTEP; break; }
                if (stp == min_step && (ftest1 < fx || dgtest <= dg)) { ls = LBERR_MINSTEP; break; }
                if (brackt && (stmax - stmin) <= xtol * stmax) { ls = LBERR_WIDTHTOOSMALL; break; }
                if (max_ls <= count) { ls = LBERR_MAXLINESEARCH; break; }
                if (fx <= ftest1 && std::fabs(dg) <= gtol * (-dginit)) { ls = count; break; }
                if (stage1 && fx <= ftest1 && (ftol <= gtol ? ftol : gtol) * dginit <= dg) stage1 = 0;
                if (stage1 && ftest1 < fx && fx <= fxx) {
                  double fm = fx - stp * dgtest;
                  double fxm = fxx - stx * dgtest;
                  double fym = fy - sty * dgtest;
                  double dgm = dg - dgtest;
                  double dgxm = dgx - dgtest;
                  double dgym = dgy - dgtest;
                  uinfo = update_trial(stx, fxm, dgxm, sty, fym, dgym, stp, fm, dgm, stmin, stmax, brackt);
                  fxx = fxm + stx * dgtest;
                  fy = fym + sty * dgtest;
                  dgx = dgxm + dgtest;
                  dgy = dgym + dgtest;
                } else {
                  double ft = fx, dt = dg;
                  uinfo = update_trial(stx, fxx, dgx, sty, fy, dgy, stp, ft, dt, stmin, stmax, brackt);
                }
                if (brackt) {
                  if (0.66 * prev_width <= std::fabs(sty - stx)) stp = stx + 0.5 * (sty - stx);
                  prev_width = width;
                  width = std::fabs(sty - stx);
                }
              }
            }
          }
          step = stp;
        }
        if (ls < 0) {
          if (xfinal) std::memcpy(xfinal, xp.data(), sizeof(double) * n);
          reverted = true;
          ret = ls;
          break;
        }
        if (gg <= geps2 * (xx < 1.0 ? 1.0 : xx)) { ret = LB_CONVERGENCE; break; }
        if (P.max_iter != 0 && P.max_iter < k + 1) { ret = LBERR_MAXITER; break; }
        const int bound = (M <= k) ? M : k;
        gram_update(end);
        coeffs(end, gg);
        dginit_next = sc[1];
        direction();
        if (trace) {
          trace->d.emplace_back(d.begin(), d.begin() + n);
          trace->g.emplace_back(g.begin(), g.begin() + n);
          std::vector<double> s_((size_t)M * n), y_((size_t)M * n);
          for (int j = 0; j < M; ++j)
            for (int i = 0; i < n; ++i) { s_[(size_t)j * n + i] = S[(size_t)j * ns + i]; y_[(size_t)j * n + i] = Y[(size_t)j * ns + i]; }
          trace->S.push_back(std::move(s_));
          trace->Y.push_back(std::move(y_));
          trace->end.push_back(end);
          trace->bound.push_back(bound);
        }
        ++k;
        end = (end + 1) & (M - 1);
        step = 1.0;
      }
    }
    if (xfinal && !reverted) std::memcpy(xfinal, x, sizeof(double) * n);
    st.ret = ret;
    st.iters = k;
    st.evals = evals;
    st.fx = fx;
    return st;
  }
};

}  // namespace wform
}  // namespace orc
