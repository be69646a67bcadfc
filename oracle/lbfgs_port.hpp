// ORACLE — TEST INFRASTRUCTURE ONLY.  Nothing in the product path may include or link this file.
//
// CPU restatement of the L-BFGS + More-Thuente line search that the reference's ViGO solve runs
// (reference: include/trajectory_planner/solver/lbfgs.hpp, lbfgs_optimize :1024-1349,
// line_search_morethuente :716-937, update_trial_interval :506-714, minimiser macros :308-391,
// vec* helpers :408-473, defaults :942-954).
//
// It is a restatement, not a copy: same arithmetic in the same order (so that results are
// bit-identical to the reference header when both are compiled with -ffp-contract=off), but
// written as one self-contained class without the callback/progress plumbing the reference's
// ViGO path never uses (proc_stepbound = proc_progress = NULL, past = 0; bsplineTraj.cpp:701).
// `oracle/_ref` builds the same oracle against the reference header itself
// (-DTP_ORACLE_REF_LBFGS) and tests/test_oracle_cpu.py::test_lbfgs_port_is_pinned_to_reference_header pins this port to it bit-for-bit.
#pragma once
#include <cmath>
#include <cstring>
#include <vector>

namespace orc {

// status codes: same numeric values as the reference enum (lbfgs.hpp:20-80)
enum {
  LB_CONVERGENCE = 0,
  LB_STOP = 1,
  LB_ALREADY_MINIMIZED = 2,
  LBERR_UNKNOWN = -1024,
  LBERR_LOGIC = -1023,
  LBERR_CANCELED = -1022,
  LBERR_INVALID_N = -1021,
  LBERR_OUTOFINTERVAL = -1010,
  LBERR_INCORRECT_TMINMAX = -1009,
  LBERR_ROUNDING = -1008,
  LBERR_MINSTEP = -1007,
  LBERR_MAXSTEP = -1006,
  LBERR_MAXLINESEARCH = -1005,
  LBERR_MAXITER = -1004,
  LBERR_WIDTHTOOSMALL = -1003,
  LBERR_INVALIDPARAMS = -1002,
  LBERR_INCREASEGRADIENT = -1001,
};

struct LbfgsParams {
  int m = 16;                 // bsplineTraj.cpp:697
  double g_eps = 0.01;        // :699
  int max_iter = 200;         // :698
  int max_linesearch = 40;    // lbfgs.hpp:948
  double min_step = 1e-20, max_step = 1e20;
  double ftol = 1e-4, gtol = 0.9, xtol = 1e-16;
};

struct LbfgsStats {
  int ret = 0, iters = 0, evals = 0;
  double fx = 0;
};

namespace detail {
// serial dot product, lbfgs.hpp:453-461
inline double dot(const double* x, const double* y, int n) {
  double s = 0.;
  for (int i = 0; i < n; ++i) s += x[i] * y[i];
  return s;
}
// cubic minimiser, lbfgs.hpp:308-324
inline double cubic_min(double u, double fu, double du, double v, double fv, double dv) {
  double d = v - u;
  double theta = (fu - fv) * 3 / d + du + dv;
  double p = std::fabs(theta), q = std::fabs(du), r = std::fabs(dv);
  double s = p >= q ? p : q;
  s = s >= r ? s : r;
  double a = theta / s;
  double gamm = s * std::sqrt(a * a - (du / s) * (dv / s));
  if (v < u) gamm = -gamm;
  p = gamm - du + theta;
  q = gamm - du + gamm + dv;
  r = p / q;
  return u + r * d;
}
// bounded cubic minimiser, lbfgs.hpp:338-366
inline double cubic_min2(double u, double fu, double du, double v, double fv, double dv,
                         double xmin, double xmax) {
  double d = v - u;
  double theta = (fu - fv) * 3 / d + du + dv;
  double p = std::fabs(theta), q = std::fabs(du), r = std::fabs(dv);
  double s = p >= q ? p : q;
  s = s >= r ? s : r;
  double a = theta / s;
  double gamm = a * a - (du / s) * (dv / s);
  gamm = gamm > 0 ? s * std::sqrt(gamm) : 0;
  if (u < v) gamm = -gamm;
  p = gamm - dv + theta;
  q = gamm - dv + gamm + du;
  r = p / q;
  if (r < 0. && gamm != 0.) return v - r * d;
  if (a < 0) return xmax;
  return xmin;
}
// quadratic minimisers, lbfgs.hpp:377-391
inline double quad_min(double u, double fu, double du, double v, double fv) {
  double a = v - u;
  return u + du / ((fu - fv) / a + du) / 2 * a;
}
inline double quad_min2(double u, double du, double v, double dv) {
  double a = u - v;
  return v + dv / (dv - du) * a;
}

// safeguarded trial-value update, lbfgs.hpp:506-714
inline int update_trial(double& x, double& fx, double& dx, double& y, double& fy, double& dy,
                        double& t, double& ft, double& dt, double tmin, double tmax, int& brackt) {
  int bound;
  int dsign = dt * (dx / std::fabs(dx)) < 0.;
  double mc, mq, newt;
  if (brackt) {
    if (t <= (x <= y ? x : y) || (x >= y ? x : y) <= t) return LBERR_OUTOFINTERVAL;
    if (0. <= dx * (t - x)) return LBERR_INCREASEGRADIENT;
    if (tmax < tmin) return LBERR_INCORRECT_TMINMAX;
  }
  if (fx < ft) {
    brackt = 1;
    bound = 1;
    mc = cubic_min(x, fx, dx, t, ft, dt);
    mq = quad_min(x, fx, dx, t, ft);
    newt = (std::fabs(mc - x) < std::fabs(mq - x)) ? mc : mc + 0.5 * (mq - mc);
  } else if (dsign) {
    brackt = 1;
    bound = 0;
    mc = cubic_min(x, fx, dx, t, ft, dt);
    mq = quad_min2(x, dx, t, dt);
    newt = (std::fabs(mc - t) > std::fabs(mq - t)) ? mc : mq;
  } else if (std::fabs(dt) < std::fabs(dx)) {
    bound = 1;
    mc = cubic_min2(x, fx, dx, t, ft, dt, tmin, tmax);
    mq = quad_min2(x, dx, t, dt);
    if (brackt)
      newt = (std::fabs(t - mc) < std::fabs(t - mq)) ? mc : mq;
    else
      newt = (std::fabs(t - mc) > std::fabs(t - mq)) ? mc : mq;
  } else {
    bound = 0;
    if (brackt)
      newt = cubic_min(t, ft, dt, y, fy, dy);
    else if (x < t)
      newt = tmax;
    else
      newt = tmin;
  }
  if (fx < ft) {
    y = t; fy = ft; dy = dt;
  } else {
    if (dsign) { y = x; fy = fx; dy = dx; }
    x = t; fx = ft; dx = dt;
  }
  if (tmax < newt) newt = tmax;
  if (newt < tmin) newt = tmin;
  if (brackt && bound) {
    mq = x + 0.66 * (y - x);
    if (x < y) { if (mq < newt) newt = mq; }
    else       { if (newt < mq) newt = mq; }
  }
  t = newt;
  return 0;
}
}  // namespace detail

// Eval: double operator()(const double* x, double* g, int n)
template <class Eval>
struct Lbfgs {
  LbfgsParams P;
  Eval& eval;
  int evals = 0;
  Lbfgs(Eval& e, const LbfgsParams& p) : P(p), eval(e) {}

  // More-Thuente, lbfgs.hpp:716-937
  int line_search(int n, double* x, double* f, double* g, const double* s, double* stp,
                  const double* xp) {
    using namespace detail;
    int count = 0, brackt = 0, stage1 = 1, uinfo = 0;
    double dg, stx, fx, dgx, sty, fy, dgy, fxm, dgxm, fym, dgym, fm, dgm;
    double finit, ftest1, dginit, dgtest, width, prev_width, stmin, stmax;
    if (*stp <= 0.) return LBERR_INVALIDPARAMS;
    dginit = dot(g, s, n);
    if (0 < dginit) return LBERR_INCREASEGRADIENT;
    finit = *f;
    dgtest = P.ftol * dginit;
    width = P.max_step - P.min_step;
    prev_width = 2.0 * width;
    stx = sty = 0.;
    fx = fy = finit;
    dgx = dgy = dginit;
    for (;;) {
      if (brackt) {
        stmin = stx <= sty ? stx : sty;
        stmax = stx >= sty ? stx : sty;
      } else {
        stmin = stx;
        stmax = *stp + 4.0 * (*stp - stx);
      }
      if (*stp < P.min_step) *stp = P.min_step;
      if (P.max_step < *stp) *stp = P.max_step;
      if ((brackt && ((*stp <= stmin || stmax <= *stp) || P.max_linesearch <= count + 1 || uinfo != 0)) ||
          (brackt && (stmax - stmin <= P.xtol * stmax)))
        *stp = stx;
      std::memcpy(x, xp, sizeof(double) * n);
      for (int i = 0; i < n; ++i) x[i] += *stp * s[i];
      *f = eval(x, g, n);
      ++evals;
      dg = dot(g, s, n);
      ftest1 = finit + *stp * dgtest;
      ++count;
      if (brackt && ((*stp <= stmin || stmax <= *stp) || uinfo != 0)) return LBERR_ROUNDING;
      if (*stp == P.max_step && *f <= ftest1 && dg <= dgtest) return LBERR_MAXSTEP;
      if (*stp == P.min_step && (ftest1 < *f || dgtest <= dg)) return LBERR_MINSTEP;
      if (brackt && (stmax - stmin) <= P.xtol * stmax) return LBERR_WIDTHTOOSMALL;
      if (P.max_linesearch <= count) return LBERR_MAXLINESEARCH;
      if (*f <= ftest1 && std::fabs(dg) <= P.gtol * (-dginit)) return count;
      if (stage1 && *f <= ftest1 && (P.ftol <= P.gtol ? P.ftol : P.gtol) * dginit <= dg) stage1 = 0;
      if (stage1 && ftest1 < *f && *f <= fx) {
        fm = *f - *stp * dgtest;
        fxm = fx - stx * dgtest;
        fym = fy - sty * dgtest;
        dgm = dg - dgtest;
        dgxm = dgx - dgtest;
        dgym = dgy - dgtest;
        uinfo = update_trial(stx, fxm, dgxm, sty, fym, dgym, *stp, fm, dgm, stmin, stmax, brackt);
        fx = fxm + stx * dgtest;
        fy = fym + sty * dgtest;
        dgx = dgxm + dgtest;
        dgy = dgym + dgtest;
      } else {
        uinfo = update_trial(stx, fx, dgx, sty, fy, dgy, *stp, *f, dg, stmin, stmax, brackt);
      }
      if (brackt) {
        if (0.66 * prev_width <= std::fabs(sty - stx)) *stp = stx + 0.5 * (sty - stx);
        prev_width = width;
        width = std::fabs(sty - stx);
      }
    }
  }

  // driver, lbfgs.hpp:1024-1349
  LbfgsStats optimize(int n, double* x) {
    using namespace detail;
    LbfgsStats st;
    if (n <= 0) { st.ret = LBERR_INVALID_N; return st; }
    const int m = P.m;
    std::vector<double> xp(n, 0.), g(n, 0.), gp(n, 0.), d(n, 0.);
    std::vector<double> S((size_t)m * n, 0.), Y((size_t)m * n, 0.), alpha(m, 0.), YS(m, 0.);
    double fx = eval(x, g.data(), n);
    ++evals;
    for (int i = 0; i < n; ++i) d[i] = -g[i];
    double xnorm = std::sqrt(dot(x, x, n)), gnorm = std::sqrt(dot(g.data(), g.data(), n));
    int ret, k = 0;
    if (xnorm < 1.0) xnorm = 1.0;
    if (gnorm / xnorm <= P.g_eps) {
      ret = LB_ALREADY_MINIMIZED;
    } else {
      double step = 1.0 / std::sqrt(dot(d.data(), d.data(), n));
      int end = 0;
      k = 1;
      for (;;) {
        std::memcpy(xp.data(), x, sizeof(double) * n);
        std::memcpy(gp.data(), g.data(), sizeof(double) * n);
        int ls = line_search(n, x, &fx, g.data(), d.data(), &step, xp.data());
        if (ls < 0) {
          std::memcpy(x, xp.data(), sizeof(double) * n);
          std::memcpy(g.data(), gp.data(), sizeof(double) * n);
          ret = ls;
          break;
        }
        xnorm = std::sqrt(dot(x, x, n));
        gnorm = std::sqrt(dot(g.data(), g.data(), n));
        if (xnorm < 1.0) xnorm = 1.0;
        if (gnorm / xnorm <= P.g_eps) { ret = LB_CONVERGENCE; break; }
        if (P.max_iter != 0 && P.max_iter < k + 1) { ret = LBERR_MAXITER; break; }
        double* s = &S[(size_t)end * n];
        double* y = &Y[(size_t)end * n];
        for (int i = 0; i < n; ++i) s[i] = x[i] - xp[i];
        for (int i = 0; i < n; ++i) y[i] = g[i] - gp[i];
        double ys = dot(y, s, n), yy = dot(y, y, n);
        YS[end] = ys;
        int bound = (m <= k) ? m : k;
        ++k;
        end = (end + 1) % m;
        for (int i = 0; i < n; ++i) d[i] = -g[i];
        int j = end;
        for (int i = 0; i < bound; ++i) {
          j = (j + m - 1) % m;
          alpha[j] = dot(&S[(size_t)j * n], d.data(), n);
          alpha[j] /= YS[j];
          const double c = -alpha[j];
          const double* yj = &Y[(size_t)j * n];
          for (int e = 0; e < n; ++e) d[e] += c * yj[e];
        }
        const double sc = ys / yy;
        for (int e = 0; e < n; ++e) d[e] *= sc;
        for (int i = 0; i < bound; ++i) {
          double beta = dot(&Y[(size_t)j * n], d.data(), n);
          beta /= YS[j];
          const double c = alpha[j] - beta;
          const double* sj = &S[(size_t)j * n];
          for (int e = 0; e < n; ++e) d[e] += c * sj[e];
          j = (j + 1) % m;
        }
        step = 1.0;
      }
    }
    st.ret = ret;
    st.iters = k;
    st.evals = evals;
    st.fx = fx;
    return st;
  }
};

}  // namespace orc
