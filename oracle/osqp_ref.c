/* ORACLE — TEST INFRASTRUCTURE ONLY.
 *
 * Thin C driver for the REFERENCE's own prebuilt QP solver, include/trajectory_planner/third_party/lib/x86/libosqp.so
 * (OSQP, C API of third_party/osqp/osqp.h; DLONG, double), compiled against the reference's headers where they lie
 * under /root/reference (oracle/Makefile target `osqp`; output oracle/_ref/libosqp_ref.so, git-ignored).  It restates
 * the call pattern of polyTrajSolver::setUpProblem / solveX (polyTrajSolver.cpp:162-223, 870-879) through OsqpEigen
 * 0.7.0: default settings (osqp_set_default_settings) with verbosity off, Hessian as its upper triangle in CSC,
 * constraint matrix in CSC, bounds l <= A x <= u, one solver per axis.
 * Used by tools/make_minsnap_golden.py to pin the numpy KKT oracle (oracle/frontend_np.py, oracle/polytraj_np.py) to
 * the reference's solver output; nothing in the product links or loads it. */
#include <stdlib.h>
#include <string.h>

#include "osqp.h"

/* Dense row-major P (n x n, symmetric; only its upper triangle is passed on) and A (m x n) -> CSC, then solve.
 * mode 0: the reference's settings (defaults, verbose off).  mode 1: the same problem driven to convergence
 *         (eps 1e-9, fixed rho schedule off the wall clock, polish) — the exact solution the defaults approximate.
 * Returns OSQP's status_val; x[n], y[m] = primal / dual solution, info[4] = {iterations, pri_res, dua_res, obj_val}. */
long long osqp_ref_solve(long long n, long long m, const double* P, const double* q, const double* A, const double* l,
                         const double* u, int mode, double* x, double* y, double* info) {
  c_int nnzP = 0, nnzA = 0;
  for (c_int j = 0; j < n; ++j)
    for (c_int i = 0; i <= j; ++i)
      if (P[i * n + j] != 0.0) ++nnzP;
  for (c_int j = 0; j < n; ++j)
    for (c_int i = 0; i < m; ++i)
      if (A[i * n + j] != 0.0) ++nnzA;
  c_float* Px = (c_float*)malloc(sizeof(c_float) * (size_t)(nnzP + 1));
  c_int* Pi = (c_int*)malloc(sizeof(c_int) * (size_t)(nnzP + 1));
  c_int* Pp = (c_int*)malloc(sizeof(c_int) * (size_t)(n + 1));
  c_float* Ax = (c_float*)malloc(sizeof(c_float) * (size_t)(nnzA + 1));
  c_int* Ai = (c_int*)malloc(sizeof(c_int) * (size_t)(nnzA + 1));
  c_int* Ap = (c_int*)malloc(sizeof(c_int) * (size_t)(n + 1));
  c_int k = 0;
  for (c_int j = 0; j < n; ++j) {
    Pp[j] = k;
    for (c_int i = 0; i <= j; ++i)
      if (P[i * n + j] != 0.0) { Px[k] = P[i * n + j]; Pi[k] = i; ++k; }
  }
  Pp[n] = k;
  k = 0;
  for (c_int j = 0; j < n; ++j) {
    Ap[j] = k;
    for (c_int i = 0; i < m; ++i)
      if (A[i * n + j] != 0.0) { Ax[k] = A[i * n + j]; Ai[k] = i; ++k; }
  }
  Ap[n] = k;

  OSQPSettings* settings = (OSQPSettings*)malloc(sizeof(OSQPSettings));
  OSQPData* data = (OSQPData*)malloc(sizeof(OSQPData));
  c_float* qq = (c_float*)malloc(sizeof(c_float) * (size_t)n);
  c_float* ll = (c_float*)malloc(sizeof(c_float) * (size_t)m);
  c_float* uu = (c_float*)malloc(sizeof(c_float) * (size_t)m);
  memcpy(qq, q, sizeof(c_float) * (size_t)n);
  memcpy(ll, l, sizeof(c_float) * (size_t)m);
  memcpy(uu, u, sizeof(c_float) * (size_t)m);
  data->n = n;
  data->m = m;
  data->P = csc_matrix(n, n, nnzP, Px, Pi, Pp);
  data->q = qq;
  data->A = csc_matrix(m, n, nnzA, Ax, Ai, Ap);
  data->l = ll;
  data->u = uu;
  osqp_set_default_settings(settings);
  settings->verbose = 0;
  if (mode == 1) {
    settings->eps_abs = 1e-9;
    settings->eps_rel = 1e-9;
    settings->max_iter = 200000;
    settings->polish = 1;
    settings->polish_refine_iter = 10;
    settings->adaptive_rho_interval = 50;   /* iteration-based: deterministic */
  }
  OSQPWorkspace* work = 0;
  c_int rc = osqp_setup(&work, data, settings);
  long long status = -100 - rc;
  if (rc == 0 && work) {
    osqp_solve(work);
    status = work->info->status_val;
    for (c_int i = 0; i < n; ++i) x[i] = work->solution->x[i];
    for (c_int i = 0; i < m; ++i) y[i] = work->solution->y[i];
    info[0] = (double)work->info->iter;
    info[1] = work->info->pri_res;
    info[2] = work->info->dua_res;
    info[3] = work->info->obj_val;
    osqp_cleanup(work);
  }
  free(data->P); free(data->A); free(data); free(settings);
  free(Px); free(Pi); free(Pp); free(Ax); free(Ai); free(Ap); free(qq); free(ll); free(uu);
  return status;
}
