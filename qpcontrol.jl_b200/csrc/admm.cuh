// admm.cuh -- OSQP-style ADMM for one dense QP per CTA, everything resident in shared memory.
//
//   min 1/2 x'Px + q'x   s.t.  lg <= G x <= ug (mg general rows),  lb <= x[n-nbx..n) <= ub (nbx box rows)
//
// Device-side replacement for `MOI.optimize!(::OSQP.Optimizer)` reached by `solve!(qpmodel)` (reference
// src/lowlevel/momentum.jl:58; solver plugged in at momentum.jl:1,15-16,27; settings test/runtests.jl:35-43,
// notebooks/Standing controller.ipynb:66-71).  Same algorithm as OSQP 0.5.x on A = [G; E_box]:
// Ruiz equilibration, per-row rho (equalities x1e3), relaxed ADMM with alpha, unscaled termination residuals every
// `check_termination` iterations, primal/dual infeasibility certificates, adaptive rho with refactorisation.
// B200-first differences: the quasi-definite KKT solve is replaced by its Schur complement
// S = P + sigma I + A' rho A (dense, n x n), Cholesky-factorised in shared memory; the factor is inverted once so the
// per-iteration solve is two triangular mat-vecs (L^-1 then L^-T) with no dependent chain; box rows are handled as a
// diagonal instead of rows of A.
#pragma once
#include "qpc_common.h"
#include "qpc_program.h"

namespace qpc {

#define QPC_INFTY 1e20
#define QPC_RHO_MIN 1e-6
#define QPC_RHO_MAX 1e6
#define QPC_RHO_EQ_FACTOR 1e3
#define QPC_RHO_TOL 1e-4
#define QPC_MIN_SCALING 1e-4
#define QPC_MAX_SCALING 1e4

#if defined(QPC_THREAD_PER_INSTANCE)
#define QPC_RED_DOUBLES 0      // no cross-thread reductions: one thread owns the QP
#else
#define QPC_RED_DOUBLES (32 * 16)
#endif
struct AdmmSmem {
  double *M;        // n x n   : P_bar -> S -> L -> L^-1 (lower) mirrored into the upper triangle
  double *Gs, *Gt;  // mg x n scaled G, n x mg its transpose
  double *D, *qs, *x, *xt, *rhs, *tv, *dx;        // n
  double *E, *l, *u, *z, *y, *rho, *w, *dy, *ax, *rinv;  // m = mg + nbx (rinv = 1 / rho: OSQP's rho_inv_vec)
  double *cb;                                     // nbx scaled box coefficient E_b D_j
  double *red;                                    // reduction scratch: 32 warps x 16
  double *sc;                                     // scalars
};
QPC_HD int admm_matrix_doubles(int n, int mg) { return n * n + 2 * mg * n; }
QPC_HD int admm_vector_doubles(int n, int mg, int nbx) {
  const int m = mg + nbx;
  return 7 * n + 10 * m + nbx + QPC_RED_DOUBLES + 32 + 8;
}
QPC_HD int admm_smem_doubles(int n, int mg, int nbx) { return admm_matrix_doubles(n, mg) + admm_vector_doubles(n, mg, nbx); }
// `mat` holds the three matrices (shared memory, or a per-CTA global scratch for QPs that do not fit), `b` the vectors
QPC_HD AdmmSmem admm_layout(double* mat, double* b, int n, int mg, int nbx) {
  const int m = mg + nbx;
  AdmmSmem s;
  s.M = mat;
  s.Gs = mat + n * n;
  s.Gt = s.Gs + mg * n;
  s.D = b;    b += n;
  s.qs = b;   b += n;
  s.x = b;    b += n;
  s.xt = b;   b += n;
  s.rhs = b;  b += n;
  s.tv = b;   b += n;
  s.dx = b;   b += n;
  s.E = b;    b += m;
  s.l = b;    b += m;
  s.u = b;    b += m;
  s.z = b;    b += m;
  s.y = b;    b += m;
  s.rho = b;  b += m;
  s.w = b;    b += m;
  s.dy = b;   b += m;
  s.ax = b;   b += m;
  s.rinv = b; b += m;
  s.cb = b;   b += nbx;
  s.red = b;  b += QPC_RED_DOUBLES;
  s.sc = b;   b += 32;
  return s;
}

// ---- block reductions of K values at once (K <= 16) ------------------------------------------------------------------
template <int K, bool IS_MAX>
QPC_DEV void block_reduce(double* v, double* red) {
#if !defined(QPC_SERIAL)
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = (blockDim.x + 31) >> 5;
#pragma unroll
  for (int k = 0; k < K; k++) {
    double a = v[k];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      double b = __shfl_xor_sync(0xffffffffu, a, o);
      a = IS_MAX ? fmax(a, b) : a + b;
    }
    if (lane == 0) red[warp * 16 + k] = a;
  }
  __syncthreads();
#pragma unroll
  for (int k = 0; k < K; k++) {
    double a = red[k];
    for (int w = 1; w < nw; w++) a = IS_MAX ? fmax(a, red[w * 16 + k]) : a + red[w * 16 + k];
    v[k] = a;
  }
  __syncthreads();
#else
  (void)v;
  (void)red;
#endif
}

// sum of `acc` over the R adjacent lanes that share one output (R in {1,2,4,8}); every lane of the warp must call
QPC_DEV double lane_group_sum(double acc, int R) {
#if !defined(QPC_SERIAL)
  for (int o = R >> 1; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
#else
  (void)R;
#endif
  return acc;
}
QPC_DEV int split_factor(int nout) {
#if !defined(QPC_SERIAL)
  int R = 1;
  while (R < 8 && nout * R * 2 <= (int)blockDim.x) R *= 2;
  return R;
#else
  (void)nout;
  return 1;
#endif
}

QPC_DEV bool finite_val(double v) { return fabs(v) <= 1.7e308; }  // false for NaN and +-Inf

QPC_DEV double limit_scaling(double v) {
  v = v < QPC_MIN_SCALING ? 1.0 : v;
  return v > QPC_MAX_SCALING ? QPC_MAX_SCALING : v;
}

struct AdmmProblem {  // this instance's slots in global memory
  const double *P, *qv, *G, *lg, *ug, *lb, *ub;
  double *x, *y;
  int *status, *iters;
  double* res;
  int* nfac = nullptr;
  // warm start (optional): previous unscaled x [n], y [mg+nbx] and the rho they ended with (<= 0: start cold);
  // rho_io receives this solve's final rho, or -1 when the solve was not accepted
  const double *x0 = nullptr, *y0 = nullptr;
  double* rho_io = nullptr;
};

// out[j] = sum_i Gs[i][j] w[i] (+ box term), j < n : one thread group per column, conflict-free column walks
QPC_DEV void admm_At_times(const AdmmSmem& s, int n, int mg, int nbx, const double* w, double* out, double sigma_x_scale,
                           const double* xadd, const double* qsub) {
  const int R = split_factor(n), G = QPC_NT / R;
  const int sidx = QPC_TID % R;
  for (int base = 0; base < n; base += G) {
    const int j = base + QPC_TID / R;
    double a = 0;
    if (j < n) {
      QPC_UNROLL8
      for (int i = sidx; i < mg; i += R) a += s.Gs[i * n + j] * w[i];
    }
    a = lane_group_sum(a, R);
    if (j < n && sidx == 0) {
      if (j >= n - nbx) a += s.cb[j - (n - nbx)] * w[mg + j - (n - nbx)];
      if (xadd) a += sigma_x_scale * xadd[j] - qsub[j];
      out[j] = a;
    }
  }
}

// out[i] = (A v)[i], i < m
QPC_DEV void admm_A_times(const AdmmSmem& s, int n, int mg, int nbx, const double* v, double* out) {
  const int R = split_factor(mg > 0 ? mg : 1), G = QPC_NT / R;
  const int sidx = QPC_TID % R;
  for (int base = 0; base < mg; base += G) {
    const int i = base + QPC_TID / R;
    double a = 0;
    if (i < mg) {
      QPC_UNROLL8
      for (int j = sidx; j < n; j += R) a += s.Gt[j * mg + i] * v[j];
    }
    a = lane_group_sum(a, R);
    if (i < mg && sidx == 0) out[i] = a;
  }
  for (int i = QPC_TID; i < nbx; i += QPC_NT) out[mg + i] = s.cb[i] * v[n - nbx + i];
}

// P_bar v = c D (P (D v)) from the unscaled P in global memory (symmetric: coalesced column walks)
QPC_DEV void admm_P_times(const AdmmSmem& s, const double* __restrict__ P, int n, double c, const double* v, double* tmp,
                          double* out) {
  for (int j = QPC_TID; j < n; j += QPC_NT) tmp[j] = s.D[j] * v[j];
  QPC_SYNC();
  for (int i = QPC_TID; i < n; i += QPC_NT) {
    double a = 0;
    QPC_UNROLL8
    for (int j = 0; j < n; j++) a += P[j * n + i] * tmp[j];
    out[i] = c * s.D[i] * a;
  }
  QPC_SYNC();
}

QPC_DEV void admm_set_rho(const AdmmSmem& s, int m, double rho) {
  for (int i = QPC_TID; i < m; i += QPC_NT) {
    double r;
    if (s.l[i] < -QPC_INFTY * QPC_MIN_SCALING && s.u[i] > QPC_INFTY * QPC_MIN_SCALING) r = QPC_RHO_MIN;
    else if (s.u[i] - s.l[i] < QPC_RHO_TOL) r = QPC_RHO_EQ_FACTOR * rho;
    else r = rho;
    s.rho[i] = r;
    s.rinv[i] = 1.0 / r;
  }
  QPC_SYNC();
}

// S = P_bar + sigma I + A' rho A -> Cholesky -> L^-1 mirrored.  If `reload`, P_bar is rebuilt from global P.
QPC_DEV void admm_factor(const AdmmSmem& s, const double* __restrict__ P, int n, int mg, int nbx, double c, double sigma,
                         bool reload) {
  const int tid = QPC_TID, nt = QPC_NT;
  double* M = s.M;
  if (reload) {
    for (int k = tid; k < n * n; k += nt) M[k] = c * s.D[k / n] * s.D[k % n] * P[k];
    QPC_SYNC();
  }
  // lower triangle += sum_k Gs[k][i] rho_k Gs[k][j]; diagonal += sigma (+ rho_b cb^2)
  for (int k = tid; k < n * n; k += nt) {
    const int i = k / n, j = k % n;
    if (j > i) continue;
    double a = M[k];
    QPC_UNROLL8
    for (int r = 0; r < mg; r++) a += s.Gs[r * n + i] * s.rho[r] * s.Gs[r * n + j];
    if (i == j) {
      a += sigma;
      if (i >= n - nbx) {
        const int b = i - (n - nbx);
        a += s.rho[mg + b] * s.cb[b] * s.cb[b];
      }
    }
    M[k] = a;
  }
  QPC_SYNC();
  // LDL'-style right-looking elimination with the column scaling deferred: one barrier per column
  for (int k = 0; k < n; k++) {
    const double dinv = 1.0 / M[k * n + k];
    const int rem = n - k - 1;
    for (int e = tid; e < rem * rem; e += nt) {
      const int i = k + 1 + e / rem, j = k + 1 + e % rem;
      if (j <= i) M[i * n + j] -= M[i * n + k] * M[j * n + k] * dinv;
    }
    QPC_SYNC();
  }
  // L[i][k] = M[i][k] / sqrt(d_k), L[k][k] = sqrt(d_k)
  for (int k = tid; k < n * n; k += nt) {
    const int i = k / n, j = k % n;
    if (j < i) M[k] = M[k] / sqrt(M[j * n + j]);
  }
  QPC_SYNC();
  for (int k = tid; k < n; k += nt) M[k * n + k] = sqrt(M[k * n + k]);
  QPC_SYNC();
  // in-place inverse of the lower-triangular factor, last column first
  double* col = s.tv;
  for (int j = n - 1; j >= 0; j--) {
    const double dj = 1.0 / M[j * n + j];
    for (int k = j + 1 + tid; k < n; k += nt) col[k] = M[k * n + j];
    QPC_SYNC();
    for (int i = j + 1 + tid; i < n; i += nt) {
      double a = 0;
      QPC_UNROLL8
      for (int k = j + 1; k <= i; k++) a += M[i * n + k] * col[k];
      M[i * n + j] = -a * dj;
    }
    if (tid == 0) M[j * n + j] = dj;
    QPC_SYNC();
  }
  for (int k = tid; k < n * n; k += nt) {
    const int i = k / n, j = k % n;
    if (j > i) M[k] = M[j * n + i];
  }
  QPC_SYNC();
}

// The solver.  `smem` must hold admm_smem_doubles(n, mg, nbx) doubles, or only admm_vector_doubles when `gmat`
// (admm_matrix_doubles of global scratch owned by this CTA) is given.
// CN / CMG / CNBX >= 0 fix the dimensions at compile time (the thread-per-instance tick of tiny_thread.cu: every loop of a
// 2 x 3 QP unrolls); -1 = the run-time arguments.
template <int CN = -1, int CMG = -1, int CNBX = -1>
QPC_DEV void admm_solve(const Settings& st, const AdmmProblem& pb, int n_, int mg_, int nbx_, double* smem,
                        double* gmat = nullptr) {
  const int n = CN >= 0 ? CN : n_, mg = CMG >= 0 ? CMG : mg_, nbx = CNBX >= 0 ? CNBX : nbx_;
  const int tid = QPC_TID, nt = QPC_NT;
  const int m = mg + nbx, nx0 = n - nbx;
  AdmmSmem s = gmat ? admm_layout(gmat, smem, n, mg, nbx) : admm_layout(smem, smem + admm_matrix_doubles(n, mg), n, mg, nbx);
  // ---- load ------------------------------------------------------------------------------------------------------
  for (int k = tid; k < n * n; k += nt) s.M[k] = pb.P[k];
  for (int k = tid; k < mg * n; k += nt) {
    const double g = pb.G[k];
    s.Gs[k] = g;
    s.Gt[(k % n) * mg + k / n] = g;
  }
  for (int j = tid; j < n; j += nt) {
    s.qs[j] = pb.qv[j];
    s.D[j] = 1.0;
    s.x[j] = 0.0;
    s.dx[j] = 0.0;
  }
  for (int i = tid; i < m; i += nt) {
    s.E[i] = 1.0;
    s.l[i] = fmax(i < mg ? pb.lg[i] : pb.lb[i - mg], -QPC_INFTY);
    s.u[i] = fmin(i < mg ? pb.ug[i] : pb.ub[i - mg], QPC_INFTY);
    s.z[i] = 0.0;
    s.y[i] = 0.0;
    s.dy[i] = 0.0;
  }
  for (int i = tid; i < nbx; i += nt) s.cb[i] = 1.0;
  QPC_SYNC();
  // ---- Ruiz equilibration of [P A'; A 0] ---------------------------------------------------------
  double c = 1.0;
  for (int it = 0; it < st.scaling; it++) {
    for (int j = tid; j < n; j += nt) {  // column norms -> tv
      double a = 0;
      for (int i = 0; i < n; i++) a = fmax(a, fabs(s.M[i * n + j]));
      for (int i = 0; i < mg; i++) a = fmax(a, fabs(s.Gs[i * n + j]));
      if (j >= nx0) a = fmax(a, fabs(s.cb[j - nx0]));
      s.tv[j] = 1.0 / sqrt(limit_scaling(a));
    }
    for (int i = tid; i < m; i += nt) {  // row norms -> ax
      double a = 0;
      if (i < mg)
        for (int j = 0; j < n; j++) a = fmax(a, fabs(s.Gt[j * mg + i]));
      else
        a = fabs(s.cb[i - mg]);
      s.ax[i] = 1.0 / sqrt(limit_scaling(a));
    }
    QPC_SYNC();
    for (int k = tid; k < n * n; k += nt) s.M[k] *= s.tv[k / n] * s.tv[k % n];
    for (int k = tid; k < mg * n; k += nt) {
      const int i = k / n, j = k % n;
      const double f = s.ax[i] * s.tv[j];
      s.Gs[k] *= f;
      s.Gt[j * mg + i] *= f;
    }
    for (int i = tid; i < nbx; i += nt) s.cb[i] *= s.ax[mg + i] * s.tv[nx0 + i];
    for (int j = tid; j < n; j += nt) {
      s.qs[j] *= s.tv[j];
      s.D[j] *= s.tv[j];
    }
    for (int i = tid; i < m; i += nt) s.E[i] *= s.ax[i];
    QPC_SYNC();
    // cost scaling: mean column norm of P_bar vs |q_bar|_inf
    double r2[2] = {0.0, 0.0}, r1[1] = {0.0};
    for (int j = tid; j < n; j += nt) {
      double a = 0;
      for (int i = 0; i < n; i++) a = fmax(a, fabs(s.M[i * n + j]));
      r2[0] += a;
      r1[0] = fmax(r1[0], fabs(s.qs[j]));
    }
    block_reduce<1, false>(r2, s.red);
    block_reduce<1, true>(r1, s.red);
    double ct = limit_scaling(n > 0 ? r2[0] / n : 1.0);
    const double qn = limit_scaling(r1[0]);
    ct = 1.0 / fmax(ct, qn);
    for (int k = tid; k < n * n; k += nt) s.M[k] *= ct;
    for (int j = tid; j < n; j += nt) s.qs[j] *= ct;
    c *= ct;
    QPC_SYNC();
  }
  const double cinv = 1.0 / c;
  for (int i = tid; i < m; i += nt) {
    s.l[i] *= s.E[i];
    s.u[i] *= s.E[i];
  }
  QPC_SYNC();
  double rho = st.rho;
  if (pb.rho_io && pb.x0 && pb.y0 && *pb.rho_io > 0.0) {  // OSQP's implicit warm start (see admm_reg.cuh)
    rho = *pb.rho_io;
    for (int j = tid; j < n; j += nt) s.x[j] = pb.x0[j] / s.D[j];
    for (int i = tid; i < m; i += nt) s.y[i] = pb.y0[i] * c / s.E[i];
    QPC_SYNC();
    admm_A_times(s, n, mg, nbx, s.x, s.z);
    QPC_SYNC();
  }
  admm_set_rho(s, m, rho);
  admm_factor(s, pb.P, n, mg, nbx, c, st.sigma, false);
  for (int i = tid; i < m; i += nt) s.w[i] = s.rho[i] * s.z[i] - s.y[i];
  QPC_SYNC();

  int status = -10, iter = 0, nfac = 1;
  double pri_res = 0, dua_res = 0;
  const double alpha = st.alpha;
  const int Rn = split_factor(n), Gn = nt / Rn, sn = tid % Rn;
  int to_check = st.check_termination, to_adapt = st.adaptive_rho ? st.adaptive_rho_interval : 0;  // countdowns (no modulo)
  for (iter = 1; iter <= st.max_iter; iter++) {
    // rhs = sigma x - q + A'(rho z - y)
    admm_At_times(s, n, mg, nbx, s.w, s.rhs, st.sigma, s.x, s.qs);
    QPC_SYNC();
    // tv = L^-1 rhs  (upper triangle holds L^-T: column walk of row k <= i)
    for (int base = 0; base < n; base += Gn) {
      const int i = base + tid / Rn;
      double a = 0;
      if (i < n) {
        QPC_UNROLL8
        for (int k = sn; k <= i; k += Rn) a += s.M[k * n + i] * s.rhs[k];
      }
      a = lane_group_sum(a, Rn);
      if (i < n && sn == 0) s.tv[i] = a;
    }
    QPC_SYNC();
    // xt = L^-T tv
    for (int base = 0; base < n; base += Gn) {
      const int j = base + tid / Rn;
      double a = 0;
      if (j < n) {
        QPC_UNROLL8
        for (int i = j + sn; i < n; i += Rn) a += s.M[i * n + j] * s.tv[i];
      }
      a = lane_group_sum(a, Rn);
      if (j < n && sn == 0) s.xt[j] = a;
    }
    QPC_SYNC();
    // zt = A xt -> ax
    admm_A_times(s, n, mg, nbx, s.xt, s.ax);
    for (int j = tid; j < n; j += nt) {
      const double xn = alpha * s.xt[j] + (1 - alpha) * s.x[j];
      s.dx[j] = xn - s.x[j];
      s.x[j] = xn;
    }
    QPC_SYNC();
    for (int i = tid; i < m; i += nt) {
      const double zr = alpha * s.ax[i] + (1 - alpha) * s.z[i];
      const double zn = fmin(fmax(zr + s.y[i] * s.rinv[i], s.l[i]), s.u[i]);
      const double d = s.rho[i] * (zr - zn);
      s.dy[i] = d;
      s.y[i] += d;
      s.z[i] = zn;
      s.w[i] = s.rho[i] * zn - s.y[i];
    }
    QPC_SYNC();
    const bool check = st.check_termination && --to_check == 0;
    const bool adapt = st.adaptive_rho && st.adaptive_rho_interval && --to_adapt == 0;
    if (check) to_check = st.check_termination;
    if (adapt) to_adapt = st.adaptive_rho_interval;
    if (!check && !adapt && iter != st.max_iter) continue;
    // ---- residuals --------------------------------------------------------------------------
    double* Px = s.xt;   // free between iterations
    double* Aty = s.rhs;
    admm_A_times(s, n, mg, nbx, s.x, s.ax);
    admm_At_times(s, n, mg, nbx, s.y, Aty, 0.0, nullptr, nullptr);
    admm_P_times(s, pb.P, n, c, s.x, s.tv, Px);
    double mx[12];
    for (int k = 0; k < 12; k++) mx[k] = 0.0;
    for (int i = tid; i < m; i += nt) {
      const double einv = 1.0 / s.E[i];
      const double r = s.ax[i] - s.z[i];
      mx[0] = fmax(mx[0], fabs(einv * r));        // unscaled primal residual
      mx[1] = fmax(mx[1], fabs(r));               // scaled
      mx[2] = fmax(mx[2], fabs(einv * s.z[i]));
      mx[3] = fmax(mx[3], fabs(einv * s.ax[i]));
      mx[4] = fmax(mx[4], fabs(s.z[i]));
      mx[5] = fmax(mx[5], fabs(s.ax[i]));
    }
    for (int j = tid; j < n; j += nt) {
      const double dinv = 1.0 / s.D[j];
      const double r = Px[j] + s.qs[j] + Aty[j];
      mx[6] = fmax(mx[6], fabs(dinv * r));        // unscaled dual residual (times c)
      mx[7] = fmax(mx[7], fabs(r));
      mx[8] = fmax(mx[8], fmax(fabs(dinv * s.qs[j]), fmax(fabs(dinv * Aty[j]), fabs(dinv * Px[j]))));
      mx[9] = fmax(mx[9], fmax(fabs(s.qs[j]), fmax(fabs(Aty[j]), fabs(Px[j]))));
      mx[10] = (mx[10] != 0.0 || !finite_val(s.x[j])) ? 1.0 : 0.0;
    }
    block_reduce<11, true>(mx, s.red);
    pri_res = mx[0];
    dua_res = cinv * mx[6];
    bool nonfinite = mx[10] != 0.0 || !finite_val(pri_res) || !finite_val(dua_res);
    if (nonfinite) {
      status = -8;
      break;
    }
    if (check || iter == st.max_iter) {
      int decided = 0;
      for (int pass = 0; pass < 2 && !decided; pass++) {
        if (pass == 1 && iter != st.max_iter) break;  // the 10x relaxed test only applies at the iteration limit
        const double f = pass ? 10.0 : 1.0;
        const double eps_abs = f * st.eps_abs, eps_rel = f * st.eps_rel;
        const double epi = f * st.eps_prim_inf, edi = f * st.eps_dual_inf;
        const bool prim_ok = m == 0 || pri_res < eps_abs + eps_rel * fmax(mx[2], mx[3]);
        const bool dual_ok = dua_res < eps_abs + eps_rel * cinv * mx[8];
        if (prim_ok && dual_ok) {
          status = pass ? 2 : 1;
          decided = 1;
          break;
        }
        if (!prim_ok) {
          // primal infeasibility certificate from delta_y
          double v2[2] = {0.0, 0.0};
          for (int i = tid; i < m; i += nt) {
            double d = s.dy[i];
            if (s.u[i] > QPC_INFTY * QPC_MIN_SCALING) d = (s.l[i] < -QPC_INFTY * QPC_MIN_SCALING) ? 0.0 : fmin(d, 0.0);
            else if (s.l[i] < -QPC_INFTY * QPC_MIN_SCALING) d = fmax(d, 0.0);
            s.ax[i] = d;
            v2[0] = fmax(v2[0], fabs(s.E[i] * d));
          }
          block_reduce<1, true>(v2, s.red);
          const double nrm = v2[0];
          if (nrm > epi) {
            double sm[1] = {0.0};
            for (int i = tid; i < m; i += nt) sm[0] += s.u[i] * fmax(s.ax[i], 0.0) + s.l[i] * fmin(s.ax[i], 0.0);
            block_reduce<1, false>(sm, s.red);
            if (sm[0] < -epi * nrm) {
              admm_At_times(s, n, mg, nbx, s.ax, s.tv, 0.0, nullptr, nullptr);
              QPC_SYNC();
              double na[1] = {0.0};
              for (int j = tid; j < n; j += nt) na[0] = fmax(na[0], fabs(s.tv[j] / s.D[j]));
              block_reduce<1, true>(na, s.red);
              if (na[0] < epi * nrm) {
                status = pass ? 3 : -3;
                decided = 1;
                break;
              }
            }
          }
        }
        if (!dual_ok) {
          // dual infeasibility certificate from delta_x
          double v1[1] = {0.0};
          for (int j = tid; j < n; j += nt) v1[0] = fmax(v1[0], fabs(s.D[j] * s.dx[j]));
          block_reduce<1, true>(v1, s.red);
          const double nrm = v1[0];
          if (nrm > edi) {
            double sm[1] = {0.0};
            for (int j = tid; j < n; j += nt) sm[0] += s.qs[j] * s.dx[j];
            block_reduce<1, false>(sm, s.red);
            if (sm[0] < -c * edi * nrm) {
              admm_P_times(s, pb.P, n, c, s.dx, s.tv, Px);
              double np[1] = {0.0};
              for (int j = tid; j < n; j += nt) np[0] = fmax(np[0], fabs(Px[j] / s.D[j]));
              block_reduce<1, true>(np, s.red);
              if (np[0] < c * edi * nrm) {
                admm_A_times(s, n, mg, nbx, s.dx, s.ax);
                QPC_SYNC();
                double bad[1] = {0.0};
                for (int i = tid; i < m; i += nt) {
                  const double a = s.ax[i] / s.E[i];
                  if ((s.u[i] < QPC_INFTY * QPC_MIN_SCALING && a > edi * nrm) ||
                      (s.l[i] > -QPC_INFTY * QPC_MIN_SCALING && a < -edi * nrm))
                    bad[0] = 1.0;
                }
                block_reduce<1, true>(bad, s.red);
                if (bad[0] == 0.0) {
                  status = pass ? 4 : -4;
                  decided = 1;
                  break;
                }
              }
            }
          }
        }
      }
      if (decided) break;
      if (iter == st.max_iter) {
        status = -2;
        break;
      }
    }
    if (adapt) {
      // rho <- rho sqrt( (r_p / max(|Ax|,|z|)) / (r_d / max(|Px|,|A'y|,|q|)) ) on scaled quantities (B.3 step 6)
      const double pr = mx[1] / (fmax(mx[4], mx[5]) + 1e-10);
      const double dr = mx[7] / (mx[9] + 1e-10);
      double rho_new = rho * sqrt(pr / (dr + 1e-10));
      rho_new = fmin(fmax(rho_new, QPC_RHO_MIN), QPC_RHO_MAX);
      if (rho_new > rho * st.adaptive_rho_tolerance || rho_new < rho / st.adaptive_rho_tolerance) {
        rho = rho_new;
        admm_set_rho(s, m, rho);
        admm_factor(s, pb.P, n, mg, nbx, c, st.sigma, true);
        nfac++;
        for (int i = tid; i < m; i += nt) s.w[i] = s.rho[i] * s.z[i] - s.y[i];
        QPC_SYNC();
      }
    }
  }
  if (iter > st.max_iter) iter = st.max_iter;
  // ---- unscale and store ---------------------------------------------------------------------------------------------
  for (int j = tid; j < n; j += nt) pb.x[j] = s.D[j] * s.x[j];
  if (pb.y)
    for (int i = tid; i < m; i += nt) pb.y[i] = cinv * s.E[i] * s.y[i];
  if (tid == 0) {
    *pb.status = status;
    if (pb.iters) *pb.iters = iter;
    if (pb.nfac) *pb.nfac = nfac;
    if (pb.rho_io) *pb.rho_io = (status == 1 || status == 2) ? rho : -1.0;
    if (pb.res) {
      pb.res[0] = pri_res;
      pb.res[1] = dua_res;
    }
  }
  QPC_SYNC();
}

}  // namespace qpc
