// ORACLE -- TEST INFRASTRUCTURE ONLY.  Never linked into, imported by, or called from the product path.
//
// CPU fp64 restatement of the OSQP 0.5.x ADMM algorithm (Stellato et al., "OSQP: an operator splitting solver for
// quadratic programs") that the reference reaches through `solve!(qpmodel)` (reference src/lowlevel/momentum.jl:58,
// optimizer configured in test/runtests.jl:35-43 and notebooks/Standing controller.ipynb:66-71).  OSQP.jl 0.5.2 /
// libosqp is an un-vendored binary dependency (reference test/Manifest.toml:440-444), so this restates the published
// algorithm with OSQP's default constants: Ruiz equilibration, rho vector with
// equality-row boost, sparse quasi-definite KKT LDL' factorisation with a fill-reducing ordering, relaxed ADMM
// iteration, unscaled termination residuals, infeasibility certificates, adaptive rho with refactorisation,
// implicit warm start between solves.  Iteration-for-iteration equality with libosqp is not a goal (different
// ordering / summation order); the same optimum and the same termination criteria are.
//
// PARITY UNPINNED: no reference outputs are available offline; tests pin this solver through KKT-optimality checks
// done independently in numpy.
#pragma once
#include <algorithm>
#include <cmath>
#include <cstdint>
#include <set>
#include <vector>

namespace orc {

struct Csc {
  int m = 0, n = 0;
  std::vector<int> p, i;
  std::vector<double> x;
  int nnz() const { return (int)i.size(); }
};

struct OsqpSettings {
  double rho = 0.1, sigma = 1e-6, alpha = 1.6;
  double eps_abs = 1e-3, eps_rel = 1e-3, eps_prim_inf = 1e-4, eps_dual_inf = 1e-4;
  int max_iter = 4000, scaling = 10, adaptive_rho = 1, adaptive_rho_interval = 25, check_termination = 25;
  double adaptive_rho_tolerance = 5.0;
  int warm_start = 1;
};

enum OsqpStatus {
  OSQP_SOLVED = 1,
  OSQP_SOLVED_INACCURATE = 2,
  OSQP_PRIMAL_INFEASIBLE_INACCURATE = 3,
  OSQP_DUAL_INFEASIBLE_INACCURATE = 4,
  OSQP_MAX_ITER_REACHED = -2,
  OSQP_PRIMAL_INFEASIBLE = -3,
  OSQP_DUAL_INFEASIBLE = -4,
  OSQP_NON_FINITE = -8,
  OSQP_UNSOLVED = -10
};

constexpr double OSQP_INFTY = 1e20, RHO_MIN = 1e-6, RHO_MAX = 1e6, RHO_EQ_OVER_RHO_INEQ = 1e3, RHO_TOL = 1e-4,
                 MIN_SCALING = 1e-4, MAX_SCALING = 1e4;

// Sparse LDL' of a symmetric quasi-definite matrix given by its upper triangle (CSC), after a symmetric
// minimum-degree permutation computed once from the pattern.  Up-looking factorisation driven by the elimination
// tree (Davis, "Algorithm 849").
class SparseLdl {
 public:
  int n = 0;
  std::vector<int> perm, pinv;           // perm[new] = old
  std::vector<int> Kp, Ki;               // permuted upper-triangular pattern
  std::vector<double> Kx;
  std::vector<int> etree, Lp, Li, Lnz;
  std::vector<double> Lx, D, Dinv;
  std::vector<int> flag, stack, pattern;
  std::vector<double> y, work;

  // pattern given as triplets (r <= c) of the upper triangle; returns for every triplet its slot in Kx
  std::vector<int> analyze(int n_, const std::vector<int>& rows, const std::vector<int>& cols) {
    n = n_;
    min_degree(rows, cols);
    // permuted triplets, sorted by column then row
    size_t nz = rows.size();
    std::vector<int> pr(nz), pc(nz);
    for (size_t k = 0; k < nz; k++) {
      int a = pinv[rows[k]], b = pinv[cols[k]];
      pr[k] = std::min(a, b);
      pc[k] = std::max(a, b);
    }
    std::vector<int> order(nz);
    for (size_t k = 0; k < nz; k++) order[k] = (int)k;
    std::sort(order.begin(), order.end(), [&](int a, int b) { return pc[a] != pc[b] ? pc[a] < pc[b] : pr[a] < pr[b]; });
    Kp.assign(n + 1, 0);
    Ki.clear();
    std::vector<int> slot(nz);
    int lastc = -1, lastr = -1;
    for (size_t t = 0; t < nz; t++) {
      int k = order[t];
      if (pc[k] == lastc && pr[k] == lastr) {  // duplicate entry: shares the slot
        slot[k] = (int)Ki.size() - 1;
        continue;
      }
      Ki.push_back(pr[k]);
      Kp[pc[k] + 1]++;
      slot[k] = (int)Ki.size() - 1;
      lastc = pc[k];
      lastr = pr[k];
    }
    for (int j = 0; j < n; j++) Kp[j + 1] += Kp[j];
    Kx.assign(Ki.size(), 0.0);
    symbolic();
    return slot;
  }

  bool factor() {
    // row k of L: solve L(0:k-1,0:k-1) D y = K(0:k-1,k)
    for (int k = 0; k < n; k++) Lnz[k] = 0;
    for (int k = 0; k < n; k++) {
      int top = n;
      flag[k] = k;
      y[k] = 0;
      double dk = 0;
      for (int p = Kp[k]; p < Kp[k + 1]; p++) {
        int i = Ki[p];
        if (i == k) {
          dk += Kx[p];
          continue;
        }
        y[i] += Kx[p];
        int len = 0;
        for (; flag[i] != k; i = etree[i]) {
          pattern[len++] = i;
          flag[i] = k;
        }
        while (len > 0) stack[--top] = pattern[--len];
      }
      for (; top < n; top++) {
        int i = stack[top];
        double yi = y[i];
        y[i] = 0;
        int p2 = Lp[i] + Lnz[i];
        for (int p = Lp[i]; p < p2; p++) y[Li[p]] -= Lx[p] * yi;
        double lki = yi * Dinv[i];
        dk -= lki * yi;
        Li[p2] = k;
        Lx[p2] = lki;
        Lnz[i]++;
      }
      if (dk == 0.0 || !std::isfinite(dk)) return false;
      D[k] = dk;
      Dinv[k] = 1.0 / dk;
    }
    return true;
  }

  // b (original ordering) overwritten with K^{-1} b
  void solve(double* b) {
    for (int i = 0; i < n; i++) work[i] = b[perm[i]];
    for (int j = 0; j < n; j++) {
      double xj = work[j];
      for (int p = Lp[j]; p < Lp[j] + Lnz[j]; p++) work[Li[p]] -= Lx[p] * xj;
    }
    for (int j = 0; j < n; j++) work[j] *= Dinv[j];
    for (int j = n - 1; j >= 0; j--) {
      double xj = work[j];
      for (int p = Lp[j]; p < Lp[j] + Lnz[j]; p++) xj -= Lx[p] * work[Li[p]];
      work[j] = xj;
    }
    for (int i = 0; i < n; i++) b[perm[i]] = work[i];
  }
  int64_t nnzL() const { return Lp.empty() ? 0 : Lp[n]; }

 private:
  void min_degree(const std::vector<int>& rows, const std::vector<int>& cols) {
    std::vector<std::set<int>> adj(n);
    for (size_t k = 0; k < rows.size(); k++)
      if (rows[k] != cols[k]) {
        adj[rows[k]].insert(cols[k]);
        adj[cols[k]].insert(rows[k]);
      }
    std::vector<char> done(n, 0);
    perm.assign(n, 0);
    pinv.assign(n, 0);
    for (int step = 0; step < n; step++) {
      int best = -1;
      size_t bestdeg = SIZE_MAX;
      for (int v = 0; v < n; v++)
        if (!done[v] && adj[v].size() < bestdeg) {
          bestdeg = adj[v].size();
          best = v;
        }
      done[best] = 1;
      perm[step] = best;
      pinv[best] = step;
      std::vector<int> nb(adj[best].begin(), adj[best].end());
      for (int a : nb) adj[a].erase(best);
      for (size_t a = 0; a < nb.size(); a++)
        for (size_t b = a + 1; b < nb.size(); b++) {
          adj[nb[a]].insert(nb[b]);
          adj[nb[b]].insert(nb[a]);
        }
      adj[best].clear();
    }
  }
  void symbolic() {
    etree.assign(n, -1);
    Lnz.assign(n, 0);
    flag.assign(n, -1);
    for (int k = 0; k < n; k++) {
      flag[k] = k;
      for (int p = Kp[k]; p < Kp[k + 1]; p++)
        for (int i = Ki[p]; flag[i] != k; i = etree[i]) {
          if (etree[i] == -1) etree[i] = k;
          Lnz[i]++;
          flag[i] = k;
        }
    }
    Lp.assign(n + 1, 0);
    for (int k = 0; k < n; k++) Lp[k + 1] = Lp[k] + Lnz[k];
    Li.assign(Lp[n], 0);
    Lx.assign(Lp[n], 0.0);
    D.assign(n, 0.0);
    Dinv.assign(n, 0.0);
    stack.assign(n, 0);
    pattern.assign(n, 0);
    y.assign(n, 0.0);
    work.assign(n, 0.0);
    flag.assign(n, -1);
  }
};

struct OsqpInfo {
  int status = OSQP_UNSOLVED, iter = 0, rho_updates = 0, factorizations = 0;
  double pri_res = 0, dua_res = 0, obj = 0, rho = 0;
};

class OsqpPort {
 public:
  OsqpSettings st;
  int n = 0, m = 0;
  Csc P, A;  // scaled working copies (P upper triangular)
  std::vector<double> q, l, u;
  std::vector<double> D, E, Dinv, Einv;
  double c = 1, cinv = 1;
  std::vector<double> x, z, y, xprev, zprev, xtilde_z, dx, dy, Ax, Px, Aty, rho_vec, rho_inv_vec, rhs;
  std::vector<int> constr_type;
  double rho = 0.1;
  SparseLdl kkt;
  std::vector<int> slotP, slotA, slotDiagX, slotDiagZ;
  OsqpInfo info;
  bool has_solution = false;
  int nfactor = 0;

  // Pu: upper triangle of P. Pattern is fixed from here on (explicit zeros allowed, as Parametron keeps them).
  void setup(const Csc& Pu, const double* q_, const Csc& A_, const double* l_, const double* u_,
             const OsqpSettings& s) {
    st = s;
    n = Pu.n;
    m = A_.m;
    P = Pu;
    A = A_;
    q.assign(q_, q_ + n);
    l.assign(l_, l_ + m);
    u.assign(u_, u_ + m);
    for (auto* vec : {&x, &xprev, &dx, &Px, &Aty}) vec->assign(n, 0.0);
    for (auto* vec : {&z, &y, &zprev, &dy, &Ax, &rho_vec, &rho_inv_vec}) vec->assign(m, 0.0);
    xtilde_z.assign(n + m, 0.0);
    rhs.assign(n + m, 0.0);
    constr_type.assign(m, 0);
    D.assign(n, 1.0);
    Dinv.assign(n, 1.0);
    E.assign(m, 1.0);
    Einv.assign(m, 1.0);
    c = cinv = 1;
    scale_data();
    rho = st.rho;
    set_rho_vec();
    // KKT pattern: [P + sigma I, A'; A, -1/rho]
    std::vector<int> rows, cols;
    for (int j = 0; j < n; j++)
      for (int p = P.p[j]; p < P.p[j + 1]; p++) {
        rows.push_back(P.i[p]);
        cols.push_back(j);
      }
    for (int j = 0; j < n; j++) {
      rows.push_back(j);
      cols.push_back(j);
    }
    for (int j = 0; j < n; j++)
      for (int p = A.p[j]; p < A.p[j + 1]; p++) {
        rows.push_back(j);
        cols.push_back(n + A.i[p]);
      }
    for (int i = 0; i < m; i++) {
      rows.push_back(n + i);
      cols.push_back(n + i);
    }
    std::vector<int> slot = kkt.analyze(n + m, rows, cols);
    int k = 0;
    slotP.assign(slot.begin() + k, slot.begin() + k + P.nnz());
    k += P.nnz();
    slotDiagX.assign(slot.begin() + k, slot.begin() + k + n);
    k += n;
    slotA.assign(slot.begin() + k, slot.begin() + k + A.nnz());
    k += A.nnz();
    slotDiagZ.assign(slot.begin() + k, slot.begin() + k + m);
    nfactor = 0;
    refactor();
    has_solution = false;
  }

  // osqp_update_P_A / lin_cost / bounds with unscaled data of the same pattern: unscale -> replace -> rescale ->
  // refactor.  Iterates and rho persist (OSQP's implicit warm start) unless st.warm_start == 0.
  void update(const double* Px_, const double* q_, const double* Ax_, const double* l_, const double* u_) {
    P.x.assign(Px_, Px_ + P.nnz());
    A.x.assign(Ax_, Ax_ + A.nnz());
    q.assign(q_, q_ + n);
    l.assign(l_, l_ + m);
    u.assign(u_, u_ + m);
    D.assign(n, 1.0);
    E.assign(m, 1.0);
    c = 1;
    scale_data();
    if (!st.warm_start) {  // cold start: zero iterates, initial rho
      std::fill(x.begin(), x.end(), 0.0);
      std::fill(z.begin(), z.end(), 0.0);
      std::fill(y.begin(), y.end(), 0.0);
      rho = st.rho;
    }
    set_rho_vec();
    nfactor = 0;
    refactor();
  }

  // unscaled warm start
  void warm_start(const double* x0, const double* y0) {
    for (int j = 0; j < n; j++) x[j] = x0[j] * Dinv[j];
    for (int i = 0; i < m; i++) y[i] = y0[i] * Einv[i] * c;
    mat_vec(A, x.data(), z.data());
  }

  int solve() {
    info = OsqpInfo{};
    int iter;
    bool done = false;
    for (iter = 1; iter <= st.max_iter; iter++) {
      xprev = x;
      zprev = z;
      // update_xz_tilde
      for (int j = 0; j < n; j++) rhs[j] = st.sigma * xprev[j] - q[j];
      for (int i = 0; i < m; i++) rhs[n + i] = zprev[i] - rho_inv_vec[i] * y[i];
      kkt.solve(rhs.data());
      for (int i = 0; i < m; i++) xtilde_z[n + i] = zprev[i] + rho_inv_vec[i] * (rhs[n + i] - y[i]);
      for (int j = 0; j < n; j++) xtilde_z[j] = rhs[j];
      // update_x, update_z, update_y
      for (int j = 0; j < n; j++) {
        x[j] = st.alpha * xtilde_z[j] + (1 - st.alpha) * xprev[j];
        dx[j] = x[j] - xprev[j];
      }
      for (int i = 0; i < m; i++) {
        double zr = st.alpha * xtilde_z[n + i] + (1 - st.alpha) * zprev[i];
        double zi = zr + rho_inv_vec[i] * y[i];
        zi = std::min(std::max(zi, l[i]), u[i]);
        z[i] = zi;
        dy[i] = rho_vec[i] * (zr - zi);
        y[i] += dy[i];
      }
      bool can_check = st.check_termination && (iter % st.check_termination == 0);
      bool updated = false;
      if (can_check) {
        update_info(iter);
        updated = true;
        if (check_termination(false)) {
          done = true;
          break;
        }
      }
      if (st.adaptive_rho && st.adaptive_rho_interval && (iter % st.adaptive_rho_interval == 0)) {
        if (!updated) update_info(iter);
        adapt_rho();
      }
    }
    if (!done) {
      iter = st.max_iter;
      update_info(iter);
      if (!check_termination(false)) {
        if (!check_termination(true)) info.status = OSQP_MAX_ITER_REACHED;
      }
    }
    info.iter = std::min(iter, st.max_iter);
    info.rho = rho;
    info.factorizations = nfactor;
    for (int j = 0; j < n; j++)
      if (!std::isfinite(x[j])) info.status = OSQP_NON_FINITE;
    has_solution = true;
    return info.status;
  }

  void solution(double* xo, double* yo) const {
    for (int j = 0; j < n; j++) xo[j] = D[j] * x[j];
    if (yo)
      for (int i = 0; i < m; i++) yo[i] = cinv * E[i] * y[i];
  }

 private:
  static double limit_scaling(double v) {
    v = v < MIN_SCALING ? 1.0 : v;
    return v > MAX_SCALING ? MAX_SCALING : v;
  }
  void scale_data() {
    D.assign(n, 1.0);
    E.assign(m, 1.0);
    c = 1.0;
    std::vector<double> Dt(n), Et(m);
    for (int it = 0; it < st.scaling; it++) {
      // inf-norm of the columns of the KKT matrix [P A'; A 0]
      std::fill(Dt.begin(), Dt.end(), 0.0);
      std::fill(Et.begin(), Et.end(), 0.0);
      for (int j = 0; j < n; j++)
        for (int p = P.p[j]; p < P.p[j + 1]; p++) {
          double a = std::fabs(P.x[p]);
          Dt[j] = std::max(Dt[j], a);
          Dt[P.i[p]] = std::max(Dt[P.i[p]], a);
        }
      for (int j = 0; j < n; j++)
        for (int p = A.p[j]; p < A.p[j + 1]; p++) {
          double a = std::fabs(A.x[p]);
          Dt[j] = std::max(Dt[j], a);
          Et[A.i[p]] = std::max(Et[A.i[p]], a);
        }
      for (int j = 0; j < n; j++) Dt[j] = 1.0 / std::sqrt(limit_scaling(Dt[j]));
      for (int i = 0; i < m; i++) Et[i] = 1.0 / std::sqrt(limit_scaling(Et[i]));
      for (int j = 0; j < n; j++)
        for (int p = P.p[j]; p < P.p[j + 1]; p++) P.x[p] *= Dt[j] * Dt[P.i[p]];
      for (int j = 0; j < n; j++)
        for (int p = A.p[j]; p < A.p[j + 1]; p++) A.x[p] *= Dt[j] * Et[A.i[p]];
      for (int j = 0; j < n; j++) {
        q[j] *= Dt[j];
        D[j] *= Dt[j];
      }
      for (int i = 0; i < m; i++) E[i] *= Et[i];
      // cost scaling: mean column inf-norm of P vs inf-norm of q
      std::vector<double> cn(n, 0.0);
      for (int j = 0; j < n; j++)
        for (int p = P.p[j]; p < P.p[j + 1]; p++) {
          double a = std::fabs(P.x[p]);
          cn[j] = std::max(cn[j], a);
          cn[P.i[p]] = std::max(cn[P.i[p]], a);
        }
      double ct = 0;
      for (int j = 0; j < n; j++) ct += cn[j];
      ct = limit_scaling(n ? ct / n : 1.0);
      double qn = 0;
      for (int j = 0; j < n; j++) qn = std::max(qn, std::fabs(q[j]));
      qn = limit_scaling(qn);
      ct = 1.0 / std::max(ct, qn);
      for (auto& v : P.x) v *= ct;
      for (auto& v : q) v *= ct;
      c *= ct;
    }
    cinv = 1.0 / c;
    for (int j = 0; j < n; j++) Dinv[j] = 1.0 / D[j];
    for (int i = 0; i < m; i++) {
      Einv[i] = 1.0 / E[i];
      l[i] *= E[i];
      u[i] *= E[i];
    }
  }
  void set_rho_vec() {
    for (int i = 0; i < m; i++) {
      if (l[i] < -OSQP_INFTY * MIN_SCALING && u[i] > OSQP_INFTY * MIN_SCALING) {
        constr_type[i] = -1;
        rho_vec[i] = RHO_MIN;
      } else if (u[i] - l[i] < RHO_TOL) {
        constr_type[i] = 1;
        rho_vec[i] = RHO_EQ_OVER_RHO_INEQ * rho;
      } else {
        constr_type[i] = 0;
        rho_vec[i] = rho;
      }
      rho_inv_vec[i] = 1.0 / rho_vec[i];
    }
  }
  void refactor() {
    std::fill(kkt.Kx.begin(), kkt.Kx.end(), 0.0);
    for (int k = 0; k < P.nnz(); k++) kkt.Kx[slotP[k]] += P.x[k];
    for (int j = 0; j < n; j++) kkt.Kx[slotDiagX[j]] += st.sigma;
    for (int k = 0; k < A.nnz(); k++) kkt.Kx[slotA[k]] += A.x[k];
    for (int i = 0; i < m; i++) kkt.Kx[slotDiagZ[i]] += -rho_inv_vec[i];
    kkt.factor();
    nfactor++;
  }
  static void mat_vec(const Csc& M, const double* v, double* out) {
    std::fill(out, out + M.m, 0.0);
    for (int j = 0; j < M.n; j++)
      for (int p = M.p[j]; p < M.p[j + 1]; p++) out[M.i[p]] += M.x[p] * v[j];
  }
  static void mat_tvec(const Csc& M, const double* v, double* out) {
    for (int j = 0; j < M.n; j++) {
      double s = 0;
      for (int p = M.p[j]; p < M.p[j + 1]; p++) s += M.x[p] * v[M.i[p]];
      out[j] = s;
    }
  }
  void sym_vec(const double* v, double* out) {  // P (upper stored) * v
    std::fill(out, out + n, 0.0);
    for (int j = 0; j < n; j++)
      for (int p = P.p[j]; p < P.p[j + 1]; p++) {
        int i = P.i[p];
        out[i] += P.x[p] * v[j];
        if (i != j) out[j] += P.x[p] * v[i];
      }
  }
  // residuals in the units OSQP reports them (unscaled, scaled_termination = 0)
  double sc_pri_res = 0, sc_dua_res = 0;  // scaled versions for the rho estimate
  void update_info(int iter) {
    mat_vec(A, x.data(), Ax.data());
    sym_vec(x.data(), Px.data());
    mat_tvec(A, y.data(), Aty.data());
    double pr = 0, spr = 0;
    for (int i = 0; i < m; i++) {
      double r = Ax[i] - z[i];
      spr = std::max(spr, std::fabs(r));
      pr = std::max(pr, std::fabs(Einv[i] * r));
    }
    double dr = 0, sdr = 0;
    for (int j = 0; j < n; j++) {
      double r = Px[j] + q[j] + Aty[j];
      sdr = std::max(sdr, std::fabs(r));
      dr = std::max(dr, std::fabs(Dinv[j] * r));
    }
    info.pri_res = pr;
    info.dua_res = cinv * dr;
    sc_pri_res = spr;
    sc_dua_res = sdr;
    double obj = 0;
    for (int j = 0; j < n; j++) obj += 0.5 * x[j] * Px[j] + q[j] * x[j];
    info.obj = obj * cinv;
    info.iter = iter;
  }
  bool check_termination(bool approximate) {
    double eps_abs = st.eps_abs, eps_rel = st.eps_rel, epi = st.eps_prim_inf, edi = st.eps_dual_inf;
    if (approximate) {
      eps_abs *= 10;
      eps_rel *= 10;
      epi *= 10;
      edi *= 10;
    }
    if (!std::isfinite(info.pri_res) || !std::isfinite(info.dua_res)) {
      info.status = OSQP_NON_FINITE;
      return true;
    }
    bool prim_ok = false, dual_ok = false, prim_inf = false, dual_inf = false;
    if (m == 0) prim_ok = true;
    else {
      double nz = 0, nAx = 0;
      for (int i = 0; i < m; i++) {
        nz = std::max(nz, std::fabs(Einv[i] * z[i]));
        nAx = std::max(nAx, std::fabs(Einv[i] * Ax[i]));
      }
      double eps_prim = eps_abs + eps_rel * std::max(nz, nAx);
      if (info.pri_res < eps_prim) prim_ok = true;
      else prim_inf = is_primal_infeasible(epi);
    }
    {
      double nq = 0, nAty = 0, nPx = 0;
      for (int j = 0; j < n; j++) {
        nq = std::max(nq, std::fabs(Dinv[j] * q[j]));
        nAty = std::max(nAty, std::fabs(Dinv[j] * Aty[j]));
        nPx = std::max(nPx, std::fabs(Dinv[j] * Px[j]));
      }
      double eps_dual = eps_abs + eps_rel * cinv * std::max(nq, std::max(nAty, nPx));
      if (info.dua_res < eps_dual) dual_ok = true;
      else dual_inf = is_dual_infeasible(edi);
    }
    if (prim_ok && dual_ok) {
      info.status = approximate ? OSQP_SOLVED_INACCURATE : OSQP_SOLVED;
      return true;
    }
    if (prim_inf) {
      info.status = approximate ? OSQP_PRIMAL_INFEASIBLE_INACCURATE : OSQP_PRIMAL_INFEASIBLE;
      return true;
    }
    if (dual_inf) {
      info.status = approximate ? OSQP_DUAL_INFEASIBLE_INACCURATE : OSQP_DUAL_INFEASIBLE;
      return true;
    }
    return false;
  }
  bool is_primal_infeasible(double eps) {
    // project delta_y onto the polar of the recession cone of [l, u]
    std::vector<double> d(m);
    double nrm = 0, lhs = 0;
    for (int i = 0; i < m; i++) {
      double v = dy[i];
      if (u[i] > OSQP_INFTY * MIN_SCALING) {
        if (l[i] < -OSQP_INFTY * MIN_SCALING) v = 0;
        else v = std::min(v, 0.0);
      } else if (l[i] < -OSQP_INFTY * MIN_SCALING) {
        v = std::max(v, 0.0);
      }
      d[i] = v;
      nrm = std::max(nrm, std::fabs(E[i] * v));
    }
    if (nrm <= eps) return false;
    for (int i = 0; i < m; i++) lhs += u[i] * std::max(d[i], 0.0) + l[i] * std::min(d[i], 0.0);
    if (lhs >= -eps * nrm) return false;
    std::vector<double> Atd(n);
    mat_tvec(A, d.data(), Atd.data());
    double na = 0;
    for (int j = 0; j < n; j++) na = std::max(na, std::fabs(Dinv[j] * Atd[j]));
    return na < eps * nrm;
  }
  bool is_dual_infeasible(double eps) {
    double nrm = 0;
    for (int j = 0; j < n; j++) nrm = std::max(nrm, std::fabs(D[j] * dx[j]));
    if (nrm <= eps) return false;
    double qdx = 0;
    for (int j = 0; j < n; j++) qdx += q[j] * dx[j];
    if (qdx >= -c * eps * nrm) return false;
    std::vector<double> Pdx(n), Adx(m);
    sym_vec(dx.data(), Pdx.data());
    double np = 0;
    for (int j = 0; j < n; j++) np = std::max(np, std::fabs(Dinv[j] * Pdx[j]));
    if (np >= c * eps * nrm) return false;
    mat_vec(A, dx.data(), Adx.data());
    for (int i = 0; i < m; i++) {
      double a = Einv[i] * Adx[i];
      if ((u[i] < OSQP_INFTY * MIN_SCALING && a > eps * nrm) || (l[i] > -OSQP_INFTY * MIN_SCALING && a < -eps * nrm))
        return false;
    }
    return true;
  }
  void adapt_rho() {
    double nz = 0, nAx = 0, nq = 0, nAty = 0, nPx = 0;
    for (int i = 0; i < m; i++) {
      nz = std::max(nz, std::fabs(z[i]));
      nAx = std::max(nAx, std::fabs(Ax[i]));
    }
    for (int j = 0; j < n; j++) {
      nq = std::max(nq, std::fabs(q[j]));
      nAty = std::max(nAty, std::fabs(Aty[j]));
      nPx = std::max(nPx, std::fabs(Px[j]));
    }
    double pr = sc_pri_res / (std::max(nz, nAx) + 1e-10);
    double dr = sc_dua_res / (std::max(nq, std::max(nAty, nPx)) + 1e-10);
    double rho_new = rho * std::sqrt(pr / (dr + 1e-10));
    rho_new = std::min(std::max(rho_new, RHO_MIN), RHO_MAX);
    if (rho_new > rho * st.adaptive_rho_tolerance || rho_new < rho / st.adaptive_rho_tolerance) {
      rho = rho_new;
      set_rho_vec();
      refactor();
      info.rho_updates++;
    }
  }
};

}  // namespace orc
