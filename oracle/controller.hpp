// ORACLE -- TEST INFRASTRUCTURE ONLY.  Never linked into, imported by, or called from the product path.
//
// CPU fp64 restatement of QPControl.jl's per-tick control path in the reference's own (lifted, slack-variable) QP
// form: reference src/lowlevel/momentum.jl (whole file), src/tasks.jl (whole file), src/contacts.jl (whole file),
// src/highlevel/standing.jl (whole file).  Parametron's lazy assembly is restated as "re-evaluate every parameter
// once, write the coefficients into a fixed sparsity pattern" (SURVEY.md appendix B.2); every entry of a matrix
// parameter is a structural nonzero, as in the reference.
//
// PARITY UNPINNED (see rbd.hpp).
#pragma once
#include <memory>
#include <string>

#include "osqp_port.hpp"
#include "rbd.hpp"

namespace orc {

enum TaskKind { T_SPATIAL = 0, T_ANGULAR = 1, T_LINEAR = 2, T_POINT = 3, T_JOINT = 4, T_MOMENTUM = 5, T_LINMOM = 6 };
enum TaskMode { M_HARD = 0, M_SCALAR = 1, M_MATRIX = 2 };

struct Task {
  int kind = 0, source = -1, target = -1, frame = -1, joint = -1, mode = 0, dim = 0, des_off = 0;
  V3 point;
  double weight = 0;
  std::vector<double> W;  // dim x dim row-major (matrix mode)
  int var0 = -1;          // first slack variable (weighted modes)
};

struct Contact {
  int body = -1;
  V3 pos, normal;
  double mu = 1;
  int var0 = -1;  // rho[N], f_local[3], wrench_world angular[3], linear[3]   (contacts.jl:46-48)
};

struct StandingParams {  // reference src/highlevel/standing.jl:1-16, defaults :23-29
  bool enabled = false;
  int linmom_task = -1, pelvis_task = -1, pelvis_body = -1;
  std::vector<int> joint_tasks, joints;
  std::vector<double> joint_kp, joint_kd, joint_ref;
  double com_kp = 10, com_kd = 2 * std::sqrt(10.0), pelvis_kp = 20, pelvis_kd = 2 * std::sqrt(20.0);
  V3 comref;
};

inline int task_dim(int kind, const Mechanism& m, int joint) {
  switch (kind) {
    case T_SPATIAL:
    case T_MOMENTUM: return 6;
    case T_JOINT: return m.nvj[joint];
    default: return 3;
  }
}

// rows of task_error = J vd + b - desired    (tasks.jl:31-44,73-84,112-123,153-171,185-189,230-236,258-262)
inline void task_rows(const Mechanism& m, const State& s, const Task& t, double* J /*dim x nv*/, double* b /*dim*/) {
  const int nv = m.nv;
  std::vector<double> J6(6 * nv);
  switch (t.kind) {
    case T_SPATIAL:
    case T_ANGULAR:
    case T_LINEAR: {
      geometric_jacobian(m, s, t.source, t.target, t.frame, J6.data());
      S6 jv = bias_in_frame(s, t.source, t.target, t.frame);
      int r0 = t.kind == T_LINEAR ? 3 : 0;
      for (int r = 0; r < t.dim; r++) {
        std::memcpy(J + r * nv, J6.data() + (r0 + r) * nv, sizeof(double) * nv);
        int rr = r0 + r;
        b[r] = rr < 3 ? jv.w[rr] : jv.v[rr - 3];
      }
      break;
    }
    case T_POINT: {
      // frame = source (base) body frame (tasks.jl:139-140,154)
      Xf to_base = inv(toroot(s, t.source));
      V3 p = (to_base * toroot(s, t.target)).R * t.point + (to_base * toroot(s, t.target)).p;
      geometric_jacobian(m, s, t.source, t.target, t.source, J6.data());
      for (int c = 0; c < nv; c++) {
        V3 w{J6[0 * nv + c], J6[1 * nv + c], J6[2 * nv + c]}, v{J6[3 * nv + c], J6[4 * nv + c], J6[5 * nv + c]};
        V3 col = cross(w, p) + v;  // point_jacobian!: J_lin + J_ang x p
        for (int r = 0; r < 3; r++) J[r * nv + c] = col[r];
      }
      S6 T = xmotion(to_base, twist_of(s, t.target) - twist_of(s, t.source));
      V3 pdot = cross(T.w, p) + T.v;  // point_velocity
      S6 jv = bias_in_frame(s, t.source, t.target, t.source);
      V3 bb = cross(T.w, pdot) + cross(jv.w, p) + jv.v;
      for (int r = 0; r < 3; r++) b[r] = bb[r];
      break;
    }
    case T_JOINT: {
      std::memset(J, 0, sizeof(double) * t.dim * nv);
      for (int r = 0; r < t.dim; r++) {
        J[r * nv + m.voff[t.joint] + r] = 1;
        b[r] = 0;
      }
      break;
    }
    case T_MOMENTUM:
    case T_LINMOM: {
      Xf w2c;
      w2c.p = -s.com;  // centroidal frame: world axes at the centre of mass (tasks.jl:209-211)
      momentum_matrix(m, s, w2c, J6.data());
      S6 hb = xforce(w2c, s.momentum_rate_bias);
      int r0 = t.kind == T_LINMOM ? 3 : 0;
      for (int r = 0; r < t.dim; r++) {
        std::memcpy(J + r * nv, J6.data() + (r0 + r) * nv, sizeof(double) * nv);
        int rr = r0 + r;
        b[r] = rr < 3 ? hb.w[rr] : hb.v[rr - 3];
      }
      break;
    }
  }
}

inline void forcebasis(double mu, int N, double* B /*3 x N row-major*/) {  // contacts.jl:16-23
  for (int i = 0; i < N; i++) {
    double th = i * (2 * M_PI / N);
    V3 v{mu * std::cos(th), mu * std::sin(th), 1.0};
    double nn = norm(v);
    for (int r = 0; r < 3; r++) B[r * N + i] = v[r] / nn;
  }
}

struct Workspace {  // one per thread: the reference's controller object is non-reentrant (momentum.jl:1-13)
  State state;
  std::vector<double> vals;  // triplet values in pattern order
  std::vector<double> Px, Ax, qv, l, u, xsol, ysol, des;
  OsqpPort solver;
  bool ready = false;
};

class Controller {
 public:
  const Mechanism* mech;
  int N, floating;  // floating = body index of the floating joint's successor, or -1
  std::vector<double> reg;
  std::vector<Task> tasks;
  std::vector<Contact> contacts;
  std::vector<double> default_weight, default_maxnf;
  StandingParams standing;
  OsqpSettings settings;
  int nvar = 0, ndes = 0;
  bool finalized = false;

  // fixed pattern
  int nrows = 0;
  std::vector<int> trow, tcol;          // A triplets
  std::vector<int> prow, pcol;          // P (upper) triplets
  std::vector<int> Amap, Pmap;          // triplet -> CSC slot
  Csc Apat, Ppat;
  int balance_row0 = -1;

  Controller(const Mechanism* m, int N_, int floating_) : mech(m), N(N_), floating(floating_), reg(m->nv, 0.0) {
    nvar = m->nv;  // vd variables first (momentum.jl:28)
  }
  int add_contact(int body, V3 pos, V3 normal, double mu) {
    Contact c;
    c.body = body;
    c.pos = pos;
    c.normal = normal;
    c.mu = mu;
    c.var0 = nvar;
    nvar += N + 9;
    contacts.push_back(c);
    default_weight.push_back(0.0);  // contacts.jl:50: disabled until the caller sets both
    default_maxnf.push_back(0.0);
    events.push_back({0, (int)contacts.size() - 1});
    return (int)contacts.size() - 1;
  }
  int add_task(Task t) {
    t.dim = task_dim(t.kind, *mech, t.joint);
    t.des_off = ndes;
    ndes += t.dim;
    if (t.mode != M_HARD) {
      t.var0 = nvar;
      nvar += t.dim;
    }
    tasks.push_back(t);
    events.push_back({1, (int)tasks.size() - 1});
    return (int)tasks.size() - 1;
  }

  // initialize!: fix the row order and sparsity (momentum.jl:150-156)
  void finalize() {
    const int nv = mech->nv;
    int row = 0;
    auto dense = [&](int r0, int nr, int c0, int nc) {
      for (int r = 0; r < nr; r++)
        for (int c = 0; c < nc; c++) {
          trow.push_back(r0 + r);
          tcol.push_back(c0 + c);
        }
    };
    auto ident = [&](int r0, int c0, int k) {
      for (int r = 0; r < k; r++) {
        trow.push_back(r0 + r);
        tcol.push_back(c0 + r);
      }
    };
    for (auto& ev : events) {
      if (ev.first == 0) {
        Contact& c = contacts[ev.second];
        int rho0 = c.var0, f0 = rho0 + N, wa0 = f0 + 3, wl0 = wa0 + 3;
        c_row0.push_back(row);
        ident(row, f0, 3);  // f - B rho == 0
        dense(row, 3, rho0, N);
        row += 3;
        ident(row, rho0, N);  // rho >= 0
        row += N;
        ident(row, rho0, N);  // rho <= maxrho
        row += N;
        ident(row, wl0, 3);  // lin(w) - R f == 0
        dense(row, 3, f0, 3);
        row += 3;
        ident(row, wa0, 3);  // ang(w) - hat(p) lin(w) == 0
        dense(row, 3, wl0, 3);
        row += 3;
      } else {
        Task& t = tasks[ev.second];
        t_row0.resize(tasks.size(), 0);
        t_row0[ev.second] = row;
        dense(row, t.dim, 0, nv);
        if (t.mode != M_HARD) ident(row, t.var0, t.dim);
        row += t.dim;
      }
    }
    if (floating >= 0) {  // add_wrench_balance_constraint! (momentum.jl:162-193)
      balance_row0 = row;
      dense(row, 6, 0, nv);
      for (auto& c : contacts) dense(row, 6, c.var0 + N + 3, 6);
      row += 6;
    }
    nrows = row;
    // P pattern: diagonal entries for every variable with a cost, full blocks for matrix weights
    for (int j = 0; j < nv; j++) {
      prow.push_back(j);
      pcol.push_back(j);
    }
    for (auto& c : contacts)
      for (int k = 0; k < 3; k++) {
        prow.push_back(c.var0 + N + k);
        pcol.push_back(c.var0 + N + k);
      }
    for (auto& t : tasks) {
      if (t.mode == M_SCALAR)
        for (int k = 0; k < t.dim; k++) {
          prow.push_back(t.var0 + k);
          pcol.push_back(t.var0 + k);
        }
      if (t.mode == M_MATRIX)
        for (int c = 0; c < t.dim; c++)
          for (int r = 0; r <= c; r++) {
            prow.push_back(t.var0 + r);
            pcol.push_back(t.var0 + c);
          }
    }
    to_csc(nrows, nvar, trow, tcol, Apat, Amap);
    to_csc(nvar, nvar, prow, pcol, Ppat, Pmap);
    finalized = true;
  }

  // StandingController functor body (standing.jl:58-85): fills the desireds of the tasks it owns
  void standing_desireds(const State& s, double* des) const {
    const Mechanism& m = *mech;
    const StandingParams& sp = standing;
    double mass = m.total_mass;
    V3 e = s.com - sp.comref;
    V3 edot = (1.0 / mass) * s.momentum.v;
    V3 cdd = (-sp.com_kp) * e + (-sp.com_kd) * edot;
    const Task& lt = tasks[sp.linmom_task];
    for (int k = 0; k < 3; k++) des[lt.des_off + k] = mass * cdd[k];
    const Task& pt = tasks[sp.pelvis_task];
    const Xf& H = toroot(s, sp.pelvis_body);
    S6 T = xmotion(inv(H), twist_of(s, sp.pelvis_body));
    V3 rv = rot_to_rotvec(H.R);
    V3 wd = (-sp.pelvis_kp) * rv + (-sp.pelvis_kd) * T.w;
    for (int k = 0; k < 3; k++) des[pt.des_off + k] = wd[k];
    for (size_t i = 0; i < sp.joint_tasks.size(); i++) {
      int j = sp.joints[i];
      double qj = s.q[m.qoff[j]], vj = s.v[m.voff[j]];
      des[tasks[sp.joint_tasks[i]].des_off] = -sp.joint_kp[i] * (qj - sp.joint_ref[i]) - sp.joint_kd[i] * vj;
    }
  }

  // Evaluate every parameter for the current state and write the lifted QP values.
  // tweight: per-tick scalar weights, one per task in addtask! order (Parameter weights, momentum.jl:107-110), or null;
  // cgeom: per-tick contact position[3], normal[3], mu per contact (Parameters of contacts.jl:39,53-61), or null
  void assemble(Workspace& w, const double* desired, const double* cweight, const double* cmaxnf,
                const double* tweight = nullptr, const double* cgeom = nullptr) const {
    const Mechanism& m = *mech;
    const State& s = w.state;
    const int nv = m.nv;
    w.vals.clear();
    w.l.assign(nrows, 0.0);
    w.u.assign(nrows, 0.0);
    w.qv.assign(nvar, 0.0);
    std::vector<double> J(6 * nv), b(6), B(3 * N);
    auto push_dense = [&](const double* M, int nr, int nc, int ld, double sgn) {
      for (int r = 0; r < nr; r++)
        for (int c = 0; c < nc; c++) w.vals.push_back(sgn * M[r * ld + c]);
    };
    auto push_ident = [&](int k, double v) {
      for (int r = 0; r < k; r++) w.vals.push_back(v);
    };
    size_t ci = 0;
    for (auto& ev : events) {
      if (ev.first == 0) {
        Contact c = contacts[ev.second];
        if (cgeom) {
          const double* gq = cgeom + 7 * ev.second;
          c.pos = V3{gq[0], gq[1], gq[2]};
          c.normal = V3{gq[3], gq[4], gq[5]};
          c.mu = gq[6];
        }
        int row = c_row0[ci++];
        forcebasis(c.mu, N, B.data());
        double maxrho = cmaxnf[ev.second] / (N * std::sqrt(c.mu * c.mu + 1));  // contacts.jl:57
        Xf zup;
        zup.R = rotation_between_z(c.normal);
        zup.p = c.pos;
        Xf T = toroot(s, c.body) * zup;  // contacts.jl:61
        push_ident(3, 1.0);
        push_dense(B.data(), 3, N, N, -1.0);
        row += 3;
        push_ident(N, 1.0);
        for (int k = 0; k < N; k++) {
          w.l[row + k] = 0;
          w.u[row + k] = OSQP_INFTY;
        }
        row += N;
        push_ident(N, 1.0);
        for (int k = 0; k < N; k++) {
          w.l[row + k] = -OSQP_INFTY;
          w.u[row + k] = maxrho;
        }
        row += N;
        push_ident(3, 1.0);
        push_dense(&T.R.a[0][0], 3, 3, 3, -1.0);
        row += 3;
        push_ident(3, 1.0);
        double hat[9] = {0, -T.p.z, T.p.y, T.p.z, 0, -T.p.x, -T.p.y, T.p.x, 0};
        push_dense(hat, 3, 3, 3, -1.0);
      } else {
        const Task& t = tasks[ev.second];
        int row = t_row0[ev.second];
        task_rows(m, s, t, J.data(), b.data());
        if (t.mode == M_HARD) {
          push_dense(J.data(), t.dim, nv, nv, 1.0);  // J vd + b - des == 0
          for (int r = 0; r < t.dim; r++) w.l[row + r] = w.u[row + r] = desired[t.des_off + r] - b[r];
        } else {
          push_dense(J.data(), t.dim, nv, nv, -1.0);  // e - (J vd + b - des) == 0  (momentum.jl:124)
          push_ident(t.dim, 1.0);
          for (int r = 0; r < t.dim; r++) w.l[row + r] = w.u[row + r] = b[r] - desired[t.des_off + r];
        }
      }
    }
    if (floating >= 0) {
      Xf I;
      momentum_matrix(m, s, I, J.data());
      S6 hb = s.momentum_rate_bias;
      V3 fg = m.total_mass * m.gravity;
      S6 Wg{cross(s.com, fg), fg};
      int o = m.voff[floating];
      std::vector<double> row(nv);
      // rows k: S_k' (A vd + Adv - Wg - sum W_c) == 0
      std::vector<double> SA(6 * nv);
      for (int k = 0; k < 6; k++) {
        const S6& S = s.Sw[o + k];
        for (int c = 0; c < nv; c++) {
          double a = 0;
          for (int r = 0; r < 3; r++) a += S.w[r] * J[r * nv + c] + S.v[r] * J[(3 + r) * nv + c];
          SA[k * nv + c] = a;
        }
      }
      for (int k = 0; k < 6; k++)
        for (int c = 0; c < nv; c++) w.vals.push_back(SA[k * nv + c]);
      for (size_t cc = 0; cc < contacts.size(); cc++)
        for (int k = 0; k < 6; k++) {
          const S6& S = s.Sw[o + k];
          for (int r = 0; r < 3; r++) w.vals.push_back(-S.w[r]);
          for (int r = 0; r < 3; r++) w.vals.push_back(-S.v[r]);
        }
      for (int k = 0; k < 6; k++) w.l[balance_row0 + k] = w.u[balance_row0 + k] = dot(s.Sw[o + k], Wg - hb);
    }
    // reorder the balance block: pattern was emitted as dense(6 x nv) then per contact dense(6 x 6)
    // (values above follow exactly that order)
    w.Ax.assign(Apat.nnz(), 0.0);
    for (size_t k = 0; k < w.vals.size(); k++) w.Ax[Amap[k]] = w.vals[k];
    // objective (momentum.jl:109,115,130 ; contacts.jl:75-79): P = 2 * weights
    std::vector<double> pv;
    for (int j = 0; j < nv; j++) pv.push_back(2 * reg[j]);
    for (size_t cc = 0; cc < contacts.size(); cc++)
      for (int k = 0; k < 3; k++) pv.push_back(2 * cweight[cc]);
    for (size_t ti = 0; ti < tasks.size(); ti++) {
      const Task& t = tasks[ti];
      if (t.mode == M_SCALAR)
        for (int k = 0; k < t.dim; k++) pv.push_back(2 * (tweight ? tweight[ti] : t.weight));
      if (t.mode == M_MATRIX)
        for (int c = 0; c < t.dim; c++)
          for (int r = 0; r <= c; r++) pv.push_back(t.W[r * t.dim + c] + t.W[c * t.dim + r]);
    }
    w.Px.assign(Ppat.nnz(), 0.0);
    for (size_t k = 0; k < pv.size(); k++) w.Px[Pmap[k]] = pv[k];
  }

  struct Result {
    int status, iter, rho_updates;
    double pri_res, dua_res;
  };

  // (controller::MomentumBasedController)(tau, t, x)  momentum.jl:41-81, preceded (if enabled) by standing.jl:58-85
  Result tick(Workspace& w, const double* q, const double* v, const double* desired_in, const double* cweight,
              const double* cmaxnf, double* tau, double* vd, double* wrenches /*ncontacts x 6 world*/,
              double* xlift /*nvar or null*/, const double* tweight = nullptr, const double* cgeom = nullptr) const {
    const Mechanism& m = *mech;
    update_state(m, q, v, w.state);
    w.des.assign(ndes, 0.0);
    if (desired_in) std::copy(desired_in, desired_in + ndes, w.des.begin());
    if (standing.enabled) standing_desireds(w.state, w.des.data());
    if (!cweight) cweight = default_weight.data();
    if (!cmaxnf) cmaxnf = default_maxnf.data();
    assemble(w, w.des.data(), cweight, cmaxnf, tweight, cgeom);
    if (!w.ready) {
      Csc P = Ppat, A = Apat;
      P.x = w.Px;
      A.x = w.Ax;
      w.solver.setup(P, w.qv.data(), A, w.l.data(), w.u.data(), settings);
      w.ready = true;
    } else {
      w.solver.st = settings;
      w.solver.update(w.Px.data(), w.qv.data(), w.Ax.data(), w.l.data(), w.u.data());
    }
    int status = w.solver.solve();
    w.xsol.resize(nvar);
    w.ysol.resize(nrows);
    w.solver.solution(w.xsol.data(), w.ysol.data());
    if (xlift) std::copy(w.xsol.begin(), w.xsol.end(), xlift);
    for (int j = 0; j < m.nv; j++) vd[j] = w.xsol[j];
    std::vector<S6> ext(m.nb);
    for (size_t c = 0; c < contacts.size(); c++) {  // momentum.jl:65-72
      const double* ww = &w.xsol[contacts[c].var0 + N + 3];
      S6 W{{ww[0], ww[1], ww[2]}, {ww[3], ww[4], ww[5]}};
      ext[contacts[c].body] = ext[contacts[c].body] + W;
      if (wrenches) std::copy(ww, ww + 6, wrenches + 6 * c);
    }
    inverse_dynamics(m, w.state, vd, ext.data(), tau);  // momentum.jl:75
    if (floating >= 0)
      for (int k = 0; k < 6; k++) tau[m.voff[floating] + k] = 0;  // momentum.jl:93-97
    const OsqpInfo& inf = w.solver.info;
    return {status, inf.iter, inf.rho_updates, inf.pri_res, inf.dua_res};
  }

  std::vector<std::pair<int, int>> events;  // call order of addcontact!/addtask! (defines variable/row order)
  std::vector<int> c_row0, t_row0;

 private:
  static void to_csc(int m, int n, const std::vector<int>& r, const std::vector<int>& c, Csc& out,
                     std::vector<int>& map) {
    size_t nz = r.size();
    std::vector<int> order(nz);
    for (size_t k = 0; k < nz; k++) order[k] = (int)k;
    std::sort(order.begin(), order.end(), [&](int a, int b) { return c[a] != c[b] ? c[a] < c[b] : r[a] < r[b]; });
    out.m = m;
    out.n = n;
    out.p.assign(n + 1, 0);
    out.i.resize(nz);
    out.x.assign(nz, 0.0);
    map.resize(nz);
    for (size_t t = 0; t < nz; t++) {
      int k = order[t];
      out.i[t] = r[k];
      out.p[c[k] + 1]++;
      map[k] = (int)t;
    }
    for (int j = 0; j < n; j++) out.p[j + 1] += out.p[j];
  }
};

}  // namespace orc
