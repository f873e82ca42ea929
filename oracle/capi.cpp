// ORACLE -- TEST INFRASTRUCTURE ONLY.  C entry points for the ctypes wrapper in oracle/oracle.py.
// Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may load this library.
#include <omp.h>

#include <chrono>
#include <cstdio>

#include "controller.hpp"

using namespace orc;

extern "C" {

void* orc_mechanism_create(int nb, const int* parent, const int* jtype, const double* axis, const double* XR,
                           const double* Xp, const double* mass, const double* com, const double* inertia_origin,
                           const double* gravity) {
  Mechanism* m = new Mechanism();
  m->nb = nb;
  for (int b = 0; b < nb; b++) {
    m->parent.push_back(parent[b]);
    m->jtype.push_back(jtype[b]);
    int nq = jtype[b] == QUAT_FLOATING ? 7 : (jtype[b] == FIXED ? 0 : 1);
    int nv = jtype[b] == QUAT_FLOATING ? 6 : (jtype[b] == FIXED ? 0 : 1);
    m->qoff.push_back(m->nq);
    m->voff.push_back(m->nv);
    m->nqj.push_back(nq);
    m->nvj.push_back(nv);
    for (int k = 0; k < nv; k++) m->vbody.push_back(b);
    m->nq += nq;
    m->nv += nv;
    m->axis.push_back({axis[3 * b], axis[3 * b + 1], axis[3 * b + 2]});
    Xf X;
    for (int i = 0; i < 3; i++)
      for (int j = 0; j < 3; j++) X.R.a[i][j] = XR[9 * b + 3 * i + j];
    X.p = {Xp[3 * b], Xp[3 * b + 1], Xp[3 * b + 2]};
    m->Xtree.push_back(X);
    SI I;
    for (int i = 0; i < 3; i++)
      for (int j = 0; j < 3; j++) I.J.a[i][j] = inertia_origin[9 * b + 3 * i + j];
    I.m = mass[b];
    I.c = {mass[b] * com[3 * b], mass[b] * com[3 * b + 1], mass[b] * com[3 * b + 2]};
    m->inertia.push_back(I);
    m->total_mass += mass[b];
  }
  m->gravity = {gravity[0], gravity[1], gravity[2]};
  return m;
}
void orc_mechanism_destroy(void* m) { delete (Mechanism*)m; }

// ---- single-state kinematic queries (used by the invariant tests) ------------------------------------------------
struct QState {
  const Mechanism* m;
  State s;
};
void* orc_state_create(void* mech) { return new QState{(const Mechanism*)mech, State{}}; }
void orc_state_destroy(void* s) { delete (QState*)s; }
void orc_state_set(void* sp, const double* q, const double* v) {
  QState* s = (QState*)sp;
  update_state(*s->m, q, v, s->s);
}
static void put6(S6 a, double* o) {
  for (int k = 0; k < 3; k++) {
    o[k] = a.w[k];
    o[3 + k] = a.v[k];
  }
}
void orc_com(void* sp, double* out) {
  QState* s = (QState*)sp;
  for (int k = 0; k < 3; k++) out[k] = s->s.com[k];
}
void orc_momentum(void* sp, double* out) { put6(((QState*)sp)->s.momentum, out); }
void orc_momentum_rate_bias(void* sp, double* out) { put6(((QState*)sp)->s.momentum_rate_bias, out); }
void orc_transform_to_root(void* sp, int body, double* R, double* p) {
  QState* s = (QState*)sp;
  const Xf& X = toroot(s->s, body);
  for (int i = 0; i < 3; i++) {
    for (int j = 0; j < 3; j++) R[3 * i + j] = X.R.a[i][j];
    p[i] = X.p[i];
  }
}
void orc_twist(void* sp, int body, double* out) { put6(twist_of(((QState*)sp)->s, body), out); }
void orc_bias_acceleration(void* sp, int body, double* out) { put6(bias_of(((QState*)sp)->s, body), out); }
void orc_geometric_jacobian(void* sp, int source, int target, int frame, double* J) {
  QState* s = (QState*)sp;
  geometric_jacobian(*s->m, s->s, source, target, frame, J);
}
void orc_bias_in_frame(void* sp, int source, int target, int frame, double* out) {
  put6(bias_in_frame(((QState*)sp)->s, source, target, frame), out);
}
void orc_momentum_matrix(void* sp, int centroidal, double* A) {
  QState* s = (QState*)sp;
  Xf X;
  if (centroidal) X.p = -s->s.com;
  momentum_matrix(*s->m, s->s, X, A);
}
void orc_mass_matrix(void* sp, double* M) {
  QState* s = (QState*)sp;
  mass_matrix(*s->m, s->s, M);
}
void orc_inverse_dynamics(void* sp, const double* vd, const double* ext /*[nb][6] world or null*/, double* tau) {
  QState* s = (QState*)sp;
  std::vector<S6> e;
  if (ext) {
    e.resize(s->m->nb);
    for (int b = 0; b < s->m->nb; b++)
      e[b] = S6{{ext[6 * b], ext[6 * b + 1], ext[6 * b + 2]}, {ext[6 * b + 3], ext[6 * b + 4], ext[6 * b + 5]}};
  }
  inverse_dynamics(*s->m, s->s, vd, ext ? e.data() : nullptr, tau);
}

// ---- controller ---------------------------------------------------------------------------------------------------
struct Ctl {
  Controller c;
  std::vector<Workspace> ws;
  Ctl(const Mechanism* m, int N, int fl) : c(m, N, fl) {}
};

void* orc_controller_create(void* mech, int N, int floating_body) {
  return new Ctl((const Mechanism*)mech, N, floating_body);
}
void orc_controller_destroy(void* c) { delete (Ctl*)c; }
int orc_add_contact(void* cp, int body, const double* pos, const double* normal, double mu) {
  return ((Ctl*)cp)->c.add_contact(body, {pos[0], pos[1], pos[2]}, {normal[0], normal[1], normal[2]}, mu);
}
void orc_set_contact_defaults(void* cp, int idx, double weight, double maxnf) {
  Ctl* c = (Ctl*)cp;
  c->c.default_weight[idx] = weight;
  c->c.default_maxnf[idx] = maxnf;
}
int orc_add_task(void* cp, int kind, int source, int target, int frame, const double* point, int joint, int mode,
                 double weight, const double* W) {
  Ctl* c = (Ctl*)cp;
  Task t;
  t.kind = kind;
  t.source = source;
  t.target = target;
  t.frame = frame;
  t.joint = joint;
  t.mode = mode;
  t.weight = weight;
  if (point) t.point = {point[0], point[1], point[2]};
  int dim = task_dim(kind, *c->c.mech, joint);
  if (mode == M_MATRIX) t.W.assign(W, W + dim * dim);
  return c->c.add_task(t);
}
void orc_regularize(void* cp, int joint, double w) {
  Ctl* c = (Ctl*)cp;
  const Mechanism& m = *c->c.mech;
  for (int k = m.voff[joint]; k < m.voff[joint] + m.nvj[joint]; k++) c->c.reg[k] += w;
}
void orc_set_standing(void* cp, int linmom_task, int pelvis_task, int pelvis_body, int nj, const int* joint_tasks,
                      const int* joints, const double* kp, const double* kd, const double* ref, double com_kp,
                      double com_kd, double pelvis_kp, double pelvis_kd, const double* comref) {
  StandingParams& s = ((Ctl*)cp)->c.standing;
  s.enabled = true;
  s.linmom_task = linmom_task;
  s.pelvis_task = pelvis_task;
  s.pelvis_body = pelvis_body;
  s.joint_tasks.assign(joint_tasks, joint_tasks + nj);
  s.joints.assign(joints, joints + nj);
  s.joint_kp.assign(kp, kp + nj);
  s.joint_kd.assign(kd, kd + nj);
  s.joint_ref.assign(ref, ref + nj);
  s.com_kp = com_kp;
  s.com_kd = com_kd;
  s.pelvis_kp = pelvis_kp;
  s.pelvis_kd = pelvis_kd;
  s.comref = {comref[0], comref[1], comref[2]};
}
void orc_set_settings(void* cp, double eps_abs, double eps_rel, int max_iter, int adaptive_rho_interval,
                      int check_termination, int scaling, int warm_start, double rho, double sigma, double alpha) {
  OsqpSettings& s = ((Ctl*)cp)->c.settings;
  s.eps_abs = eps_abs;
  s.eps_rel = eps_rel;
  s.max_iter = max_iter;
  s.adaptive_rho_interval = adaptive_rho_interval;
  s.check_termination = check_termination;
  s.scaling = scaling;
  s.warm_start = warm_start;
  s.rho = rho;
  s.sigma = sigma;
  s.alpha = alpha;
}
void orc_finalize(void* cp) {
  Ctl* c = (Ctl*)cp;
  if (!c->c.finalized) c->c.finalize();
}
void orc_dims(void* cp, int* nvar, int* nrows, int* ndes, int* ncontacts) {
  Ctl* c = (Ctl*)cp;
  *nvar = c->c.nvar;
  *nrows = c->c.nrows;
  *ndes = c->c.ndes;
  *ncontacts = (int)c->c.contacts.size();
}
// rows of one task at a state: error = J vd + b - desired
void orc_task_rows(void* cp, void* sp, int task, double* J, double* b) {
  Ctl* c = (Ctl*)cp;
  QState* s = (QState*)sp;
  task_rows(*c->c.mech, s->s, c->c.tasks[task], J, b);
}
// desireds the standing controller would set for this state
void orc_standing_desireds(void* cp, void* sp, double* des) {
  Ctl* c = (Ctl*)cp;
  c->c.standing_desireds(((QState*)sp)->s, des);
}
// dense copy of the lifted QP for one state (tests: KKT verification in numpy)
void orc_lifted_qp(void* cp, const double* q, const double* v, const double* desired, const double* cweight,
                   const double* cmaxnf, double* Pd /*nvar^2*/, double* qd, double* Ad /*nrows x nvar*/, double* l,
                   double* u) {
  Ctl* c = (Ctl*)cp;
  Controller& k = c->c;
  Workspace w;
  update_state(*k.mech, q, v, w.state);
  w.des.assign(k.ndes, 0.0);
  if (desired) std::copy(desired, desired + k.ndes, w.des.begin());
  if (k.standing.enabled) k.standing_desireds(w.state, w.des.data());
  k.assemble(w, w.des.data(), cweight ? cweight : k.default_weight.data(), cmaxnf ? cmaxnf : k.default_maxnf.data());
  std::fill(Pd, Pd + (size_t)k.nvar * k.nvar, 0.0);
  std::fill(Ad, Ad + (size_t)k.nrows * k.nvar, 0.0);
  for (int j = 0; j < k.nvar; j++) {
    for (int p = k.Ppat.p[j]; p < k.Ppat.p[j + 1]; p++) {
      Pd[(size_t)k.Ppat.i[p] * k.nvar + j] += w.Px[p];
      if (k.Ppat.i[p] != j) Pd[(size_t)j * k.nvar + k.Ppat.i[p]] += w.Px[p];
    }
    for (int p = k.Apat.p[j]; p < k.Apat.p[j + 1]; p++) Ad[(size_t)k.Apat.i[p] * k.nvar + j] += w.Ax[p];
    qd[j] = w.qv[j];
  }
  for (int i = 0; i < k.nrows; i++) {
    l[i] = w.l[i];
    u[i] = w.u[i];
  }
}

// Batched control tick: `Threads.@threads over instances`, one private workspace per thread.
// Strides of 0 for desired/cweight/cmaxnf broadcast one row to every instance; null pointers use the defaults.
// Returns wall seconds spent in the parallel region.
// per-tick Parameters beyond desireds / contact weight / maxnormalforce: task weights [B][ntasks] and contact geometry
// [B][ncontacts][7] (position, normal, mu), both optional; used by orc_solve_batch when set (test infrastructure)
static const double* g_tweight = nullptr;
static const double* g_cgeom = nullptr;
static int64_t g_tweight_stride = 0, g_cgeom_stride = 0;
void orc_set_tick_parameters(const double* tweight, int64_t tweight_stride, const double* cgeom, int64_t cgeom_stride) {
  g_tweight = tweight;
  g_tweight_stride = tweight_stride;
  g_cgeom = cgeom;
  g_cgeom_stride = cgeom_stride;
}
double orc_solve_batch(void* cp, int64_t B, const double* q, const double* v, const double* desired,
                       int64_t desired_stride, const double* cweight, const double* cmaxnf, int64_t contact_stride,
                       double* tau, double* vd, double* wrenches, int32_t* status, int32_t* iters, double* res,
                       int32_t* rho_updates, double* xlift, int nthreads) {
  Ctl* c = (Ctl*)cp;
  Controller& k = c->c;
  if (!k.finalized) k.finalize();
  if (nthreads <= 0) nthreads = omp_get_max_threads();
  if ((int)c->ws.size() < nthreads) c->ws.resize(nthreads);
  const Mechanism& m = *k.mech;
  const int nc = (int)k.contacts.size();
  auto t0 = std::chrono::steady_clock::now();
#pragma omp parallel for schedule(dynamic, 4) num_threads(nthreads)
  for (int64_t i = 0; i < B; i++) {
    Workspace& w = c->ws[omp_get_thread_num()];
    Controller::Result r = k.tick(
        w, q + i * m.nq, v + i * m.nv, desired ? desired + i * desired_stride : nullptr,
        cweight ? cweight + i * contact_stride : nullptr, cmaxnf ? cmaxnf + i * contact_stride : nullptr,
        tau + i * m.nv, vd + i * m.nv, wrenches ? wrenches + i * nc * 6 : nullptr, xlift ? xlift + i * k.nvar : nullptr,
        g_tweight ? g_tweight + i * g_tweight_stride : nullptr, g_cgeom ? g_cgeom + i * g_cgeom_stride : nullptr);
    status[i] = r.status;
    if (iters) iters[i] = r.iter;
    if (rho_updates) rho_updates[i] = r.rho_updates;
    if (res) {
      res[2 * i] = r.pri_res;
      res[2 * i + 1] = r.dua_res;
    }
  }
  return std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
}
void orc_reset_workspaces(void* cp) { ((Ctl*)cp)->ws.clear(); }

// ---- raw QPs: dense row-major P (n x n), A (m x n) per instance ------------------------------
double orc_solve_dense_qp_batch(int64_t B, int n, int m, const double* P, const double* q, const double* A,
                                const double* l, const double* u, double eps_abs, double eps_rel, int max_iter,
                                double* x, double* y, int32_t* status, int32_t* iters, double* res, int nthreads) {
  if (nthreads <= 0) nthreads = omp_get_max_threads();
  auto t0 = std::chrono::steady_clock::now();
#pragma omp parallel for schedule(dynamic, 1) num_threads(nthreads)
  for (int64_t b = 0; b < B; b++) {
    Csc Pc, Ac;
    Pc.m = Pc.n = n;
    Pc.p.assign(n + 1, 0);
    Ac.m = m;
    Ac.n = n;
    Ac.p.assign(n + 1, 0);
    const double* Pb = P + (size_t)b * n * n;
    const double* Ab = A + (size_t)b * m * n;
    for (int j = 0; j < n; j++) {
      for (int i = 0; i <= j; i++) {
        Pc.i.push_back(i);
        Pc.x.push_back(Pb[(size_t)i * n + j]);
      }
      Pc.p[j + 1] = (int)Pc.i.size();
      for (int i = 0; i < m; i++) {
        Ac.i.push_back(i);
        Ac.x.push_back(Ab[(size_t)i * n + j]);
      }
      Ac.p[j + 1] = (int)Ac.i.size();
    }
    OsqpSettings s;
    s.eps_abs = eps_abs;
    s.eps_rel = eps_rel;
    s.max_iter = max_iter;
    s.warm_start = 0;
    OsqpPort solver;
    solver.setup(Pc, q + (size_t)b * n, Ac, l + (size_t)b * m, u + (size_t)b * m, s);
    status[b] = solver.solve();
    solver.solution(x + (size_t)b * n, y ? y + (size_t)b * m : nullptr);
    if (iters) iters[b] = solver.info.iter;
    if (res) {
      res[2 * b] = solver.info.pri_res;
      res[2 * b + 1] = solver.info.dua_res;
    }
  }
  return std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
}

int orc_max_threads() { return omp_get_max_threads(); }

}  // extern "C"
