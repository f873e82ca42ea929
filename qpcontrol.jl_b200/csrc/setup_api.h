// setup_api.h -- the setup-time half of the C ABI (include/qpcontrol_b200.h): records the controller description on
// the host.  Included by api.cu (CUDA backend) and by tests/emu/emu.cpp (CPU emulation of the kernel bodies, test
// infrastructure); each defines `struct Backend` before including this file.
#pragma once
#include <mutex>
#include <cstdio>
#include <string>

#include "../../include/qpcontrol_b200.h"
#include "host_program.h"

static thread_local std::string g_qpc_error;
static int qpc_fail(int code, const std::string& msg) {
  g_qpc_error = msg;
  return code;
}

struct qpc_mechanism {
  qpc::HostMechanism hm;
};
struct qpc_controller {
  qpc::HostController hc;
  qpc::DevProgram prog;
  bool finalized = false;
  Backend be;
};

extern "C" {

int qpc_version(void) { return 100; }
const char* qpc_last_error(void) { return g_qpc_error.c_str(); }

void qpc_default_settings(qpc_settings* s) {
  qpc::Settings d;
  qpc::default_settings(d);
  s->rho = d.rho;
  s->sigma = d.sigma;
  s->alpha = d.alpha;
  s->eps_abs = d.eps_abs;
  s->eps_rel = d.eps_rel;
  s->eps_prim_inf = d.eps_prim_inf;
  s->eps_dual_inf = d.eps_dual_inf;
  s->adaptive_rho_tolerance = d.adaptive_rho_tolerance;
  s->max_iter = d.max_iter;
  s->scaling = d.scaling;
  s->adaptive_rho = d.adaptive_rho;
  s->adaptive_rho_interval = d.adaptive_rho_interval;
  s->check_termination = d.check_termination;
  s->reserved[0] = s->reserved[1] = s->reserved[2] = 0;
}

static void qpc_copy_settings(const qpc_settings* s, qpc::Settings& d) {
  d.rho = s->rho;
  d.sigma = s->sigma;
  d.alpha = s->alpha;
  d.eps_abs = s->eps_abs;
  d.eps_rel = s->eps_rel;
  d.eps_prim_inf = s->eps_prim_inf;
  d.eps_dual_inf = s->eps_dual_inf;
  d.adaptive_rho_tolerance = s->adaptive_rho_tolerance;
  d.max_iter = s->max_iter;
  d.scaling = s->scaling;
  d.adaptive_rho = s->adaptive_rho;
  d.adaptive_rho_interval = s->adaptive_rho_interval;
  d.check_termination = s->check_termination;
}

qpc_mechanism* qpc_mechanism_create(int32_t nb, const int32_t* parent, const int32_t* jtype, const double* axis,
                                    const double* X_R, const double* X_p, const double* mass, const double* com,
                                    const double* inertia_origin, const double gravity[3]) {
  if (nb <= 0 || !parent || !jtype) {
    qpc_fail(QPC_ERR_ARG, "qpc_mechanism_create: bad arguments");
    return nullptr;
  }
  qpc_mechanism* m = new qpc_mechanism();
  qpc::HostMechanism& h = m->hm;
  h.nb = nb;
  for (int b = 0; b < nb; b++) {
    if (parent[b] >= b || parent[b] < -1 || jtype[b] < 0 || jtype[b] > 3) {
      delete m;
      qpc_fail(QPC_ERR_ARG, "qpc_mechanism_create: bodies must be topologically sorted with parent < index");
      return nullptr;
    }
    const int nq = jtype[b] == QPC_QUAT_FLOATING ? 7 : (jtype[b] == QPC_FIXED ? 0 : 1);
    const int nv = jtype[b] == QPC_QUAT_FLOATING ? 6 : (jtype[b] == QPC_FIXED ? 0 : 1);
    h.parent.push_back(parent[b]);
    h.jtype.push_back(jtype[b]);
    h.qoff.push_back(h.nq);
    h.voff.push_back(h.nv);
    h.nqj.push_back(nq);
    h.nvj.push_back(nv);
    h.nq += nq;
    h.nv += nv;
    h.mass.push_back(mass[b]);
    h.total_mass += mass[b];
  }
  h.axis.assign(axis, axis + 3 * nb);
  h.XR.assign(X_R, X_R + 9 * nb);
  h.Xp.assign(X_p, X_p + 3 * nb);
  h.com.assign(com, com + 3 * nb);
  h.inertia_origin.assign(inertia_origin, inertia_origin + 9 * nb);
  for (int i = 0; i < 3; i++) h.gravity[i] = gravity[i];
  return m;
}
void qpc_mechanism_destroy(qpc_mechanism* m) { delete m; }
int qpc_mechanism_dims(const qpc_mechanism* m, int32_t* nb, int32_t* nq, int32_t* nv) {
  if (!m) return qpc_fail(QPC_ERR_ARG, "null mechanism");
  *nb = m->hm.nb;
  *nq = m->hm.nq;
  *nv = m->hm.nv;
  return QPC_OK;
}

qpc_controller* qpc_controller_create(qpc_mechanism* m, int32_t N, int32_t floating_body, const qpc_settings* s) {
  if (!m || N < 1 || N > QPC_MAXN || floating_body >= m->hm.nb) {
    qpc_fail(QPC_ERR_ARG, "qpc_controller_create: bad arguments");
    return nullptr;
  }
  qpc_controller* c = new qpc_controller();
  c->hc.mech = &m->hm;
  c->hc.N = N;
  c->hc.floating = floating_body;
  c->hc.reg.assign(m->hm.nv, 0.0);
  qpc::default_settings(c->hc.settings);
  if (s) qpc_copy_settings(s, c->hc.settings);
  return c;
}

#define QPC_CHECK_OPEN(c)                                        \
  if (!(c)) return qpc_fail(QPC_ERR_ARG, "null controller");     \
  if ((c)->finalized) return qpc_fail(QPC_ERR_STATE, "controller already finalized")

int qpc_add_contact(qpc_controller* c, int32_t body, const double position[3], const double normal[3], double mu) {
  QPC_CHECK_OPEN(c);
  if (body < 0 || body >= c->hc.mech->nb) return qpc_fail(QPC_ERR_ARG, "qpc_add_contact: body out of range");
  if ((int)c->hc.contacts.size() >= QPC_MAXC) return qpc_fail(QPC_ERR_LIMIT, "too many contacts");
  qpc::HostContact h;
  h.body = body;
  for (int i = 0; i < 3; i++) {
    h.pos[i] = position[i];
    h.normal[i] = normal[i];
  }
  h.mu = mu;
  h.weight = 0.0;  // contacts.jl:50 -- disabled until the caller sets both
  h.maxnf = 0.0;
  c->hc.contacts.push_back(h);
  return (int)c->hc.contacts.size() - 1;
}

int qpc_set_contact_params(qpc_controller* c, int32_t contact, double weight, double maxnormalforce) {
  if (!c || contact < 0 || contact >= (int)c->hc.contacts.size()) return qpc_fail(QPC_ERR_ARG, "bad contact index");
  // post-finalize setters mutate the program a concurrent qpc_solve_batch uploads: same mutex as the tick
  std::unique_lock<std::mutex> lock(c->be.mu, std::defer_lock);
  if (c->finalized) lock.lock();
  const bool changed = c->hc.contacts[contact].weight != weight || c->hc.contacts[contact].maxnf != maxnormalforce;
  c->hc.contacts[contact].weight = weight;
  c->hc.contacts[contact].maxnf = maxnormalforce;
  if (c->finalized && changed) {  // re-pushing unchanged values every tick (host mirrors do) costs no program upload
    c->prog.def_cweight[contact] = weight;
    c->prog.def_cmaxnf[contact] = maxnormalforce;
    c->be.dirty = true;
  }
  return QPC_OK;
}

int qpc_add_task(qpc_controller* c, int32_t kind, int32_t source_body, int32_t target_body, int32_t frame_body,
                 const double point[3], int32_t joint, int32_t mode, double weight, const double* W) {
  QPC_CHECK_OPEN(c);
  const qpc::HostMechanism& m = *c->hc.mech;
  if (kind < 0 || kind > 6 || mode < 0 || mode > 2) return qpc_fail(QPC_ERR_ARG, "qpc_add_task: bad kind/mode");
  if (kind <= 3 && (source_body < -1 || source_body >= m.nb || target_body < -1 || target_body >= m.nb ||
                    frame_body < -1 || frame_body >= m.nb))
    return qpc_fail(QPC_ERR_ARG, "qpc_add_task: body out of range");
  if (kind == 4 && (joint < 0 || joint >= m.nb)) return qpc_fail(QPC_ERR_ARG, "qpc_add_task: joint out of range");
  if ((int)c->hc.tasks.size() >= QPC_MAXT) return qpc_fail(QPC_ERR_LIMIT, "too many tasks");
  qpc::HostTask t;
  t.kind = kind;
  t.source = source_body;
  t.target = target_body;
  t.frame = frame_body;
  t.joint = joint;
  t.mode = mode;
  t.weight = weight;
  t.dim = qpc::task_dim(kind, m, joint);
  t.des_off = c->hc.ndes;
  for (int i = 0; i < 3; i++) t.point[i] = point ? point[i] : 0.0;
  if (mode == QPC_MODE_MATRIX_WEIGHT) {
    if (!W) return qpc_fail(QPC_ERR_ARG, "qpc_add_task: matrix weight missing");
    t.W.assign(W, W + t.dim * t.dim);
  }
  t.desired.assign(t.dim, 0.0);
  c->hc.ndes += t.dim;
  c->hc.tasks.push_back(t);
  return (int)c->hc.tasks.size() - 1;
}

int qpc_set_task_desired(qpc_controller* c, int32_t task, const double* desired) {
  if (!c || task < 0 || task >= (int)c->hc.tasks.size()) return qpc_fail(QPC_ERR_ARG, "bad task index");
  if (!desired) return qpc_fail(QPC_ERR_ARG, "qpc_set_task_desired: null desired");
  std::unique_lock<std::mutex> lock(c->be.mu, std::defer_lock);
  if (c->finalized) lock.lock();
  qpc::HostTask& t = c->hc.tasks[task];
  bool changed = (int)t.desired.size() != t.dim;
  for (int i = 0; i < t.dim && !changed; i++) changed = t.desired[i] != desired[i];
  t.desired.assign(desired, desired + t.dim);
  if (c->finalized && changed) {
    for (int i = 0; i < t.dim; i++) c->prog.def_desired[t.des_off + i] = desired[i];
    c->be.dirty = true;
  }
  return QPC_OK;
}

int qpc_regularize(qpc_controller* c, int32_t joint, double weight) {
  QPC_CHECK_OPEN(c);
  const qpc::HostMechanism& m = *c->hc.mech;
  if (joint < 0 || joint >= m.nb) return qpc_fail(QPC_ERR_ARG, "qpc_regularize: joint out of range");
  for (int k = m.voff[joint]; k < m.voff[joint] + m.nvj[joint]; k++) c->hc.reg[k] += weight;
  return QPC_OK;
}

int qpc_standing_setup(qpc_controller* c, int32_t linmom_task, int32_t pelvis_task, int32_t pelvis_body,
                       int32_t njoints, const int32_t* joint_tasks, const int32_t* joints, const double* kp,
                       const double* kd, const double* qref, double com_kp, double com_kd, double pelvis_kp,
                       double pelvis_kd, const double comref[3]) {
  QPC_CHECK_OPEN(c);
  const int nt = (int)c->hc.tasks.size();
  if (linmom_task < 0 || linmom_task >= nt || pelvis_task < 0 || pelvis_task >= nt || njoints > QPC_MAXV)
    return qpc_fail(QPC_ERR_ARG, "qpc_standing_setup: bad task indices");
  if (c->hc.tasks[linmom_task].kind != QPC_TASK_LINEAR_MOMENTUM_RATE || c->hc.tasks[pelvis_task].kind != QPC_TASK_ANGULAR)
    return qpc_fail(QPC_ERR_ARG, "qpc_standing_setup: task kinds do not match the standing controller");
  {
    // every index the program compiler later dereferences is checked here (standing.jl:46-49: one 1-dof
    // JointAccelerationTask per position-controlled joint)
    const qpc::HostMechanism& m = *c->hc.mech;
    if (njoints < 0 || (njoints > 0 && (!joint_tasks || !joints || !kp || !kd || !qref)) || !comref)
      return qpc_fail(QPC_ERR_ARG, "qpc_standing_setup: bad joint arrays");
    if (pelvis_body < 0 || pelvis_body >= m.nb) return qpc_fail(QPC_ERR_ARG, "qpc_standing_setup: pelvis body out of range");
    for (int i = 0; i < njoints; i++) {
      if (joint_tasks[i] < 0 || joint_tasks[i] >= nt || joints[i] < 0 || joints[i] >= m.nb)
        return qpc_fail(QPC_ERR_ARG, "qpc_standing_setup: joint / joint task index out of range");
      const qpc::HostTask& jt = c->hc.tasks[joint_tasks[i]];
      if (jt.kind != QPC_TASK_JOINT || jt.dim != 1 || jt.joint != joints[i] || m.nvj[joints[i]] != 1)
        return qpc_fail(QPC_ERR_ARG, "qpc_standing_setup: joint tasks must be 1-dof JointAccelerationTasks of the listed joints");
    }
  }
  qpc::HostStanding& s = c->hc.standing;
  s.enabled = true;
  s.linmom_task = linmom_task;
  s.pelvis_task = pelvis_task;
  s.pelvis_body = pelvis_body;
  s.joint_tasks.assign(joint_tasks, joint_tasks + njoints);
  s.joints.assign(joints, joints + njoints);
  s.kp.assign(kp, kp + njoints);
  s.kd.assign(kd, kd + njoints);
  s.ref.assign(qref, qref + njoints);
  s.com_kp = com_kp;
  s.com_kd = com_kd;
  s.pelvis_kp = pelvis_kp;
  s.pelvis_kd = pelvis_kd;
  for (int i = 0; i < 3; i++) s.comref[i] = comref[i];
  return QPC_OK;
}

// ---- SE3PDController on the device (se3pdcontroller.jl:1-18) ------------------------------------------------------------
static const char* fill_trajectory(qpc::DevTraj& tr, int32_t n, const qpc_interp_piece* pcs, int32_t piecewise,
                                   double break_end, bool rotation) {
  if (n < 1 || n > QPC_MAXSEG || !pcs) return "between 1 and QPC_MAX_PIECES pieces per trajectory";
  if (!piecewise && n != 1) return "a trajectory that is not piecewise has exactly one piece";
  std::memset(&tr, 0, sizeof(tr));
  tr.nseg = n;
  tr.piecewise = piecewise ? 1 : 0;
  tr.brk_end = break_end;
  for (int i = 0; i < n; i++) {
    const qpc_interp_piece& s = pcs[i];
    qpc::DevInterp& d = tr.seg[i];
    if (!(s.xf > s.x0) || !std::isfinite(s.x0) || !std::isfinite(s.xf)) return "piece with xf <= x0";
    if (s.ncoeffs < 0 || s.ncoeffs > 6) return "interpolator polynomials have at most 6 coefficients";
    if (piecewise && ((i > 0 && !(s.break_start >= pcs[i - 1].break_start)) || !(break_end >= s.break_start)))
      return "Piecewise breaks must be sorted";
    if (rotation) {
      const double nq = s.y0[0] * s.y0[0] + s.y0[1] * s.y0[1] + s.y0[2] * s.y0[2] + s.y0[3] * s.y0[3];
      const double na = s.dy[0] * s.dy[0] + s.dy[1] * s.dy[1] + s.dy[2] * s.dy[2];
      if (std::fabs(nq - 1.0) > 1e-9 || std::fabs(na - 1.0) > 1e-9) return "rotation pieces need a unit quaternion and a unit axis";
    }
    d.brk = piecewise ? s.break_start : 0.0;
    d.x0 = s.x0;
    d.xf = s.xf;
    for (int k = 0; k < 4; k++) d.y0[k] = s.y0[k];
    for (int k = 0; k < 3; k++) d.dy[k] = s.dy[k];
    d.angle = s.angle;
    for (int k = 0; k < 6; k++) d.c[k] = k < s.ncoeffs ? s.coeffs[k] : 0.0;
    d.nc = s.ncoeffs;
  }
  return nullptr;
}

int qpc_add_se3pd(qpc_controller* c, int32_t task, int32_t base_body, int32_t body, const double gains[36],
                  int32_t n_angular, const qpc_interp_piece* angular, int32_t angular_piecewise, double angular_break_end,
                  int32_t n_linear, const qpc_interp_piece* linear, int32_t linear_piecewise, double linear_break_end) {
  QPC_CHECK_OPEN(c);
  const qpc::HostMechanism& m = *c->hc.mech;
  if (task < 0 || task >= (int)c->hc.tasks.size() || c->hc.tasks[task].kind != QPC_TASK_SPATIAL)
    return qpc_fail(QPC_ERR_ARG, "qpc_add_se3pd: `task` must be a SpatialAccelerationTask added before");
  if (body < 0 || body >= m.nb || base_body < -1 || base_body >= m.nb) return qpc_fail(QPC_ERR_ARG, "qpc_add_se3pd: body out of range");
  if (!gains) return qpc_fail(QPC_ERR_ARG, "qpc_add_se3pd: null gains");
  if ((int)c->hc.se3.size() >= QPC_MAXSE3) return qpc_fail(QPC_ERR_LIMIT, "too many SE3PDControllers");
  for (auto& o : c->hc.se3)
    if (o.task == task) return qpc_fail(QPC_ERR_ARG, "qpc_add_se3pd: the task is already driven by an SE3PDController");
  qpc::HostSE3PD h;
  h.task = task;
  h.base = base_body;
  h.body = body;
  std::memcpy(h.K, gains, sizeof(h.K));
  if (const char* e = fill_trajectory(h.ang, n_angular, angular, angular_piecewise, angular_break_end, true))
    return qpc_fail(QPC_ERR_ARG, std::string("qpc_add_se3pd: angular: ") + e);
  if (const char* e = fill_trajectory(h.lin, n_linear, linear, linear_piecewise, linear_break_end, false))
    return qpc_fail(QPC_ERR_ARG, std::string("qpc_add_se3pd: linear: ") + e);
  c->hc.se3.push_back(h);
  return (int)c->hc.se3.size() - 1;
}

int qpc_se3pd_update(qpc_controller* c, int32_t se3pd, const double* gains,
                     int32_t n_angular, const qpc_interp_piece* angular, int32_t angular_piecewise, double angular_break_end,
                     int32_t n_linear, const qpc_interp_piece* linear, int32_t linear_piecewise, double linear_break_end) {
  if (!c || se3pd < 0 || se3pd >= (int)c->hc.se3.size()) return qpc_fail(QPC_ERR_ARG, "qpc_se3pd_update: bad index");
  std::unique_lock<std::mutex> lock(c->be.mu, std::defer_lock);
  if (c->finalized) lock.lock();
  qpc::HostSE3PD h = c->hc.se3[se3pd];
  if (gains) std::memcpy(h.K, gains, sizeof(h.K));
  if (angular)
    if (const char* e = fill_trajectory(h.ang, n_angular, angular, angular_piecewise, angular_break_end, true))
      return qpc_fail(QPC_ERR_ARG, std::string("qpc_se3pd_update: angular: ") + e);
  if (linear)
    if (const char* e = fill_trajectory(h.lin, n_linear, linear, linear_piecewise, linear_break_end, false))
      return qpc_fail(QPC_ERR_ARG, std::string("qpc_se3pd_update: linear: ") + e);
  const qpc::HostSE3PD& old = c->hc.se3[se3pd];
  const bool changed = std::memcmp(old.K, h.K, sizeof(h.K)) || std::memcmp(&old.ang, &h.ang, sizeof(h.ang)) ||
                       std::memcmp(&old.lin, &h.lin, sizeof(h.lin));
  c->hc.se3[se3pd] = h;
  if (c->finalized && changed) {
    qpc::DevSE3PD& d = c->prog.se3[se3pd];
    std::memcpy(d.K, h.K, sizeof(d.K));
    d.ang = h.ang;
    d.lin = h.lin;
    c->be.dirty = true;
  }
  return QPC_OK;
}

int qpc_set_settings(qpc_controller* c, const qpc_settings* s) {
  if (!c || !s) return qpc_fail(QPC_ERR_ARG, "null argument");
  std::unique_lock<std::mutex> lock(c->be.mu, std::defer_lock);
  if (c->finalized) lock.lock();
  qpc_copy_settings(s, c->hc.settings);
  if (c->finalized) {
    c->prog.settings = c->hc.settings;
    c->be.dirty = true;
  }
  return QPC_OK;
}

int qpc_controller_dims(const qpc_controller* c, int32_t* nq, int32_t* nv, int32_t* ndes, int32_t* ncontacts,
                        int32_t* n, int32_t* mg, int32_t* nbox) {
  if (!c || !c->finalized) return qpc_fail(QPC_ERR_STATE, "controller not finalized");
  *nq = c->prog.nq;
  *nv = c->prog.nv;
  *ndes = c->prog.ndes;
  *ncontacts = c->prog.ncontacts;
  *n = c->prog.n;
  *mg = c->prog.mg;
  *nbox = c->prog.nbx;
  return QPC_OK;
}

// doubles of one row of qpc_batch_in.task_weight_matrix: sum of dim^2 over the matrix-weighted tasks (addtask! order)
int qpc_controller_weight_matrix_doubles(const qpc_controller* c) { return (c && c->finalized) ? c->prog.nwmat : 0; }

}  // extern "C"
