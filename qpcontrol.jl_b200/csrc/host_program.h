// host_program.h -- host-side recorder of the setup-time API and its compilation into the flat device tables.
// Plain C++ (no CUDA): shared by api.cu and the CPU emulation harness under tests/emu.
#pragma once
#include <cmath>
#include <cstring>
#include <string>
#include <vector>

#include "qpc_common.h"
#include "qpc_program.h"

namespace qpc {

struct HostMechanism {
  int nb = 0, nq = 0, nv = 0;
  std::vector<int> parent, jtype, qoff, voff, nvj, nqj;
  std::vector<double> axis, XR, Xp, mass, com, inertia_origin;
  double gravity[3] = {0, 0, -9.81};
  double total_mass = 0;
};

struct HostTask {
  int kind, source, target, frame, joint, mode, dim, des_off;
  double point[3];
  double weight;
  std::vector<double> W, desired;
};
struct HostContact {
  int body;
  double pos[3], normal[3], mu, weight, maxnf;
};
struct HostStanding {
  bool enabled = false;
  int linmom_task = -1, pelvis_task = -1, pelvis_body = -1;
  std::vector<int> joint_tasks, joints;
  std::vector<double> kp, kd, ref;
  double com_kp = 0, com_kd = 0, pelvis_kp = 0, pelvis_kd = 0, comref[3] = {0, 0, 0};
};

struct HostSE3PD {
  int task, base, body;
  double K[36];
  DevTraj ang, lin;
};

struct HostController {
  const HostMechanism* mech = nullptr;
  int N = 4, floating = -1;
  std::vector<HostTask> tasks;
  std::vector<HostContact> contacts;
  std::vector<double> reg;
  HostStanding standing;
  std::vector<HostSE3PD> se3;
  Settings settings;
  int ndes = 0;
};

inline void default_settings(Settings& s) {
  s.rho = 0.1;
  s.sigma = 1e-6;
  s.alpha = 1.6;
  s.eps_abs = 1e-3;
  s.eps_rel = 1e-3;
  s.eps_prim_inf = 1e-4;
  s.eps_dual_inf = 1e-4;
  s.adaptive_rho_tolerance = 5.0;
  s.max_iter = 4000;
  s.scaling = 10;
  s.adaptive_rho = 1;
  s.adaptive_rho_interval = 25;
  s.check_termination = 25;
}

inline int task_dim(int kind, const HostMechanism& m, int joint) {
  if (kind == 0 || kind == 5) return 6;
  if (kind == 4) return m.nvj[joint];
  return 3;
}

// Rotations.rotation_between((0,0,1), v)  (reference src/contacts.jl:11)
inline void rotation_between_z(const double* v, double* R) {
  double n = std::sqrt(v[0] * v[0] + v[1] * v[1] + v[2] * v[2]);
  double t[3] = {v[0] / n, v[1] / n, v[2] / n};
  double ax[3] = {-t[1], t[0], 0.0};  // (0,0,1) x t
  double s = std::sqrt(ax[0] * ax[0] + ax[1] * ax[1]), c = t[2];
  for (int i = 0; i < 9; i++) R[i] = (i % 4 == 0) ? 1.0 : 0.0;
  if (s < 1e-14) {
    if (c < 0) {
      R[4] = -1;
      R[8] = -1;
    }
    return;
  }
  V3 k = mk3(ax[0] / s, ax[1] / s, 0.0);
  axis_angle_to_rot(k, std::atan2(s, c), R);
}

// returns "" on success, otherwise a description of the violated limit / inconsistency
inline std::string compile_program(const HostController& hc, DevProgram& p) {
  const HostMechanism& m = *hc.mech;
  std::memset(&p, 0, sizeof(p));
  if (m.nb > QPC_MAXB) return "too many bodies";
  if (m.nv > QPC_MAXV || m.nq > QPC_MAXQ) return "too many degrees of freedom";
  if ((int)hc.tasks.size() > QPC_MAXT) return "too many tasks";
  if ((int)hc.contacts.size() > QPC_MAXC) return "too many contacts";
  if (hc.N > QPC_MAXN || hc.N < 1) return "unsupported number of friction-cone generators";
  if (hc.ndes > QPC_MAXDES) return "too many desired values";
  p.nb = m.nb;
  p.nq = m.nq;
  p.nv = m.nv;
  std::vector<int> depth(m.nb, 0);
  int maxd = 0;
  for (int b = 0; b < m.nb; b++) {
    p.parent[b] = m.parent[b];
    p.jtype[b] = m.jtype[b];
    p.qoff[b] = m.qoff[b];
    p.voff[b] = m.voff[b];
    p.nvj[b] = m.nvj[b];
    for (int k = 0; k < m.nvj[b]; k++) p.vbody[m.voff[b] + k] = b;
    for (int k = 0; k < 3; k++) {
      p.axis[3 * b + k] = m.axis[3 * b + k];
      p.Xp[3 * b + k] = m.Xp[3 * b + k];
    }
    for (int k = 0; k < 9; k++) p.XR[9 * b + k] = m.XR[9 * b + k];
    const double* J = &m.inertia_origin[9 * b];
    double* I = &p.inertia[10 * b];
    I[0] = J[0];
    I[1] = 0.5 * (J[1] + J[3]);
    I[2] = 0.5 * (J[2] + J[6]);
    I[3] = J[4];
    I[4] = 0.5 * (J[5] + J[7]);
    I[5] = J[8];
    for (int k = 0; k < 3; k++) I[6 + k] = m.mass[b] * m.com[3 * b + k];
    I[9] = m.mass[b];
    depth[b] = m.parent[b] < 0 ? 0 : depth[m.parent[b]] + 1;
    if (depth[b] > maxd) maxd = depth[b];
  }
  p.nlevels = m.nb ? maxd + 1 : 0;
  int k = 0;
  for (int d = 0; d <= maxd; d++) {
    p.level_ptr[d] = k;
    for (int b = 0; b < m.nb; b++)
      if (depth[b] == d) p.level_body[k++] = b;
  }
  p.level_ptr[maxd + 1] = k;
  {  // ancestor chains (root first, the body itself last) and descendant lists (deepest first)
    int ka = 0, kd = 0;
    for (int b = 0; b < m.nb; b++) {
      p.anc_ptr[b] = ka;
      if (ka + depth[b] + 1 > QPC_MAXANC) return "kinematic tree too deep for the ancestor tables";
      int chain[QPC_MAXB], len = 0;
      for (int a = b; a >= 0; a = m.parent[a]) chain[len++] = a;
      for (int i = len - 1; i >= 0; i--) p.anc_idx[ka++] = chain[i];
      p.desc_ptr[b] = kd;
      for (int d = maxd; d > depth[b]; d--)
        for (int c = 0; c < m.nb; c++) {
          if (depth[c] != d) continue;
          bool below = false;
          for (int a = m.parent[c]; a >= 0; a = m.parent[a]) below = below || a == b;
          if (below) {
            if (kd >= QPC_MAXANC) return "kinematic tree too deep for the descendant tables";
            p.desc_idx[kd++] = c;
          }
        }
    }
    p.anc_ptr[m.nb] = ka;
    p.desc_ptr[m.nb] = kd;
  }
  k = 0;
  for (int b = 0; b < m.nb; b++) {
    p.child_ptr[b] = k;
    for (int c = 0; c < m.nb; c++)
      if (m.parent[c] == b) p.child_idx[k++] = c;
  }
  p.child_ptr[m.nb] = k;
  for (int i = 0; i < 3; i++) p.gravity[i] = m.gravity[i];
  p.total_mass = m.total_mass;

  p.N = hc.N;
  p.floating = hc.floating;
  p.ntasks = (int)hc.tasks.size();
  p.ncontacts = (int)hc.contacts.size();
  p.ndes = hc.ndes;
  for (int i = 0; i < m.nv; i++) {
    p.reg[i] = hc.reg[i];
    p.vcol[i] = 0;
    p.vfix_des[i] = -1;
  }
  // velocities fixed by hard JointAccelerationTasks are substituted out of the QP
  for (auto& t : hc.tasks)
    if (t.kind == 4 && t.mode == 0)
      for (int r = 0; r < t.dim; r++) {
        int vi = m.voff[t.joint] + r;
        if (p.vfix_des[vi] >= 0) return "two hard JointAccelerationTasks on one joint";
        p.vfix_des[vi] = t.des_off + r;
      }
  p.nvf = 0;
  for (int i = 0; i < m.nv; i++) p.vcol[i] = p.vfix_des[i] >= 0 ? -1 : p.nvf++;
  p.nbx = p.ncontacts * p.N;
  // weighted tasks keep the reference's slack variables e == task_error with cost w e'e / e'We (momentum.jl:107-126):
  // condensing them into J'WJ squares the Jacobian scale into P and makes OSQP's relative dual tolerance far looser
  // than it is on the reference's own QP
  p.ne = 0;
  for (auto& t : hc.tasks)
    if (t.mode != 0) p.ne += t.dim;
  p.n = p.nvf + p.ne + p.nbx;
  int row = 0, npath = 0, nw = 0, scol = p.nvf;
  for (int ti = 0; ti < p.ntasks; ti++) {
    const HostTask& t = hc.tasks[ti];
    DevTask& d = p.tasks[ti];
    d.kind = t.kind;
    d.mode = t.mode;
    d.dim = t.dim;
    d.source = t.source;
    d.target = t.target;
    d.frame = t.kind == 3 ? t.source : t.frame;
    d.joint = t.joint;
    d.des_off = t.des_off;
    d.weight = t.weight;
    for (int i = 0; i < 3; i++) d.point[i] = t.point[i];
    d.eliminated = (t.kind == 4 && t.mode == 0) ? 1 : 0;
    d.row0 = -1;
    d.scol0 = -1;
    if (!d.eliminated) {
      d.row0 = row;
      row += t.dim;
    }
    if (t.mode != 0) {
      d.scol0 = scol;
      scol += t.dim;
    }
    d.w_off = nw;
    if (t.mode == 2) {
      if (nw + t.dim * t.dim > QPC_MAXW) return "matrix weights exceed storage";
      for (int i = 0; i < t.dim * t.dim; i++) p.Wbuf[nw + i] = t.W[i];
      nw += t.dim * t.dim;
    }
    d.path_ptr = npath;
    d.path_len = 0;
    if (t.kind <= 3) {
      std::vector<int> up, down;
      for (int b = t.source; b >= 0; b = m.parent[b]) up.push_back(b);
      for (int b = t.target; b >= 0; b = m.parent[b]) down.push_back(b);
      while (!up.empty() && !down.empty() && up.back() == down.back()) {
        up.pop_back();
        down.pop_back();
      }
      if (npath + (int)(up.size() + down.size()) > QPC_MAXPATH) return "task paths exceed storage";
      for (int b : up) {
        p.path_body[npath] = b;
        p.path_sign[npath++] = -1;
      }
      for (int i = (int)down.size() - 1; i >= 0; i--) {
        p.path_body[npath] = down[i];
        p.path_sign[npath++] = +1;
      }
      d.path_len = npath - d.path_ptr;
    }
    for (int c = 0; c < QPC_MAXV; c++) p.tsign[ti][c] = 0;
    for (int e = d.path_ptr; e < d.path_ptr + d.path_len; e++) {
      const int b = p.path_body[e];
      for (int c = m.voff[b]; c < m.voff[b] + m.nvj[b]; c++) p.tsign[ti][c] = (signed char)p.path_sign[e];
    }
    for (int i = 0; i < t.dim; i++) p.def_desired[t.des_off + i] = t.desired.empty() ? 0.0 : t.desired[i];
  }
  p.nfixv = 0;
  for (int i = 0; i < m.nv; i++)
    if (p.vcol[i] < 0) p.fixv[p.nfixv++] = i;
  p.nactive = 0;
  for (int ti = 0; ti < p.ntasks; ti++)
    if (!p.tasks[ti].eliminated) p.active[p.nactive++] = ti;
  p.balance_row0 = -1;
  if (hc.floating >= 0) {
    if (m.jtype[hc.floating] != 2) return "floating joint must be a quaternion floating joint";
    p.balance_row0 = row;
    row += 6;
  }
  p.mg = row;
  p.nwmat = nw;
  for (int c = 0; c < p.ncontacts; c++) {
    const HostContact& hcn = hc.contacts[c];
    DevContact& d = p.contacts[c];
    d.body = hcn.body;
    d.col0 = p.nvf + p.ne + c * p.N;
    rotation_between_z(hcn.normal, d.Rz);
    for (int i = 0; i < 3; i++) d.pos[i] = hcn.pos[i];
    for (int g = 0; g < p.N; g++) {  // forcebasis (contacts.jl:16-23): unit vectors on the cone
      double th = g * (2 * M_PI / p.N);
      double v[3] = {hcn.mu * std::cos(th), hcn.mu * std::sin(th), 1.0};
      double nn = std::sqrt(v[0] * v[0] + v[1] * v[1] + v[2] * v[2]);
      for (int r = 0; r < 3; r++) d.B[r * p.N + g] = v[r] / nn;
    }
    for (int a = 0; a < p.N; a++)
      for (int b = 0; b < p.N; b++) {
        double s = 0;
        for (int r = 0; r < 3; r++) s += d.B[r * p.N + a] * d.B[r * p.N + b];
        d.BtB[a * p.N + b] = s;
      }
    d.maxrho_factor = 1.0 / (p.N * std::sqrt(hcn.mu * hcn.mu + 1.0));
    p.def_cweight[c] = hcn.weight;
    p.def_cmaxnf[c] = hcn.maxnf;
  }
  // fast-path eligibility (admm_reg.cuh, ELIM): every free velocity regularised, every weighted task scalar-weighted
  // with a positive weight, at least one kept (contact) variable.  All general rows of a controller program are
  // equalities (hard tasks, slack definitions, wrench balance), which the elimination relies on.
  p.nel = p.nvf + p.ne;
  for (int i = 0; i < m.nv; i++)
    if (p.vcol[i] >= 0 && !(p.reg[i] > 0.0)) p.nel = 0;
  for (auto& t : hc.tasks)
    if (t.mode == 2 || (t.mode == 1 && !(t.weight > 0.0))) p.nel = 0;
  if (p.nbx == 0) p.nel = 0;
  const HostStanding& st = hc.standing;
  p.standing = st.enabled ? 1 : 0;
  if (st.enabled) {
    p.st_linmom_des = hc.tasks[st.linmom_task].des_off;
    p.st_pelvis_des = hc.tasks[st.pelvis_task].des_off;
    p.st_pelvis_body = st.pelvis_body;
    p.st_nj = (int)st.joints.size();
    for (int i = 0; i < p.st_nj; i++) {
      p.st_jq[i] = m.qoff[st.joints[i]];
      p.st_jv[i] = m.voff[st.joints[i]];
      p.st_jdes[i] = hc.tasks[st.joint_tasks[i]].des_off;
      p.st_kp[i] = st.kp[i];
      p.st_kd[i] = st.kd[i];
      p.st_ref[i] = st.ref[i];
    }
    p.st_com_kp = st.com_kp;
    p.st_com_kd = st.com_kd;
    p.st_pelvis_kp = st.pelvis_kp;
    p.st_pelvis_kd = st.pelvis_kd;
    for (int i = 0; i < 3; i++) p.st_comref[i] = st.comref[i];
  }
  p.nse3 = (int)hc.se3.size();
  for (int i = 0; i < p.nse3; i++) {
    const HostSE3PD& h = hc.se3[i];
    DevSE3PD& d = p.se3[i];
    d.body = h.body;
    d.base = h.base;
    d.task = h.task;
    d.des_off = hc.tasks[h.task].des_off;
    std::memcpy(d.K, h.K, sizeof(d.K));
    d.ang = h.ang;
    d.lin = h.lin;
  }
  p.settings = hc.settings;
  return "";
}

}  // namespace qpc
