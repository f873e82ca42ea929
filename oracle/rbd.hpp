// ORACLE -- TEST INFRASTRUCTURE ONLY.  Never linked into, imported by, or called from the product path.
//
// CPU fp64 restatement of the RigidBodyDynamics.jl 2.2.0 quantities that QPControl.jl's control tick evaluates
// (SURVEY.md section 8(a) row a15 and appendix B.1).  RigidBodyDynamics is an un-vendored third-party dependency
// of the reference (reference Manifest.toml:130-134), so this file restates its published algorithms; the call
// sites in the reference that each function serves are cited next to it.
//
// PARITY UNPINNED: the reference ships no golden vectors and cannot run offline (no Julia); this restatement is
// pinned by the reference's own invariant tests restated in tests/ (finite differences, energy/momentum
// consistency, forward-dynamics round trips), not by reference outputs.
#pragma once
#include <cmath>
#include <cstring>
#include <vector>

namespace orc {

struct V3 {
  double x = 0, y = 0, z = 0;
  double& operator[](int i) { return (&x)[i]; }
  double operator[](int i) const { return (&x)[i]; }
};
inline V3 operator+(V3 a, V3 b) { return {a.x + b.x, a.y + b.y, a.z + b.z}; }
inline V3 operator-(V3 a, V3 b) { return {a.x - b.x, a.y - b.y, a.z - b.z}; }
inline V3 operator-(V3 a) { return {-a.x, -a.y, -a.z}; }
inline V3 operator*(double s, V3 a) { return {s * a.x, s * a.y, s * a.z}; }
inline double dot(V3 a, V3 b) { return a.x * b.x + a.y * b.y + a.z * b.z; }
inline V3 cross(V3 a, V3 b) { return {a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x}; }
inline double norm(V3 a) { return std::sqrt(dot(a, a)); }

struct M3 {
  double a[3][3] = {{0, 0, 0}, {0, 0, 0}, {0, 0, 0}};
  static M3 eye() {
    M3 m;
    m.a[0][0] = m.a[1][1] = m.a[2][2] = 1;
    return m;
  }
};
inline V3 operator*(const M3& m, V3 v) {
  return {m.a[0][0] * v.x + m.a[0][1] * v.y + m.a[0][2] * v.z, m.a[1][0] * v.x + m.a[1][1] * v.y + m.a[1][2] * v.z,
          m.a[2][0] * v.x + m.a[2][1] * v.y + m.a[2][2] * v.z};
}
inline M3 operator*(const M3& p, const M3& q) {
  M3 r;
  for (int i = 0; i < 3; i++)
    for (int j = 0; j < 3; j++) r.a[i][j] = p.a[i][0] * q.a[0][j] + p.a[i][1] * q.a[1][j] + p.a[i][2] * q.a[2][j];
  return r;
}
inline M3 operator+(const M3& p, const M3& q) {
  M3 r;
  for (int i = 0; i < 3; i++)
    for (int j = 0; j < 3; j++) r.a[i][j] = p.a[i][j] + q.a[i][j];
  return r;
}
inline M3 transpose(const M3& p) {
  M3 r;
  for (int i = 0; i < 3; i++)
    for (int j = 0; j < 3; j++) r.a[i][j] = p.a[j][i];
  return r;
}
inline V3 tmul(const M3& m, V3 v) { return transpose(m) * v; }
// hat(a)^2 = a a' - (a.a) I
inline M3 hat_squared(V3 a) {
  M3 r;
  double d = dot(a, a);
  for (int i = 0; i < 3; i++)
    for (int j = 0; j < 3; j++) r.a[i][j] = a[i] * a[j] - (i == j ? d : 0.0);
  return r;
}

// Transform3D(from -> to): x_to = R x_from + p
struct Xf {
  M3 R = M3::eye();
  V3 p;
};
inline Xf operator*(const Xf& a, const Xf& b) { return {a.R * b.R, a.R * b.p + a.p}; }
inline Xf inv(const Xf& a) { return {transpose(a.R), -(tmul(a.R, a.p))}; }

// spatial vector (angular; linear): twists, accelerations, wrenches, momenta
struct S6 {
  V3 w, v;
};
inline S6 operator+(S6 a, S6 b) { return {a.w + b.w, a.v + b.v}; }
inline S6 operator-(S6 a, S6 b) { return {a.w - b.w, a.v - b.v}; }
inline S6 operator-(S6 a) { return {-a.w, -a.v}; }
inline S6 operator*(double s, S6 a) { return {s * a.w, s * a.v}; }
inline double dot(S6 a, S6 b) { return dot(a.w, b.w) + dot(a.v, b.v); }
// motion transform: w' = R w, v' = R v + p x (R w)      
inline S6 xmotion(const Xf& X, S6 t) {
  V3 w = X.R * t.w;
  return {w, X.R * t.v + cross(X.p, w)};
}
// force transform: f' = R f, tau' = R tau + p x (R f)
inline S6 xforce(const Xf& X, S6 f) {
  V3 l = X.R * f.v;
  return {X.R * f.w + cross(X.p, l), l};
}
// se(3) commutator / motion cross product [x, y]
inline S6 cross_motion(S6 x, S6 y) { return {cross(x.w, y.w), cross(x.w, y.v) + cross(x.v, y.w)}; }
// twist x* momentum
inline S6 cross_force(S6 t, S6 h) { return {cross(t.w, h.w) + cross(t.v, h.v), cross(t.w, h.v)}; }

// SpatialInertia(frame, moment J about the frame origin, cross part c = m*com, mass)
struct SI {
  M3 J;
  V3 c;
  double m = 0;
};
inline SI operator+(const SI& a, const SI& b) { return {a.J + b.J, a.c + b.c, a.m + b.m}; }
inline S6 operator*(const SI& I, S6 t) { return {I.J * t.w + cross(I.c, t.v), I.m * t.v - cross(I.c, t.w)}; }
inline SI xinertia(const Xf& X, const SI& I) {
  SI r;
  V3 Rc = X.R * I.c;
  r.m = I.m;
  r.c = Rc + I.m * X.p;
  r.J = X.R * I.J * transpose(X.R);
  if (I.m > 0) {
    M3 a = hat_squared(Rc), b = hat_squared(r.c);
    for (int i = 0; i < 3; i++)
      for (int j = 0; j < 3; j++) r.J.a[i][j] += (a.a[i][j] - b.a[i][j]) / I.m;
  }
  return r;
}
inline S6 newton_euler(const SI& I, S6 accel, S6 twist) { return I * accel + cross_force(twist, I * twist); }

enum JointType { REVOLUTE = 0, PRISMATIC = 1, QUAT_FLOATING = 2, FIXED = 3 };

struct Mechanism {
  int nb = 0, nq = 0, nv = 0;
  std::vector<int> parent, jtype, qoff, voff, nqj, nvj;
  std::vector<int> vbody;  // velocity index -> body
  std::vector<V3> axis;
  std::vector<Xf> Xtree;
  std::vector<SI> inertia;  // body frame
  V3 gravity;
  double total_mass = 0;
};

inline M3 quat_to_rot(double w, double x, double y, double z) {
  M3 R;
  R.a[0][0] = 1 - 2 * (y * y + z * z);
  R.a[0][1] = 2 * (x * y - w * z);
  R.a[0][2] = 2 * (x * z + w * y);
  R.a[1][0] = 2 * (x * y + w * z);
  R.a[1][1] = 1 - 2 * (x * x + z * z);
  R.a[1][2] = 2 * (y * z - w * x);
  R.a[2][0] = 2 * (x * z - w * y);
  R.a[2][1] = 2 * (y * z + w * x);
  R.a[2][2] = 1 - 2 * (x * x + y * y);
  return R;
}
inline M3 axis_angle_to_rot(V3 k, double th) {
  double c = std::cos(th), s = std::sin(th), v = 1 - c;
  M3 R;
  R.a[0][0] = k.x * k.x * v + c;
  R.a[0][1] = k.x * k.y * v - k.z * s;
  R.a[0][2] = k.x * k.z * v + k.y * s;
  R.a[1][0] = k.y * k.x * v + k.z * s;
  R.a[1][1] = k.y * k.y * v + c;
  R.a[1][2] = k.y * k.z * v - k.x * s;
  R.a[2][0] = k.z * k.x * v - k.y * s;
  R.a[2][1] = k.z * k.y * v + k.x * s;
  R.a[2][2] = k.z * k.z * v + c;
  return R;
}
// rotation vector (Rotations.RodriguesVec) of a rotation matrix, via the unit quaternion with w >= 0
inline V3 rot_to_rotvec(const M3& R) {
  double w, x, y, z;
  double tr = R.a[0][0] + R.a[1][1] + R.a[2][2];
  if (tr > 0) {
    double s = std::sqrt(tr + 1.0) * 2;
    w = 0.25 * s;
    x = (R.a[2][1] - R.a[1][2]) / s;
    y = (R.a[0][2] - R.a[2][0]) / s;
    z = (R.a[1][0] - R.a[0][1]) / s;
  } else if (R.a[0][0] > R.a[1][1] && R.a[0][0] > R.a[2][2]) {
    double s = std::sqrt(1.0 + R.a[0][0] - R.a[1][1] - R.a[2][2]) * 2;
    w = (R.a[2][1] - R.a[1][2]) / s;
    x = 0.25 * s;
    y = (R.a[0][1] + R.a[1][0]) / s;
    z = (R.a[0][2] + R.a[2][0]) / s;
  } else if (R.a[1][1] > R.a[2][2]) {
    double s = std::sqrt(1.0 + R.a[1][1] - R.a[0][0] - R.a[2][2]) * 2;
    w = (R.a[0][2] - R.a[2][0]) / s;
    x = (R.a[0][1] + R.a[1][0]) / s;
    y = 0.25 * s;
    z = (R.a[1][2] + R.a[2][1]) / s;
  } else {
    double s = std::sqrt(1.0 + R.a[2][2] - R.a[0][0] - R.a[1][1]) * 2;
    w = (R.a[1][0] - R.a[0][1]) / s;
    x = (R.a[0][2] + R.a[2][0]) / s;
    y = (R.a[1][2] + R.a[2][1]) / s;
    z = 0.25 * s;
  }
  if (w < 0) {
    w = -w;
    x = -x;
    y = -y;
    z = -z;
  }
  double n = std::sqrt(x * x + y * y + z * z);
  if (n < 1e-15) return {2 * x, 2 * y, 2 * z};
  double th = 2 * std::atan2(n, w);
  return {x * th / n, y * th / n, z * th / n};
}

// Rotations.rotation_between(u = (0,0,1), v): the minimal rotation taking u to v (reference src/contacts.jl:11)
inline M3 rotation_between_z(V3 v) {
  double n = norm(v);
  V3 t = (1.0 / n) * v;
  V3 u{0, 0, 1};
  V3 ax = cross(u, t);
  double s = norm(ax), c = dot(u, t);
  if (s < 1e-14) {
    if (c > 0) return M3::eye();
    M3 R = M3::eye();  // v = -u: undefined in the reference; pick the rotation by pi about x
    R.a[1][1] = -1;
    R.a[2][2] = -1;
    return R;
  }
  return axis_angle_to_rot((1.0 / s) * ax, std::atan2(s, c));
}

// Cached kinematic state: what RBD's MechanismState caches after `copyto!(state, x)` (momentum.jl:57).
struct State {
  std::vector<double> q, v;
  std::vector<Xf> toroot;    // transform_to_root(body)
  std::vector<S6> twist;     // twist_wrt_world(body), world frame
  std::vector<S6> bias;      // bias_acceleration(body), world frame
  std::vector<S6> Sw;        // [nv] motion subspace columns in world frame
  std::vector<SI> Iw, Ic;    // world-frame body inertia, composite (crb) inertia
  V3 com;
  S6 momentum;
  S6 momentum_rate_bias;
};

inline const Xf& toroot(const State& s, int body) {
  static const Xf I;
  return body < 0 ? I : s.toroot[body];
}
inline S6 twist_of(const State& s, int body) { return body < 0 ? S6{} : s.twist[body]; }
inline S6 bias_of(const State& s, int body) { return body < 0 ? S6{} : s.bias[body]; }

inline Xf joint_transform(const Mechanism& m, int b, const double* q) {
  Xf X;
  const double* qj = q + m.qoff[b];
  switch (m.jtype[b]) {
    case REVOLUTE: X.R = axis_angle_to_rot(m.axis[b], qj[0]); break;
    case PRISMATIC: X.p = qj[0] * m.axis[b]; break;
    case QUAT_FLOATING: {
      double n = std::sqrt(qj[0] * qj[0] + qj[1] * qj[1] + qj[2] * qj[2] + qj[3] * qj[3]);
      X.R = quat_to_rot(qj[0] / n, qj[1] / n, qj[2] / n, qj[3] / n);
      X.p = {qj[4], qj[5], qj[6]};
      break;
    }
    default: break;
  }
  return X;
}

inline void update_state(const Mechanism& m, const double* q, const double* v, State& s) {
  s.q.assign(q, q + m.nq);
  s.v.assign(v, v + m.nv);
  s.toroot.resize(m.nb);
  s.twist.resize(m.nb);
  s.bias.resize(m.nb);
  s.Sw.resize(m.nv);
  s.Iw.resize(m.nb);
  s.Ic.resize(m.nb);
  for (int b = 0; b < m.nb; b++) {
    int p = m.parent[b];
    s.toroot[b] = toroot(s, p) * m.Xtree[b] * joint_transform(m, b, q);
    const Xf& H = s.toroot[b];
    S6 jt;  // joint twist (body wrt parent) in world frame
    int o = m.voff[b];
    switch (m.jtype[b]) {
      case REVOLUTE: s.Sw[o] = xmotion(H, S6{m.axis[b], V3{}}); break;
      case PRISMATIC: s.Sw[o] = xmotion(H, S6{V3{}, m.axis[b]}); break;
      case QUAT_FLOATING:
        for (int k = 0; k < 3; k++) {
          V3 e;
          e[k] = 1;
          s.Sw[o + k] = xmotion(H, S6{e, V3{}});
          s.Sw[o + 3 + k] = xmotion(H, S6{V3{}, e});
        }
        break;
      default: break;
    }
    for (int k = 0; k < m.nvj[b]; k++) jt = jt + v[o + k] * s.Sw[o + k];
    s.twist[b] = twist_of(s, p) + jt;
    // world-frame bias acceleration: parent's + T_body x (S v); the joints' own bias is zero
    s.bias[b] = bias_of(s, p) + cross_motion(s.twist[b], jt);
    s.Iw[b] = xinertia(H, m.inertia[b]);
  }
  for (int b = 0; b < m.nb; b++) s.Ic[b] = s.Iw[b];
  for (int b = m.nb - 1; b >= 0; b--)
    if (m.parent[b] >= 0) s.Ic[m.parent[b]] = s.Ic[m.parent[b]] + s.Ic[b];
  V3 mc;
  s.momentum = S6{};
  s.momentum_rate_bias = S6{};
  for (int b = 0; b < m.nb; b++) {
    mc = mc + s.Iw[b].c;
    s.momentum = s.momentum + s.Iw[b] * s.twist[b];
    s.momentum_rate_bias = s.momentum_rate_bias + newton_euler(s.Iw[b], s.bias[b], s.twist[b]);
  }
  s.com = (1.0 / m.total_mass) * mc;
}

// path(mechanism, source, target) as (body-of-joint, sign) pairs
inline void tree_path(const Mechanism& m, int source, int target, std::vector<int>& joints, std::vector<int>& signs) {
  joints.clear();
  signs.clear();
  std::vector<int> up, down;
  for (int b = source; b >= 0; b = m.parent[b]) up.push_back(b);
  for (int b = target; b >= 0; b = m.parent[b]) down.push_back(b);
  while (!up.empty() && !down.empty() && up.back() == down.back()) {
    up.pop_back();
    down.pop_back();
  }
  for (int b : up) {
    joints.push_back(b);
    signs.push_back(-1);
  }
  for (int i = (int)down.size() - 1; i >= 0; i--) {
    joints.push_back(down[i]);
    signs.push_back(+1);
  }
}

// geometric_jacobian!(J, state, path, world -> frame): 6 x nv row-major (rows 0-2 angular, 3-5 linear)
// (reference src/tasks.jl:34,76,115)
inline void geometric_jacobian(const Mechanism& m, const State& s, int source, int target, int frame, double* J) {
  std::memset(J, 0, sizeof(double) * 6 * m.nv);
  std::vector<int> joints, signs;
  tree_path(m, source, target, joints, signs);
  Xf X = inv(toroot(s, frame));
  for (size_t k = 0; k < joints.size(); k++) {
    int b = joints[k];
    for (int c = m.voff[b]; c < m.voff[b] + m.nvj[b]; c++) {
      S6 col = (double)signs[k] * xmotion(X, s.Sw[c]);
      for (int r = 0; r < 3; r++) {
        J[r * m.nv + c] = col.w[r];
        J[(3 + r) * m.nv + c] = col.v[r];
      }
    }
  }
}

// transform(state, accel(body=target, base=source, frame=world), to frame) (RBD):
// X_{world->F} [ accel + (-T_F) x_m (T_target - T_source) ]         (reference src/tasks.jl:36-39)
inline S6 bias_in_frame(const State& s, int source, int target, int frame) {
  S6 a = bias_of(s, target) - bias_of(s, source);
  S6 rel = twist_of(s, target) - twist_of(s, source);
  S6 old_wrt_new = -twist_of(s, frame);
  return xmotion(inv(toroot(s, frame)), a + cross_motion(old_wrt_new, rel));
}

// momentum_matrix!(A, state[, world -> frame]): 6 x nv row-major (reference src/tasks.jl:212, momentum.jl:168)
inline void momentum_matrix(const Mechanism& m, const State& s, const Xf& world_to_frame, double* A) {
  for (int c = 0; c < m.nv; c++) {
    S6 col = xforce(world_to_frame, s.Ic[m.vbody[c]] * s.Sw[c]);
    for (int r = 0; r < 3; r++) {
      A[r * m.nv + c] = col.w[r];
      A[(3 + r) * m.nv + c] = col.v[r];
    }
  }
}

// mass_matrix (composite rigid body algorithm) -- needed only by the test-side forward dynamics
inline void mass_matrix(const Mechanism& m, const State& s, double* M) {
  std::memset(M, 0, sizeof(double) * m.nv * m.nv);
  for (int i = 0; i < m.nv; i++) {
    int bi = m.vbody[i];
    S6 F = s.Ic[bi] * s.Sw[i];
    for (int b = bi; b >= 0; b = m.parent[b])
      for (int j = m.voff[b]; j < m.voff[b] + m.nvj[b]; j++) {
        if (b == bi && j > i) continue;
        double val = dot(s.Sw[j], F);
        M[i * m.nv + j] = val;
        M[j * m.nv + i] = val;
      }
  }
}

// inverse_dynamics!(tau, jointwrenches, accelerations, state, vd, externalwrenches) (reference momentum.jl:75):
// recursive Newton-Euler, gravity through the root acceleration, external wrenches per body in world frame.
inline void inverse_dynamics(const Mechanism& m, const State& s, const double* vd, const S6* ext /*[nb] or null*/,
                             double* tau) {
  std::vector<S6> acc(m.nb), jw(m.nb);
  S6 root{V3{}, -m.gravity};
  for (int b = 0; b < m.nb; b++) {
    int p = m.parent[b];
    S6 a = p < 0 ? root : acc[p];
    for (int k = m.voff[b]; k < m.voff[b] + m.nvj[b]; k++) a = a + vd[k] * s.Sw[k];
    // spatial acceleration = parent's + S vd + (bias_b - bias_parent)
    a = a + (s.bias[b] - bias_of(s, p));
    acc[b] = a;
    jw[b] = newton_euler(s.Iw[b], a, s.twist[b]);
    if (ext) jw[b] = jw[b] - ext[b];
  }
  for (int b = m.nb - 1; b >= 0; b--) {
    for (int k = m.voff[b]; k < m.voff[b] + m.nvj[b]; k++) tau[k] = dot(s.Sw[k], jw[b]);
    if (m.parent[b] >= 0) jw[m.parent[b]] = jw[m.parent[b]] + jw[b];
  }
}

}  // namespace orc
