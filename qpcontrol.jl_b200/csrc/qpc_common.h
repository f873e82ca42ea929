// qpc_common.h -- execution-model macros and small fp64 spatial algebra shared by all kernels.
//
// Every kernel body in this directory is written as "block-cooperative" code: loops strided by QPC_TID / QPC_NT over
// shared-memory arrays, phases separated by QPC_SYNC().  Compiled by nvcc for sm_100a these map to threadIdx.x /
// blockDim.x / __syncthreads().  tests/emu compiles the very same bodies with g++ as a single "thread"
// (QPC_TID = 0, QPC_NT = 1, QPC_SYNC = no-op) so the arithmetic can be debugged without a GPU; that build is test
// infrastructure only and is never loaded by the product path.
#pragma once
#include <math.h>
#include <stdint.h>

#if defined(__CUDACC__)
#define QPC_HD __host__ __device__ __forceinline__
#define QPC_DEV __device__ __forceinline__
#define QPC_DEVN __device__ __noinline__
#if defined(QPC_THREAD_PER_INSTANCE)  // tiny_thread.cu: one THREAD per robot instance, workspace in local memory
#define QPC_TID 0
#define QPC_NT 1
#define QPC_SYNC() ((void)0)
#define QPC_SERIAL 1
#else
#define QPC_TID ((int)threadIdx.x)
#define QPC_NT ((int)blockDim.x)
#if defined(QPC_WARP_PER_INSTANCE)  // kin_warp.cu: blockDim = (32, instances per CTA), one warp per robot instance
#define QPC_SYNC() __syncwarp()
#else
#define QPC_SYNC() __syncthreads()
#endif
#endif
#define QPC_LDG(p) __ldg(p)
#define QPC_UNROLL8 _Pragma("unroll 8")  // product loops: eight loads in flight instead of one dependent load per FMA
#else
#define QPC_HD inline
#define QPC_DEV inline
#define QPC_DEVN inline
#define QPC_TID 0
#define QPC_NT 1
#define QPC_SYNC() ((void)0)
#define QPC_SERIAL 1
#define QPC_LDG(p) (*(p))
#define QPC_UNROLL8
#endif

namespace qpc {

struct V3 {
  double x, y, z;
};
QPC_HD V3 mk3(double x, double y, double z) {
  V3 r;
  r.x = x;
  r.y = y;
  r.z = z;
  return r;
}
QPC_HD V3 ld3(const double* p) { return mk3(p[0], p[1], p[2]); }
QPC_HD void st3(double* p, V3 a) {
  p[0] = a.x;
  p[1] = a.y;
  p[2] = a.z;
}
QPC_HD V3 operator+(V3 a, V3 b) { return mk3(a.x + b.x, a.y + b.y, a.z + b.z); }
QPC_HD V3 operator-(V3 a, V3 b) { return mk3(a.x - b.x, a.y - b.y, a.z - b.z); }
QPC_HD V3 operator-(V3 a) { return mk3(-a.x, -a.y, -a.z); }
QPC_HD V3 operator*(double s, V3 a) { return mk3(s * a.x, s * a.y, s * a.z); }
QPC_HD double dot(V3 a, V3 b) { return a.x * b.x + a.y * b.y + a.z * b.z; }
QPC_HD V3 cross(V3 a, V3 b) { return mk3(a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x); }

// rigid transform from -> to, stored as 12 doubles: R row-major (9) then p (3); x_to = R x_from + p
struct Xf {
  double R[9];
  V3 p;
};
QPC_HD Xf xf_identity() {
  Xf X;
  for (int i = 0; i < 9; i++) X.R[i] = (i % 4 == 0) ? 1.0 : 0.0;
  X.p = mk3(0, 0, 0);
  return X;
}
QPC_HD Xf xf_load(const double* s) {
  Xf X;
  for (int i = 0; i < 9; i++) X.R[i] = s[i];
  X.p = ld3(s + 9);
  return X;
}
QPC_HD void xf_store(double* s, const Xf& X) {
  for (int i = 0; i < 9; i++) s[i] = X.R[i];
  st3(s + 9, X.p);
}
QPC_HD V3 rot(const double* R, V3 v) {
  return mk3(R[0] * v.x + R[1] * v.y + R[2] * v.z, R[3] * v.x + R[4] * v.y + R[5] * v.z,
             R[6] * v.x + R[7] * v.y + R[8] * v.z);
}
QPC_HD V3 rot_t(const double* R, V3 v) {
  return mk3(R[0] * v.x + R[3] * v.y + R[6] * v.z, R[1] * v.x + R[4] * v.y + R[7] * v.z,
             R[2] * v.x + R[5] * v.y + R[8] * v.z);
}
QPC_HD Xf xf_mul(const Xf& a, const Xf& b) {
  Xf r;
  for (int i = 0; i < 3; i++)
    for (int j = 0; j < 3; j++) r.R[3 * i + j] = a.R[3 * i] * b.R[j] + a.R[3 * i + 1] * b.R[3 + j] + a.R[3 * i + 2] * b.R[6 + j];
  r.p = rot(a.R, b.p) + a.p;
  return r;
}
QPC_HD Xf xf_inv(const Xf& a) {
  Xf r;
  for (int i = 0; i < 3; i++)
    for (int j = 0; j < 3; j++) r.R[3 * i + j] = a.R[3 * j + i];
  r.p = -rot_t(a.R, a.p);
  return r;
}

// spatial vector (angular; linear)
struct S6 {
  V3 w, v;
};
QPC_HD S6 mk6(V3 w, V3 v) {
  S6 r;
  r.w = w;
  r.v = v;
  return r;
}
QPC_HD S6 s6_zero() { return mk6(mk3(0, 0, 0), mk3(0, 0, 0)); }
QPC_HD S6 ld6(const double* p) { return mk6(ld3(p), ld3(p + 3)); }
QPC_HD void st6(double* p, S6 a) {
  st3(p, a.w);
  st3(p + 3, a.v);
}
QPC_HD S6 operator+(S6 a, S6 b) { return mk6(a.w + b.w, a.v + b.v); }
QPC_HD S6 operator-(S6 a, S6 b) { return mk6(a.w - b.w, a.v - b.v); }
QPC_HD S6 operator*(double s, S6 a) { return mk6(s * a.w, s * a.v); }
QPC_HD double dot(S6 a, S6 b) { return dot(a.w, b.w) + dot(a.v, b.v); }
QPC_HD double s6_get(const S6& a, int i) {
  switch (i) {
    case 0: return a.w.x;
    case 1: return a.w.y;
    case 2: return a.w.z;
    case 3: return a.v.x;
    case 4: return a.v.y;
    default: return a.v.z;
  }
}
// twist / acceleration change of frame
QPC_HD S6 xmotion(const Xf& X, S6 t) {
  V3 w = rot(X.R, t.w);
  return mk6(w, rot(X.R, t.v) + cross(X.p, w));
}
// wrench / momentum change of frame
QPC_HD S6 xforce(const Xf& X, S6 f) {
  V3 l = rot(X.R, f.v);
  return mk6(rot(X.R, f.w) + cross(X.p, l), l);
}
QPC_HD S6 cross_motion(S6 a, S6 b) { return mk6(cross(a.w, b.w), cross(a.w, b.v) + cross(a.v, b.w)); }
QPC_HD S6 cross_force(S6 t, S6 h) { return mk6(cross(t.w, h.w) + cross(t.v, h.v), cross(t.w, h.v)); }

// spatial inertia as 10 doubles: symmetric moment about the frame origin (xx, xy, xz, yy, yz, zz), m*com (3), mass
struct SI {
  double J[6];
  V3 c;
  double m;
};
QPC_HD SI si_load(const double* s) {
  SI I;
  for (int i = 0; i < 6; i++) I.J[i] = s[i];
  I.c = ld3(s + 6);
  I.m = s[9];
  return I;
}
QPC_HD void si_store(double* s, const SI& I) {
  for (int i = 0; i < 6; i++) s[i] = I.J[i];
  st3(s + 6, I.c);
  s[9] = I.m;
}
QPC_HD V3 sym_mul(const double* J, V3 v) {
  return mk3(J[0] * v.x + J[1] * v.y + J[2] * v.z, J[1] * v.x + J[3] * v.y + J[4] * v.z,
             J[2] * v.x + J[4] * v.y + J[5] * v.z);
}
QPC_HD S6 si_mul(const SI& I, S6 t) { return mk6(sym_mul(I.J, t.w) + cross(I.c, t.v), I.m * t.v - cross(I.c, t.w)); }
QPC_HD S6 newton_euler(const SI& I, S6 a, S6 t) { return si_mul(I, a) + cross_force(t, si_mul(I, t)); }
// express an inertia given in the `from` frame of X in its `to` frame
QPC_HD SI si_transform(const Xf& X, const SI& I) {
  SI r;
  const double* R = X.R;
  // R J R'
  double JR[9];  // J * R'
  double Jf[9] = {I.J[0], I.J[1], I.J[2], I.J[1], I.J[3], I.J[4], I.J[2], I.J[4], I.J[5]};
  for (int i = 0; i < 3; i++)
    for (int j = 0; j < 3; j++) JR[3 * i + j] = Jf[3 * i] * R[3 * j] + Jf[3 * i + 1] * R[3 * j + 1] + Jf[3 * i + 2] * R[3 * j + 2];
  double o[6];
  int k = 0;
  for (int i = 0; i < 3; i++)
    for (int j = i; j < 3; j++) o[k++] = R[3 * i] * JR[j] + R[3 * i + 1] * JR[3 + j] + R[3 * i + 2] * JR[6 + j];
  V3 rc = rot(R, I.c);
  V3 cn = rc + I.m * X.p;
  if (I.m > 0) {
    // + (hat(rc)^2 - hat(cn)^2) / m, hat(a)^2 = a a' - |a|^2 I
    double inv = 1.0 / I.m;
    double d1 = dot(rc, rc), d2 = dot(cn, cn);
    o[0] += (rc.x * rc.x - d1 - cn.x * cn.x + d2) * inv;
    o[1] += (rc.x * rc.y - cn.x * cn.y) * inv;
    o[2] += (rc.x * rc.z - cn.x * cn.z) * inv;
    o[3] += (rc.y * rc.y - d1 - cn.y * cn.y + d2) * inv;
    o[4] += (rc.y * rc.z - cn.y * cn.z) * inv;
    o[5] += (rc.z * rc.z - d1 - cn.z * cn.z + d2) * inv;
  }
  for (int i = 0; i < 6; i++) r.J[i] = o[i];
  r.c = cn;
  r.m = I.m;
  return r;
}

QPC_HD void quat_to_rot(double w, double x, double y, double z, double* R) {
  R[0] = 1 - 2 * (y * y + z * z);
  R[1] = 2 * (x * y - w * z);
  R[2] = 2 * (x * z + w * y);
  R[3] = 2 * (x * y + w * z);
  R[4] = 1 - 2 * (x * x + z * z);
  R[5] = 2 * (y * z - w * x);
  R[6] = 2 * (x * z - w * y);
  R[7] = 2 * (y * z + w * x);
  R[8] = 1 - 2 * (x * x + y * y);
}
// Rotations.rotation_between((0,0,1), v)  (reference src/contacts.jl:11): Rodrigues with sin / cos taken from the cross
// and dot products directly.  Same result as host_program.h: rotation_between_z up to rounding.
QPC_HD void rot_between_z(V3 v, double* R) {
  const double n = sqrt(dot(v, v));
  const V3 t = (1.0 / n) * v;
  const double ax = -t.y, ay = t.x;  // (0,0,1) x t
  const double s = sqrt(ax * ax + ay * ay), c = t.z;
  for (int i = 0; i < 9; i++) R[i] = (i % 4 == 0) ? 1.0 : 0.0;
  if (s < 1e-14) {
    if (c < 0) {
      R[4] = -1;
      R[8] = -1;
    }
    return;
  }
  const double kx = ax / s, ky = ay / s, u = 1 - c;
  R[0] = kx * kx * u + c;
  R[1] = kx * ky * u;
  R[2] = ky * s;
  R[3] = ky * kx * u;
  R[4] = ky * ky * u + c;
  R[5] = -kx * s;
  R[6] = -ky * s;
  R[7] = kx * s;
  R[8] = c;
}
QPC_HD void axis_angle_to_rot(V3 k, double th, double* R) {
  double s, c;
#if defined(__CUDA_ARCH__)
  sincos(th, &s, &c);
#else
  s = sin(th);
  c = cos(th);
#endif
  double t = 1 - c;
  R[0] = k.x * k.x * t + c;
  R[1] = k.x * k.y * t - k.z * s;
  R[2] = k.x * k.z * t + k.y * s;
  R[3] = k.y * k.x * t + k.z * s;
  R[4] = k.y * k.y * t + c;
  R[5] = k.y * k.z * t - k.x * s;
  R[6] = k.z * k.x * t - k.y * s;
  R[7] = k.z * k.y * t + k.x * s;
  R[8] = k.z * k.z * t + c;
}
// rotation vector of a rotation matrix (through the unit quaternion with w >= 0)
QPC_HD V3 rot_to_rotvec(const double* R) {
  double w, x, y, z;
  double tr = R[0] + R[4] + R[8];
  if (tr > 0) {
    double s = sqrt(tr + 1.0) * 2;
    w = 0.25 * s;
    x = (R[7] - R[5]) / s;
    y = (R[2] - R[6]) / s;
    z = (R[3] - R[1]) / s;
  } else if (R[0] > R[4] && R[0] > R[8]) {
    double s = sqrt(1.0 + R[0] - R[4] - R[8]) * 2;
    w = (R[7] - R[5]) / s;
    x = 0.25 * s;
    y = (R[1] + R[3]) / s;
    z = (R[2] + R[6]) / s;
  } else if (R[4] > R[8]) {
    double s = sqrt(1.0 + R[4] - R[0] - R[8]) * 2;
    w = (R[2] - R[6]) / s;
    x = (R[1] + R[3]) / s;
    y = 0.25 * s;
    z = (R[5] + R[7]) / s;
  } else {
    double s = sqrt(1.0 + R[8] - R[0] - R[4]) * 2;
    w = (R[3] - R[1]) / s;
    x = (R[2] + R[6]) / s;
    y = (R[5] + R[7]) / s;
    z = 0.25 * s;
  }
  if (w < 0) {
    w = -w;
    x = -x;
    y = -y;
    z = -z;
  }
  double n = sqrt(x * x + y * y + z * z);
  if (n < 1e-15) return mk3(2 * x, 2 * y, 2 * z);
  double th = 2 * atan2(n, w);
  return mk3(x * th / n, y * th / n, z * th / n);
}

}  // namespace qpc
