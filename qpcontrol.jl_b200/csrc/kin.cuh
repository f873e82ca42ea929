// kin.cuh -- kinematics / dynamics terms of one robot instance, one instance per CTA, tree state in shared memory.
//
// Device-side replacement for the RigidBodyDynamics state queries the reference's Parameters evaluate once per
// `solve!` (call sites reference src/tasks.jl:32-39,74-81,113-120,156-169,208-219,
// src/contacts.jl:58-61, src/lowlevel/momentum.jl:168-177, src/highlevel/standing.jl:60-76): transforms to root,
// world-frame motion subspaces, twists, bias accelerations, world / composite inertias, centre of mass, momentum,
// momentum-rate bias, centroidal momentum matrix, contact frames.  The sweeps are level-synchronous: bodies of one
// tree depth are processed by different threads, QPC_SYNC() between depths.
#pragma once
#include "qpc_common.h"
#include "qpc_program.h"

namespace qpc {

struct KinSmem {
  double *q, *v, *des, *cw, *cm;
  double *H, *TW, *BI, *IW, *IC, *SW;  // per body: 12, 6, 6, 10, 10 ; per velocity: 6
  double *scr;                          // nb*12 scratch (per-body momentum / Newton-Euler terms)
  double *tot;                          // com(3) momentum(6) hb(6) Wg(6)
  double *A;                            // 6 x nv world-frame momentum matrix
  double *GC;                           // ncontacts * N * 6 world wrench of each unit generator
  double *Jt, *bt;                      // 6 x nv task Jacobian scratch, 16 scalars
  double *tw;                           // QPC_MAXT scalar task weights of this tick
  const double* wm;                     // matrix weights of this tick (Wbuf layout): the per-tick Parameter or pg->Wbuf
  double *ct;                           // per contact (kin_ct_stride): z_up rotation 9, position 3, B 3N, B'B N^2, maxrho factor
};
// per contact: z_up rotation 9, position 3, force basis B 3N, maxrho factor 1 (B'B is formed where it is used)
QPC_HD int kin_ct_stride(int N) { return 13 + 3 * N; }

// Shared memory per instance bounds the resident instances per SM of the assembly kernel: the task-row buffer Jt shares the
// composite-inertia area (dead once the momentum matrix exists) and B'B is not staged -- 22.3 KB instead of 25.1 KB for
// Atlas, nine CTAs per SM instead of eight (measured: 7 -> 8 CTAs was worth 9 %, 8 -> 9 nothing).
QPC_HD int kin_ic_doubles(int nb, int nv) { return 10 * nb > 6 * nv ? 10 * nb : 6 * nv; }
QPC_HD int kin_smem_doubles(int nb, int nq, int nv, int ndes, int nc, int N) {
  const int na = nv > nc ? nv : nc;
  return nq + nv + ndes + 2 * nc + nb * (12 + 6 + 6 + 10) + kin_ic_doubles(nb, nv) + nv * 6 + nb * 12 + 24 + 6 * na +
         nc * N * 6 + 16 + QPC_MAXT + nc * kin_ct_stride(N);
}
QPC_HD KinSmem kin_layout(double* b, int nb, int nq, int nv, int ndes, int nc, int N) {
  KinSmem s;
  s.q = b;      b += nq;
  s.v = b;      b += nv;
  s.des = b;    b += ndes;
  s.cw = b;     b += nc;
  s.cm = b;     b += nc;
  s.H = b;      b += nb * 12;
  s.TW = b;     b += nb * 6;
  s.BI = b;     b += nb * 6;
  s.IW = b;     b += nb * 10;
  s.IC = b;     s.Jt = b;  b += kin_ic_doubles(nb, nv);  // Jt (6 nv) is first written after the last read of IC
  s.SW = b;     b += nv * 6;
  s.scr = b;    b += nb * 12;
  s.tot = b;    b += 24;
  s.A = b;      b += 6 * (nv > nc ? nv : nc);
  s.GC = b;     b += nc * N * 6;
  s.bt = b;     b += 16;
  s.tw = b;     b += QPC_MAXT;
  s.ct = b;     b += nc * kin_ct_stride(N);
  return s;
}

// ---- hand-off of the kinematic state from the assembly kernel to the inverse-dynamics kernel ---------------------------------
// The epilogue (kin_inverse_dynamics) needs SW, BI, IW, TW and the contact generator wrenches GC of the same state the
// assembly already swept; instead of repeating the forward sweep (half of the old inverse-dynamics kernel) the assembly
// kernel saves them (8.7 KB per Atlas instance, coalesced; 0.3 GB per 16,384-instance tick = 0.05 ms of HBM time) and the
// epilogue runs on a 12.7 KB shared-memory footprint instead of 25 KB -- twice the resident instances.
QPC_HD int kin_save_doubles(int nb, int nv, int nc, int N) { return 6 * nv + 22 * nb + 6 * nc * N; }
QPC_HD int kin_id_smem_doubles(int nb, int nv, int ndes, int nc, int N) {
  return ndes + kin_save_doubles(nb, nv, nc, N) + 12 * nb + 6 * nc + nv + 2;
}
QPC_HD KinSmem kin_id_layout(double* b, int nb, int nv, int ndes, int nc, int N) {
  KinSmem s;
  s.q = s.v = s.cw = s.cm = s.H = s.IC = s.tot = s.bt = s.tw = s.ct = nullptr;
  s.wm = nullptr;
  s.SW = b;     b += nv * 6;      // the saved block: SW | BI | IW | TW | GC, contiguous
  s.BI = b;     b += nb * 6;
  s.IW = b;     b += nb * 10;
  s.TW = b;     b += nb * 6;
  s.GC = b;     b += nc * N * 6;
  s.des = b;    b += ndes;
  s.scr = b;    b += nb * 12;
  s.A = b;      b += 6 * nc;
  s.Jt = b;     b += nv;
  return s;
}
// assembly side: cooperative, coalesced copy of the saved block (same order as kin_id_layout)
QPC_DEV void kin_save(const DevProgram* __restrict__ pg, const KinSmem& s, double* __restrict__ dst) {
  const int nb = pg->nb, nv = pg->nv, ng = pg->ncontacts * pg->N * 6;
  const int t = QPC_TID, nt = QPC_NT;
  for (int i = t; i < 6 * nv; i += nt) dst[i] = s.SW[i];
  dst += 6 * nv;
  for (int i = t; i < 6 * nb; i += nt) dst[i] = s.BI[i];
  dst += 6 * nb;
  for (int i = t; i < 10 * nb; i += nt) dst[i] = s.IW[i];
  dst += 10 * nb;
  for (int i = t; i < 6 * nb; i += nt) dst[i] = s.TW[i];
  dst += 6 * nb;
  for (int i = t; i < ng; i += nt) dst[i] = s.GC[i];
}
// epilogue side: the saved block and the desireds into shared memory
QPC_DEV void kin_id_load(const DevProgram* __restrict__ pg, KinSmem& s, const double* __restrict__ src,
                         const double* __restrict__ des) {
  const int tot = kin_save_doubles(pg->nb, pg->nv, pg->ncontacts, pg->N);
#if defined(__CUDA_ARCH__) && !defined(QPC_THREAD_PER_INSTANCE)
  // asynchronous copies global -> shared: every element of the 8.7 KB block in flight at once (a register-staged loop pays
  // one DRAM round trip per unrolled group, and this epilogue is all latency)
  for (int i = QPC_TID; i < tot; i += QPC_NT)
    asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"((unsigned)__cvta_generic_to_shared(s.SW + i)), "l"(src + i));
  for (int i = QPC_TID; i < pg->ndes; i += QPC_NT)
    asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"((unsigned)__cvta_generic_to_shared(s.des + i)), "l"(des + i));
  asm volatile("cp.async.commit_group;");
  asm volatile("cp.async.wait_group 0;" ::: "memory");
#else
  for (int i = QPC_TID; i < tot; i += QPC_NT) s.SW[i] = src[i];
  for (int i = QPC_TID; i < pg->ndes; i += QPC_NT) s.des[i] = des[i];
#endif
  QPC_SYNC();
}

QPC_DEV Xf body_to_root(const KinSmem& s, int body) { return body < 0 ? xf_identity() : xf_load(s.H + 12 * body); }
QPC_DEV S6 body_twist(const KinSmem& s, int body) { return body < 0 ? s6_zero() : ld6(s.TW + 6 * body); }
QPC_DEV S6 body_bias(const KinSmem& s, int body) { return body < 0 ? s6_zero() : ld6(s.BI + 6 * body); }

// load the instance inputs (q, v, desireds, contact parameters) into shared memory
QPC_DEV void kin_load(const DevProgram* __restrict__ pg, const BatchIO& io, long long inst, KinSmem& s) {
  const int t = QPC_TID, nt = QPC_NT;
  for (int i = t; i < pg->nq; i += nt) s.q[i] = io.q[inst * pg->nq + i];
  for (int i = t; i < pg->nv; i += nt) s.v[i] = io.v[inst * pg->nv + i];
  for (int i = t; i < pg->ndes; i += nt)
    s.des[i] = io.desired ? io.desired[inst * io.desired_stride + i] : pg->def_desired[i];
  for (int i = t; i < pg->ncontacts; i += nt) {
    s.cw[i] = io.cweight ? io.cweight[inst * io.contact_stride + i] : pg->def_cweight[i];
    s.cm[i] = io.cmaxnf ? io.cmaxnf[inst * io.contact_stride + i] : pg->def_cmaxnf[i];
  }
  for (int i = t; i < pg->ntasks; i += nt)
    s.tw[i] = io.tweight ? io.tweight[inst * io.tweight_stride + i] : pg->tasks[i].weight;
  s.wm = io.twmat ? io.twmat + inst * io.twmat_stride : pg->Wbuf;
  // contact table of this tick: the setup-time ContactPoint geometry, or the per-tick Parameters (contacts.jl:53-61)
  const int N = pg->N, cts = kin_ct_stride(N);
  for (int c = t; c < pg->ncontacts; c += nt) {
    const DevContact& dc = pg->contacts[c];
    double* ct = s.ct + c * cts;
    if (!io.cgeom) {
      for (int i = 0; i < 9; i++) ct[i] = dc.Rz[i];
      for (int i = 0; i < 3; i++) ct[9 + i] = dc.pos[i];
      for (int i = 0; i < 3 * N; i++) ct[12 + i] = dc.B[i];
      ct[12 + 3 * N] = dc.maxrho_factor;
    } else {
      const double* gq = io.cgeom + inst * io.cgeom_stride + 7 * c;
      const double mu = gq[6];
      rot_between_z(ld3(gq + 3), ct);
      for (int i = 0; i < 3; i++) ct[9 + i] = gq[i];
      double* B = ct + 12;
      for (int g = 0; g < N; g++) {  // forcebasis (contacts.jl:16-23)
        const double th = g * (6.283185307179586476925 / N);
        const V3 b = mk3(mu * cos(th), mu * sin(th), 1.0);
        const double inv = 1.0 / sqrt(dot(b, b));
        B[g] = b.x * inv;
        B[N + g] = b.y * inv;
        B[2 * N + g] = b.z * inv;
      }
      ct[12 + 3 * N] = 1.0 / (N * sqrt(mu * mu + 1.0));  // contacts.jl:57
    }
  }
  QPC_SYNC();
}

// forward sweep: transform_to_root, motion subspaces, twist_wrt_world, bias_acceleration, world inertias.
// Only what depends on the parent is done level by level (one 3x3 product for the transform, one cross product for the
// bias): the joint transforms (sincos), the motion subspaces, the world inertias and the Newton-Euler terms are computed
// for ALL bodies at once between the two level sweeps.  In the first version every body did all of it inside its level
// step -- 8 levels x ~400 dependent flops with 60 of the 64 threads waiting at the barrier (38 % of the kernel's stalls).
QPC_DEV void kin_forward(const DevProgram* __restrict__ pg, KinSmem& s) {
  const int nb = pg->nb;
  double* loc = s.scr;  // nb * 12 local transforms (scr is free until phase 5)
  double* jtwp = s.IC;  // joint twists, 6 of every body's 10 slots (IC is free until kin_composite)
  // phase 1 (all bodies): local transform X_tree * X_joint(q)
  for (int b = QPC_TID; b < nb; b += QPC_NT) {
    const int jt = pg->jtype[b];
    const double* qj = s.q + pg->qoff[b];
    Xf Xj = xf_identity();
    const V3 ax = ld3(pg->axis + 3 * b);
    if (jt == 0) {
      axis_angle_to_rot(ax, qj[0], Xj.R);
    } else if (jt == 1) {
      Xj.p = qj[0] * ax;
    } else if (jt == 2) {
      double nn = 1.0 / sqrt(qj[0] * qj[0] + qj[1] * qj[1] + qj[2] * qj[2] + qj[3] * qj[3]);
      quat_to_rot(qj[0] * nn, qj[1] * nn, qj[2] * nn, qj[3] * nn, Xj.R);
      Xj.p = mk3(qj[4], qj[5], qj[6]);
    }
    Xf Xt;
    for (int i = 0; i < 9; i++) Xt.R[i] = pg->XR[9 * b + i];
    Xt.p = ld3(pg->Xp + 3 * b);
    xf_store(loc + 12 * b, xf_mul(Xt, Xj));
  }
  QPC_SYNC();
  // phase 2 (all bodies, NO barrier between tree levels): every body composes its own ancestor chain, root first -- the
  // products of a level-by-level sweep in the same order, so the same bits -- then its world-frame motion subspaces, its
  // joint twist and its world inertia.  A warp runs the chain loop once per level either way; what goes away is one CTA
  // barrier per level (Atlas: 11 levels; barriers were 37 % of this kernel's stall samples, profiles/r2_asm_v2_*).
  for (int b = QPC_TID; b < nb; b += QPC_NT) {
    const int a0 = pg->anc_ptr[b], a1 = pg->anc_ptr[b + 1];
    Xf H = xf_load(loc + 12 * pg->anc_idx[a0]);
    for (int k = a0 + 1; k < a1; k++) H = xf_mul(H, xf_load(loc + 12 * pg->anc_idx[k]));
    xf_store(s.H + 12 * b, H);
    const int jt = pg->jtype[b];
    const V3 ax = ld3(pg->axis + 3 * b);
    const int o = pg->voff[b];
    S6 jtw = s6_zero();
    if (jt == 0) {
      S6 S = xmotion(H, mk6(ax, mk3(0, 0, 0)));
      st6(s.SW + 6 * o, S);
      jtw = s.v[o] * S;
    } else if (jt == 1) {
      S6 S = xmotion(H, mk6(mk3(0, 0, 0), ax));
      st6(s.SW + 6 * o, S);
      jtw = s.v[o] * S;
    } else if (jt == 2) {
      for (int c = 0; c < 3; c++) {
        V3 e = mk3(c == 0, c == 1, c == 2);
        S6 Sa = xmotion(H, mk6(e, mk3(0, 0, 0)));
        S6 Sl = xmotion(H, mk6(mk3(0, 0, 0), e));
        st6(s.SW + 6 * (o + c), Sa);
        st6(s.SW + 6 * (o + 3 + c), Sl);
        jtw = jtw + s.v[o + c] * Sa + s.v[o + 3 + c] * Sl;
      }
    }
    st6(jtwp + 10 * b, jtw);
    si_store(s.IW + 10 * b, si_transform(H, si_load(pg->inertia + 10 * b)));
  }
  QPC_SYNC();
  // phase 3 (all bodies): twist = sum of the chain's joint twists, bias = sum of (twist x joint twist) along the chain,
  // root first; then the body's momentum and Newton-Euler bias terms (summed in kin_composite)
  for (int b = QPC_TID; b < nb; b += QPC_NT) {
    S6 tw = s6_zero(), bi = s6_zero();
    for (int k = pg->anc_ptr[b]; k < pg->anc_ptr[b + 1]; k++) {
      const S6 jtw = ld6(jtwp + 10 * pg->anc_idx[k]);
      tw = tw + jtw;
      bi = bi + cross_motion(tw, jtw);
    }
    st6(s.TW + 6 * b, tw);
    st6(s.BI + 6 * b, bi);
    const SI Iw = si_load(s.IW + 10 * b);
    st6(s.scr + 12 * b, si_mul(Iw, tw));
    st6(s.scr + 12 * b + 6, newton_euler(Iw, bi, tw));
  }
  QPC_SYNC();
}

// composite rigid-body inertias (every body sums its own subtree, deepest descendants first: one pass, no barrier per
// level), centre of mass, momentum, momentum_rate_bias, gravity wrench, world-frame momentum matrix
QPC_DEV void kin_composite(const DevProgram* __restrict__ pg, KinSmem& s) {
  for (int b = QPC_TID; b < pg->nb; b += QPC_NT) {
    double acc[10];
    for (int i = 0; i < 10; i++) acc[i] = 0.0;
    for (int k = pg->desc_ptr[b]; k < pg->desc_ptr[b + 1]; k++) {
      const int d = pg->desc_idx[k];
      for (int i = 0; i < 10; i++) acc[i] += s.IW[10 * d + i];
    }
    for (int i = 0; i < 10; i++) s.IC[10 * b + i] = acc[i] + s.IW[10 * b + i];
  }
  QPC_SYNC();
  // totals: 12 sums over bodies + centre of mass from the roots' composite inertia
  for (int c = QPC_TID; c < 12; c += QPC_NT) {
    double a = 0;
    for (int b = 0; b < pg->nb; b++) a += s.scr[12 * b + c];
    s.tot[3 + c] = a;  // momentum (3..8), momentum_rate_bias (9..14)
  }
  for (int c = QPC_TID; c < 3; c += QPC_NT) {
    double a = 0;
    for (int b = 0; b < pg->nb; b++)
      if (pg->parent[b] < 0) a += s.IC[10 * b + 6 + c];
    s.tot[c] = a / pg->total_mass;
  }
  // momentum matrix column c = I^c_{succ(joint)} S_c (world frame)   (momentum_matrix!, momentum.jl:168)
  for (int c = QPC_TID; c < pg->nv; c += QPC_NT) {
    S6 col = si_mul(si_load(s.IC + 10 * pg->vbody[c]), ld6(s.SW + 6 * c));
    for (int r = 0; r < 6; r++) s.A[r * pg->nv + c] = s6_get(col, r);
  }
  QPC_SYNC();
  if (QPC_TID == 0) {
    V3 fg = pg->total_mass * ld3(pg->gravity);
    st6(s.tot + 15, mk6(cross(ld3(s.tot), fg), fg));  // Wg = (c x m g, m g)  (momentum.jl:170)
  }
  QPC_SYNC();
}

// StandingController PD laws (standing.jl:60-85) -> desireds of the tasks it owns
QPC_DEV void kin_standing(const DevProgram* __restrict__ pg, KinSmem& s) {
  if (!pg->standing) return;
  const double mass = pg->total_mass;
  for (int k = QPC_TID; k < pg->st_nj + 2; k += QPC_NT) {
    if (k == pg->st_nj) {  // centre of mass: m * (-k (c - cref) - d hlin / m)
      for (int c = 0; c < 3; c++) {
        double e = s.tot[c] - pg->st_comref[c];
        double ed = s.tot[3 + 3 + c] / mass;
        s.des[pg->st_linmom_des + c] = mass * (-pg->st_com_kp * e - pg->st_com_kd * ed);
      }
    } else if (k == pg->st_nj + 1) {  // pelvis orientation: -k rotvec(R) - d omega_body
      Xf H = body_to_root(s, pg->st_pelvis_body);
      S6 T = xmotion(xf_inv(H), body_twist(s, pg->st_pelvis_body));
      V3 rv = rot_to_rotvec(H.R);
      V3 wd = (-pg->st_pelvis_kp) * rv + (-pg->st_pelvis_kd) * T.w;
      st3(s.des + pg->st_pelvis_des, wd);
    } else {
      s.des[pg->st_jdes[k]] = -pg->st_kp[k] * (s.q[pg->st_jq[k]] - pg->st_ref[k]) - pg->st_kd[k] * s.v[pg->st_jv[k]];
    }
  }
  QPC_SYNC();
}

// ---- SE3PDController on the device (reference src/lowlevel/se3pdcontroller.jl:13-18) ---------------------------------------
// One `Interpolated` / `Piecewise` trajectory with two derivatives (interpolated.jl:19-60, piecewise.jl:25-40).
// rotation: y[4] quaternion, d1 / d2 angular velocity / acceleration (Lie derivative, interpolated.jl:75-82);
// vector: y[0..2], d1, d2.
QPC_DEV void traj_eval(const DevTraj& tr, double x, bool rotation, double* y, V3& d1, V3& d2) {
  int k = 0;
  if (tr.piecewise) {
    const double xc = fmin(fmax(x, tr.seg[0].brk), tr.brk_end);
    for (int i = 1; i < tr.nseg; i++)
      if (tr.seg[i].brk <= xc) k = i;
    x = xc - tr.seg[k].brk;
  }
  const DevInterp& g = tr.seg[k];
  const double dx = g.xf - g.x0;
  double th = (x - g.x0) / dx, dth = 1.0 / dx;
  if (th <= 0.0) {
    th = 0.0;
    dth = 0.0;
  } else if (th >= 1.0) {
    th = 1.0;
    dth = 0.0;
  }
  double al = th, al1 = 1.0, al2 = 0.0;
  if (g.nc > 0) {  // Horner on the polynomial and its first two derivatives
    al = al1 = al2 = 0.0;
    for (int i = g.nc - 1; i >= 0; i--) {
      al2 = al2 * th + 2.0 * al1;
      al1 = al1 * th + al;
      al = al * th + g.c[i];
    }
  }
  const double a1 = al1 * dth, a2 = al2 * dth * dth;
  const V3 ax = ld3(g.dy);
  if (rotation) {
    double sh, ch;
    sincos(0.5 * al * g.angle, &sh, &ch);
    const double w1 = g.y0[0], x1 = g.y0[1], y1 = g.y0[2], z1 = g.y0[3];
    const double x2 = sh * ax.x, y2 = sh * ax.y, z2 = sh * ax.z;
    y[0] = w1 * ch - x1 * x2 - y1 * y2 - z1 * z2;
    y[1] = w1 * x2 + x1 * ch + y1 * z2 - z1 * y2;
    y[2] = w1 * y2 - x1 * z2 + y1 * ch + z1 * x2;
    y[3] = w1 * z2 + x1 * y2 - y1 * x2 + z1 * ch;
    d1 = (a1 * g.angle) * ax;
    d2 = (a2 * g.angle) * ax;
  } else {
    st3(y, ld3(g.y0) + al * ax);
    d1 = a1 * ax;
    d2 = a2 * ax;
  }
}
QPC_DEV V3 mat3_mul(const double* M, V3 v) { return rot(M, v); }

// Desired spatial acceleration of every SE3PDController at time t: Tdref + pd(gains, H, Href, T, Tref), double-geodesic
// PD in the body frame (RigidBodyDynamics.PDControl, restated in qpcontrol.jl_b200/se3pd.py: pd_se3), written over the
// desired of the SpatialAccelerationTask the controller drives.  Runs after kin_forward (transforms and twists).
// (not inlined, scalar arguments only: keeps its registers out of the assembly kernels, which run it for few programs)
static QPC_DEVN void kin_se3pd_eval(const DevProgram* __restrict__ pg, double t, const double* Hs, const double* TWs, double* des) {
  KinSmem s;
  s.H = const_cast<double*>(Hs);
  s.TW = const_cast<double*>(TWs);
  s.des = des;
  for (int c = QPC_TID; c < pg->nse3; c += QPC_NT) {
    const DevSE3PD& sc = pg->se3[c];
    // reference pose, twist and spatial acceleration in the desired body frame (se3.jl:8-27)
    double quat[4], pdes[3], Rd[9];
    V3 w_base, wd_base, pd1, pd2;
    traj_eval(sc.ang, t, true, quat, w_base, wd_base);
    traj_eval(sc.lin, t, false, pdes, pd1, pd2);
    quat_to_rot(quat[0], quat[1], quat[2], quat[3], Rd);
    const V3 w_des = rot_t(Rd, w_base), nu_des = rot_t(Rd, pd1);
    const V3 wd_des = rot_t(Rd, wd_base), nud_des = rot_t(Rd, pd2) + cross(w_des, nu_des);
    // actual pose of body in base and relative twist in the body frame (se3pdcontroller.jl:15-16)
    const Xf Hb = body_to_root(s, sc.body), Ha = body_to_root(s, sc.base);
    const Xf H = xf_mul(xf_inv(Ha), Hb);
    const S6 T = xmotion(xf_inv(Hb), body_twist(s, sc.body) - body_twist(s, sc.base));
    // error pose e = inv(x_des) * x and the reference twist re-expressed in the body frame
    double Re[9];
    for (int i = 0; i < 3; i++)
      for (int j = 0; j < 3; j++) Re[3 * i + j] = Rd[i] * H.R[j] + Rd[3 + i] * H.R[3 + j] + Rd[6 + i] * H.R[6 + j];
    const V3 pe = rot_t(Rd, H.p - ld3(pdes));
    const V3 pe_b = rot_t(Re, pe);
    const V3 w_des_b = rot_t(Re, w_des);
    const V3 nu_des_b = rot_t(Re, nu_des) + cross(-pe_b, w_des_b);
    const V3 ang = -mat3_mul(sc.K, rot_to_rotvec(Re)) - mat3_mul(sc.K + 9, T.w - w_des_b);
    const V3 lin = -mat3_mul(sc.K + 18, pe_b) - mat3_mul(sc.K + 27, T.v - nu_des_b);
    st3(s.des + sc.des_off, wd_des + ang);
    st3(s.des + sc.des_off + 3, nud_des + lin);
  }
}
QPC_DEV void kin_se3pd(const DevProgram* __restrict__ pg, const BatchIO& io, long long inst, KinSmem& s) {
  if (!pg->nse3) return;
  kin_se3pd_eval(pg, (io.time ? io.time[inst * io.time_stride] : 0.0) + io.time_offset, s.H, s.TW, s.des);
  QPC_SYNC();
}

// contact frames and unit-generator wrenches: toroot = transform_to_root(body) * z_up (contacts.jl:61);
// generator g of contact c produces the world wrench (p x (R b_g); R b_g)  (contacts.jl:63,66,67)
QPC_DEV void kin_contacts(const DevProgram* __restrict__ pg, KinSmem& s) {
  const int N = pg->N;
  for (int k = QPC_TID; k < pg->ncontacts * N; k += QPC_NT) {
    const int c = k / N, g = k % N;
    const DevContact& dc = pg->contacts[c];
    const double* ct = s.ct + c * kin_ct_stride(N);
    Xf Z;
    for (int i = 0; i < 9; i++) Z.R[i] = ct[i];
    Z.p = ld3(ct + 9);
    Xf T = xf_mul(body_to_root(s, dc.body), Z);
    const double* B = ct + 12;
    V3 f = rot(T.R, mk3(B[g], B[N + g], B[2 * N + g]));
    st6(s.GC + 6 * k, mk6(cross(T.p, f), f));
  }
  QPC_SYNC();
}

// rows of task_error = J vd + b - desired for one task into Jt (dim x nv) and bt (dim)
// (tasks.jl:31-44,73-84,112-123,153-171,185-189,230-236,258-262).  One velocity COLUMN per thread (every entry of Jt is
// written, zeros included: no clearing pass, and the six columns of a floating joint go to six threads); the bias by the
// first thread past the columns.  No barrier inside: the caller separates producing and consuming a buffer.
QPC_DEV void kin_task_rows(const DevProgram* __restrict__ pg, const KinSmem& s, int ti, double* Jt, double* bt) {
  const DevTask& t = pg->tasks[ti];
  const int nv = pg->nv, kind = t.kind, dim = t.dim;
  const int bias_tid = QPC_NT > nv ? nv : 0;
  if (kind <= 3) {  // path tasks: geometric Jacobian in `frame` (point task: the base frame)
    const Xf Xinv = xf_inv(body_to_root(s, t.frame));
    V3 p = mk3(0, 0, 0);
    if (kind == 3) {
      Xf rel = xf_mul(Xinv, body_to_root(s, t.target));
      p = rot(rel.R, ld3(t.point)) + rel.p;
    }
    const int r0 = kind == 2 ? 3 : 0;
    for (int c = QPC_TID; c < nv; c += QPC_NT) {
      const int sg = pg->tsign[ti][c];
      if (sg == 0) {
        for (int r = 0; r < dim; r++) Jt[r * nv + c] = 0.0;
        continue;
      }
      S6 col = (double)sg * xmotion(Xinv, ld6(s.SW + 6 * c));
      if (kind == 3) {
        V3 pj = cross(col.w, p) + col.v;  // point_jacobian!: J_lin + J_ang x p
        Jt[0 * nv + c] = pj.x;
        Jt[1 * nv + c] = pj.y;
        Jt[2 * nv + c] = pj.z;
      } else {
        for (int r = 0; r < dim; r++) Jt[r * nv + c] = s6_get(col, r0 + r);
      }
    }
    if (QPC_TID == bias_tid) {
      // transform(state, -bias(source) + bias(target), frame): X [ a + (-T_frame) x (T_target - T_source) ]
      S6 rel = body_twist(s, t.target) - body_twist(s, t.source);
      S6 a = body_bias(s, t.target) - body_bias(s, t.source);
      S6 neg = (-1.0) * body_twist(s, t.frame);
      S6 jv = xmotion(Xinv, a + cross_motion(neg, rel));
      if (kind == 3) {
        S6 T = xmotion(Xinv, rel);
        V3 pd = cross(T.w, p) + T.v;
        V3 bb = cross(T.w, pd) + cross(jv.w, p) + jv.v;
        st3(bt, bb);
      } else {
        for (int r = 0; r < dim; r++) bt[r] = s6_get(jv, r0 + r);
      }
    }
  } else if (kind == 4) {
    const int o = pg->voff[t.joint];
    for (int c = QPC_TID; c < nv; c += QPC_NT)
      for (int r = 0; r < dim; r++) Jt[r * nv + c] = (c == o + r) ? 1.0 : 0.0;
    if (QPC_TID == bias_tid)
      for (int r = 0; r < dim; r++) bt[r] = 0.0;
  } else {  // momentum-rate tasks: force-transform of the world momentum matrix to the centroidal frame
    const V3 com = ld3(s.tot);
    const int r0 = kind == 6 ? 3 : 0;
    for (int c = QPC_TID; c < nv; c += QPC_NT) {
      V3 aw = mk3(s.A[0 * nv + c], s.A[1 * nv + c], s.A[2 * nv + c]);
      V3 al = mk3(s.A[3 * nv + c], s.A[4 * nv + c], s.A[5 * nv + c]);
      S6 col = mk6(aw - cross(com, al), al);
      for (int r = 0; r < dim; r++) Jt[r * nv + c] = s6_get(col, r0 + r);
    }
    if (QPC_TID == bias_tid) {
      S6 hb = ld6(s.tot + 9);
      S6 hc = mk6(hb.w - cross(com, hb.v), hb.v);
      for (int r = 0; r < dim; r++) bt[r] = s6_get(hc, r0 + r);
    }
  }
}
// one task's rows into the QP from a finished (Jt, bt) buffer: residual r = b - desired + J_fixed vd_fixed, bounds, the free
// columns of G, slack columns and slack cost
QPC_DEV void kin_task_emit(const DevProgram* __restrict__ pg, const KinSmem& s, int ti, const double* Jt, const double* bt,
                           double* P, double* G, double* lg, double* ug) {
  const DevTask& t = pg->tasks[ti];
  const int n = pg->n, nv = pg->nv, dim = t.dim, t0 = QPC_TID, nt = QPC_NT;
  // the residual rows go to the LAST threads (the first ones are still busy with the next task's columns)
  for (int k = nt - 1 - t0; k < dim; k += nt) {
    double a = bt[k] - s.des[t.des_off + k];
    for (int j = 0; j < pg->nfixv; j++) {
      const int i = pg->fixv[j];
      a += Jt[k * nv + i] * s.des[pg->vfix_des[i]];
    }
    lg[t.row0 + k] = ug[t.row0 + k] = -a;
  }
  for (int k = t0; k < dim * nv; k += nt) {
    const int row = k / nv, i = k % nv;
    if (pg->vcol[i] >= 0) G[(t.row0 + row) * n + pg->vcol[i]] = Jt[k];
  }
  if (t.mode != 0) {
    for (int k = t0; k < dim; k += nt) G[(t.row0 + k) * n + t.scol0 + k] = -1.0;
    if (t.mode == 1) {
      for (int k = t0; k < dim; k += nt) P[(t.scol0 + k) * n + t.scol0 + k] = 2.0 * s.tw[ti];
    } else {  // e'We with a possibly unsymmetric W: the Hessian is W + W'
      const double* W = s.wm + t.w_off;
      for (int k = t0; k < dim * dim; k += nt) {
        const int a = k / dim, b = k % dim;
        P[(t.scol0 + a) * n + t.scol0 + b] = W[a * dim + b] + W[b * dim + a];
      }
    }
  }
}

// ---- QP assembly -------------------------------------------------------------------------------------------------------
// x = (free vd, task-error slacks e, rho).  Cost: regularisation 2 reg on vd (momentum.jl:128-131), the weighted
// tasks' w e'e / e'We on their slacks (momentum.jl:107-117), 2 w_c B'B on each contact's rho block (contacts.jl:75-79
// with f = B rho substituted).  General rows (all equalities): hard tasks J vd = -r; weighted tasks J vd - e = -r
// (momentum.jl:124); wrench balance S'(A vd - sum G_c rho_c) = S'(Wg - Adot v) (momentum.jl:162-193).
// Box rows 0 <= rho <= maxrho (contacts.jl:64-65).  r = b - desired + J_fixed vd_fixed accounts for velocities fixed
// by hard JointAccelerationTasks, which are substituted out.  Contact force / wrench variables of the reference's
// lifted QP (contacts.jl:46-48,63,66,67) are eliminated through their defining equalities.
// P, qv, G, lg, ug, lb, ub point at this instance's slots (global or shared memory).
// prezeroed: P and G already hold zeros wherever this function never writes (the controller's own workspace, zeroed at
// allocation: every tick writes the same positions, so 4,081 of the 4,300 stores per Atlas instance were clearing zeros).
QPC_DEV void kin_assemble(const DevProgram* __restrict__ pg, KinSmem& s, double* P, double* qv, double* G, double* lg,
                          double* ug, double* lb, double* ub, bool prezeroed = false) {
  const int n = pg->n, nv = pg->nv, mg = pg->mg, N = pg->N;
  const int t0 = QPC_TID, nt = QPC_NT;
  if (!prezeroed) {
    for (int i = t0; i < n * n; i += nt) P[i] = 0.0;
    for (int i = t0; i < mg * n; i += nt) G[i] = 0.0;
  }
  for (int i = t0; i < n; i += nt) qv[i] = 0.0;
  QPC_SYNC();
  // regularisation and contact-force cost
  for (int i = t0; i < nv; i += nt)
    if (pg->vcol[i] >= 0) P[pg->vcol[i] * n + pg->vcol[i]] = 2.0 * pg->reg[i];
  for (int k = t0; k < pg->ncontacts * N * N; k += nt) {
    const int c = k / (N * N), a = (k / N) % N, b = k % N;
    const DevContact& dc = pg->contacts[c];
    const double* Bc = s.ct + c * kin_ct_stride(N) + 12;  // B'B of the contact's force basis (contacts.jl:75-79)
    P[(dc.col0 + a) * n + dc.col0 + b] = 2.0 * s.cw[c] * (Bc[a] * Bc[b] + Bc[N + a] * Bc[N + b] + Bc[2 * N + a] * Bc[2 * N + b]);
  }
  for (int k = t0; k < pg->ncontacts * N; k += nt) {
    const int c = k / N;
    lb[k] = 0.0;
    ub[k] = s.cm[c] * s.ct[c * kin_ct_stride(N) + 12 + 3 * N];
  }
  QPC_SYNC();
  // Task rows, software-pipelined over two (Jt, bt) buffers: in one barrier interval the threads produce task j's rows
  // (one column each) and emit task j-1's into the QP -- one barrier per task instead of four (the per-task loop was 25 %
  // of this kernel: profiles/r2_asm_v2_stalls_by_line.txt).  The second buffer is the dead per-body scratch when it is
  // large enough, otherwise both steps use s.Jt with a barrier in between.
  {
    const bool two = 12 * pg->nb >= 6 * nv;
    double *jc = s.Jt, *jp = s.scr, *bc = s.bt, *bp = s.bt + 8;  // buffer being produced / being emitted
    int prev = -1;
    for (int ia = 0; ia < pg->nactive; ia++) {
      const int ti = pg->active[ia];
      if (!two) {
        kin_task_rows(pg, s, ti, jc, bc);
        QPC_SYNC();
        kin_task_emit(pg, s, ti, jc, bc, P, G, lg, ug);
        QPC_SYNC();
        continue;
      }
      kin_task_rows(pg, s, ti, jc, bc);
      if (prev >= 0) kin_task_emit(pg, s, prev, jp, bp, P, G, lg, ug);
      QPC_SYNC();
      prev = ti;
      double* tj = jc;
      jc = jp;
      jp = tj;
      double* tb = bc;
      bc = bp;
      bp = tb;
    }
    if (prev >= 0) kin_task_emit(pg, s, prev, jp, bp, P, G, lg, ug);
    QPC_SYNC();
  }
  if (pg->floating >= 0) {  // add_wrench_balance_constraint! (momentum.jl:162-193)
    const int o = pg->voff[pg->floating], row0 = pg->balance_row0;
    // SA = S'A (6 x nv) into Jt
    for (int k = t0; k < 6 * nv; k += nt) {
      const int kk = k / nv, c = k % nv;
      S6 S = ld6(s.SW + 6 * (o + kk));
      s.Jt[k] = S.w.x * s.A[c] + S.w.y * s.A[nv + c] + S.w.z * s.A[2 * nv + c] + S.v.x * s.A[3 * nv + c] +
                S.v.y * s.A[4 * nv + c] + S.v.z * s.A[5 * nv + c];
    }
    QPC_SYNC();
    for (int k = t0; k < 6 * nv; k += nt) {
      const int kk = k / nv, i = k % nv;
      if (pg->vcol[i] >= 0) G[(row0 + kk) * n + pg->vcol[i]] = s.Jt[k];
    }
    for (int k = t0; k < 6 * pg->ncontacts * N; k += nt) {
      const int kk = k / (pg->ncontacts * N), cg = k % (pg->ncontacts * N);
      G[(row0 + kk) * n + pg->contacts[cg / N].col0 + cg % N] = -dot(ld6(s.SW + 6 * (o + kk)), ld6(s.GC + 6 * cg));
    }
    for (int kk = t0; kk < 6; kk += nt) {
      double a = dot(ld6(s.SW + 6 * (o + kk)), ld6(s.tot + 15) - ld6(s.tot + 9));
      for (int i = 0; i < nv; i++)
        if (pg->vcol[i] < 0) a -= s.Jt[kk * nv + i] * s.des[pg->vfix_des[i]];
      lg[row0 + kk] = ug[row0 + kk] = a;
    }
    QPC_SYNC();
  }
}

// RNEA core (RigidBodyDynamics.inverse_dynamics!, momentum.jl:75): vd [nv], ext [nb][6] external wrench per body in world
// frame (overwritten with the joint wrenches), acc [nb][6] scratch.  zero_floating: zero_floating_joint_torques!
// (momentum.jl:93-97) -- the forward dynamics below needs those rows.
QPC_DEV void kin_rnea(const DevProgram* __restrict__ pg, KinSmem& s, const double* vd, double* ext, double* acc,
                      double* tau_out, bool zero_floating) {
  const int nv = pg->nv, nb = pg->nb;
  // forward, as a chain walk (no barrier per tree level): with gravity as the root acceleration,
  //   a_b = a_root + bias_b + sum over the ancestor chain of vd_c S_c
  // (the bias differences of a level-by-level sweep telescope), then the body's own wrench  I a + T x* I T - ext  into `acc`
  const S6 a0 = mk6(mk3(0, 0, 0), -ld3(pg->gravity));
  for (int b = QPC_TID; b < nb; b += QPC_NT) {
    S6 a = a0;
    for (int k = pg->anc_ptr[b]; k < pg->anc_ptr[b + 1]; k++) {
      const int ab = pg->anc_idx[k];
      for (int c = pg->voff[ab]; c < pg->voff[ab] + pg->nvj[ab]; c++) a = a + vd[c] * ld6(s.SW + 6 * c);
    }
    a = a + body_bias(s, b);
    st6(acc + 6 * b, newton_euler(si_load(s.IW + 10 * b), a, ld6(s.TW + 6 * b)) - ld6(ext + 6 * b));
  }
  QPC_SYNC();
  // backward: the joint wrench of body b is the sum of the own wrenches over its subtree (deepest descendants first)
  for (int b = QPC_TID; b < nb; b += QPC_NT) {
    S6 w = s6_zero();
    for (int k = pg->desc_ptr[b]; k < pg->desc_ptr[b + 1]; k++) w = w + ld6(acc + 6 * pg->desc_idx[k]);
    st6(ext + 6 * b, w + ld6(acc + 6 * b));
  }
  QPC_SYNC();
  for (int c = QPC_TID; c < nv; c += QPC_NT) {
    const int b = pg->vbody[c];
    double tau = dot(ld6(s.SW + 6 * c), ld6(ext + 6 * b));
    if (zero_floating && b == pg->floating) tau = 0.0;  // zero_floating_joint_torques! (momentum.jl:93-97)
    tau_out[c] = tau;
  }
  QPC_SYNC();
}

// ---- epilogue: contact wrenches, inverse dynamics (momentum.jl:62-80) ------------------------------------------------
// x = condensed solution.  Writes vd (nv), per-contact world wrenches (nc x 6) and tau (nv) to the given slots.
QPC_DEV void kin_inverse_dynamics(const DevProgram* __restrict__ pg, KinSmem& s, const double* x, double* vd_out,
                                  double* wrench_out, double* tau_out) {
  const int nv = pg->nv, nb = pg->nb, N = pg->N;
  double* vd = s.Jt;            // nv
  double* ext = s.scr;          // nb * 6 external wrench per body (then joint wrenches)
  double* acc = s.scr + 6 * nb; // nb * 6 spatial accelerations
  for (int i = QPC_TID; i < nv; i += QPC_NT) {
    const int ci = pg->vcol[i];
    vd[i] = ci >= 0 ? x[ci] : s.des[pg->vfix_des[i]];
    if (vd_out) vd_out[i] = vd[i];
  }
  for (int i = QPC_TID; i < 6 * nb; i += QPC_NT) ext[i] = 0.0;
  QPC_SYNC();
  // value(model, point.wrench_world) = sum_g rho_g * generator wrench; summed per body (momentum.jl:65-72)
  for (int k = QPC_TID; k < pg->ncontacts * 6; k += QPC_NT) {
    const int c = k / 6, r = k % 6;
    double a = 0;
    for (int g = 0; g < N; g++) a += x[pg->contacts[c].col0 + g] * s.GC[6 * (c * N + g) + r];
    if (wrench_out) wrench_out[k] = a;
    s.A[k] = a;  // A is free now: stash per-contact wrenches
  }
  QPC_SYNC();
  for (int k = QPC_TID; k < nb * 6; k += QPC_NT) {
    const int b = k / 6, r = k % 6;
    double a = 0;
    for (int c = 0; c < pg->ncontacts; c++)
      if (pg->contacts[c].body == b) a += s.A[6 * c + r];
    ext[k] = a;
  }
  QPC_SYNC();
  kin_rnea(pg, s, vd, ext, acc, tau_out, true);
}


// ---- forward dynamics under a soft ground contact: the plant of the closed loop ---------------------
// notebooks/Standing controller.ipynb:202-214 runs simulate(state, T, PeriodicController(tau, dt, controller)): the
// simulator applies the commanded torques to the mechanism and the environment answers with contact forces.  Here:
//   vd = M(q)^-1 (tau - c(q, v, w_ext)),   M by the composite-rigid-body algorithm, c by RNEA at vd = 0 with the contact
//   wrenches as external wrenches, the solve by Cholesky in shared memory;
//   contact: every ContactPoint of the controller against the half-space z >= ground_z, normal force
//   max(0, -k phi - d phidot) on penetration phi < 0 (spring-damper), tangential force -mu f_n v_t / max(|v_t|, v_eps)
//   (Coulomb friction regularised below v_eps).
// Restated from the published algorithms of RigidBodyDynamics.dynamics! / RigidBodySim [dep-memory]; the soft-contact
// model stands for RigidBodyDynamics.Contact's (Hunt-Crossley normal + viscoelastic Coulomb friction with state), of
// which it keeps the half-space geometry, the penalty normal force and the friction cone, not the friction state.
struct ContactModel {
  double k, d, mu, v_eps, ground_z;
};
QPC_HD int kin_fd_extra_doubles(int nv) { return nv * nv + 2 * nv; }  // M, rhs / tau_bias
// s: full KinSmem (kin_load + kin_forward + kin_composite done); M [nv*nv], rhs [nv], tb [nv] scratch; tau [nv] applied
// torques (global or shared); vd_out [nv]; fc_out optional [ncontacts][3] world contact forces.
// anchor: optional per-contact state [ncontacts][3] = (x, y, in-contact flag) of the tangential spring (stick): with it the
// tangential force is -k (p_t - anchor) - d v_t clipped to the cone mu f_n (the anchor slides along when clipped), the
// viscoelastic Coulomb model; without it the force is the regularised Coulomb law, whose stick regime is a damper of
// coefficient mu f_n / v_eps that an explicit integrator only tolerates for light loads.
QPC_DEV void kin_forward_dynamics(const DevProgram* __restrict__ pg, KinSmem& s, const ContactModel& cm,
                                  const double* tau, double* M, double* rhs, double* tb, double* vd_out, double* fc_out,
                                  double* anchor = nullptr) {
  const int nv = pg->nv, nb = pg->nb;
  double* ext = s.scr;
  double* acc = s.scr + 6 * nb;
  for (int i = QPC_TID; i < 6 * nb; i += QPC_NT) ext[i] = 0.0;
  for (int i = QPC_TID; i < nv; i += QPC_NT) rhs[i] = 0.0;  // vd = 0 for the bias pass
  QPC_SYNC();
  // soft ground contact at the controller's contact points (world frame); one thread per point, then per-body sums
  for (int c = QPC_TID; c < pg->ncontacts; c += QPC_NT) {
    const DevContact& dc = pg->contacts[c];
    const Xf H = body_to_root(s, dc.body);
    const V3 p = rot(H.R, ld3(dc.pos)) + H.p;
    const S6 T = body_twist(s, dc.body);
    const V3 vp = T.v + cross(T.w, p);
    const double phi = p.z - cm.ground_z;
    V3 f = mk3(0, 0, 0);
    if (phi < 0.0) {
      const double fn = fmax(0.0, -cm.k * phi - cm.d * vp.z);
      if (anchor) {
        double* a = anchor + 3 * c;
        if (a[2] == 0.0) {  // touch-down: the tangential spring starts unloaded
          a[0] = p.x;
          a[1] = p.y;
          a[2] = 1.0;
        }
        double fx = -cm.k * (p.x - a[0]) - cm.d * vp.x, fy = -cm.k * (p.y - a[1]) - cm.d * vp.y;
        const double ft = sqrt(fx * fx + fy * fy), lim = cm.mu * fn;
        if (ft > lim) {  // sliding: force on the cone, anchor dragged so that the spring carries exactly that force
          const double sc = ft > 0.0 ? lim / ft : 0.0;
          fx *= sc;
          fy *= sc;
          if (cm.k > 0.0) {
            a[0] = p.x + (fx + cm.d * vp.x) / cm.k;
            a[1] = p.y + (fy + cm.d * vp.y) / cm.k;
          }
        }
        f = mk3(fx, fy, fn);
      } else {
        const V3 vt = mk3(vp.x, vp.y, 0.0);
        const double nvt = sqrt(dot(vt, vt));
        const double sc = -cm.mu * fn / fmax(nvt, cm.v_eps);
        f = mk3(sc * vt.x, sc * vt.y, fn);
      }
    } else if (anchor) {
      anchor[3 * c + 2] = 0.0;  // lift-off
    }
    st6(s.A + 6 * c, mk6(cross(p, f), f));  // A (6 x max(nv, nc)) is free after the momentum matrix is no longer needed
    if (fc_out) st3(fc_out + 3 * c, f);
  }
  QPC_SYNC();
  for (int k = QPC_TID; k < nb * 6; k += QPC_NT) {
    const int b = k / 6, r = k % 6;
    double a = 0;
    for (int c = 0; c < pg->ncontacts; c++)
      if (pg->contacts[c].body == b) a += s.A[6 * c + r];
    ext[k] = a;
  }
  QPC_SYNC();
  kin_rnea(pg, s, rhs, ext, acc, tb, false);  // tb = c(q, v) - J'w: bias torques at zero acceleration
  // mass matrix, composite-rigid-body algorithm: M[i][j] = S_i . (Ic_{body(i)} S_j) when body(j) is body(i) or one of its
  // ancestors (and symmetric), 0 for velocities on different branches
  for (int k = QPC_TID; k < nv * nv; k += QPC_NT) {
    const int i = k / nv, j = k % nv;
    const int bi = pg->vbody[i], bj = pg->vbody[j];
    // deeper body = the one whose composite inertia applies; walk up from it to see whether the other is an ancestor
    int lo_ = bi, hi_ = bj, ci = i, cj = j;
    bool related = false;
    for (int pass = 0; pass < 2 && !related; pass++) {
      for (int b = lo_; b >= 0; b = pg->parent[b])
        if (b == hi_) {
          related = true;
          break;
        }
      if (!related) {
        const int t_ = lo_;
        lo_ = hi_;
        hi_ = t_;
        const int tc = ci;
        ci = cj;
        cj = tc;
      }
    }
    double m = 0.0;
    if (related) m = dot(ld6(s.SW + 6 * cj), si_mul(si_load(s.IC + 10 * lo_), ld6(s.SW + 6 * ci)));
    M[k] = m;
  }
  for (int i = QPC_TID; i < nv; i += QPC_NT) rhs[i] = tau[i] - tb[i];
  QPC_SYNC();
  // Cholesky M = L L' in place (lower triangle), column by column, then the two triangular solves
  for (int j = 0; j < nv; j++) {
    if (QPC_TID == 0) M[j * nv + j] = sqrt(M[j * nv + j]);
    QPC_SYNC();
    const double d = M[j * nv + j];
    for (int i = j + 1 + QPC_TID; i < nv; i += QPC_NT) M[i * nv + j] /= d;
    QPC_SYNC();
    for (int k = QPC_TID; k < (nv - j - 1) * (nv - j - 1); k += QPC_NT) {
      const int i = j + 1 + k / (nv - j - 1), c = j + 1 + k % (nv - j - 1);
      if (c <= i) M[i * nv + c] -= M[i * nv + j] * M[c * nv + j];
    }
    QPC_SYNC();
  }
  if (QPC_TID == 0) {  // nv <= 64: the substitutions are a serial 2 nv^2 flops
    for (int i = 0; i < nv; i++) {
      double a = rhs[i];
      for (int c = 0; c < i; c++) a -= M[i * nv + c] * rhs[c];
      rhs[i] = a / M[i * nv + i];
    }
    for (int i = nv - 1; i >= 0; i--) {
      double a = rhs[i];
      for (int c = i + 1; c < nv; c++) a -= M[c * nv + i] * rhs[c];
      rhs[i] = a / M[i * nv + i];
    }
  }
  QPC_SYNC();
  for (int i = QPC_TID; i < nv; i += QPC_NT) vd_out[i] = rhs[i];
  QPC_SYNC();
}

}  // namespace qpc
