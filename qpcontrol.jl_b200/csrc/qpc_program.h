// qpc_program.h -- the static "controller program": flat device tables compiled once from the setup-time API calls
// (MomentumBasedController / addtask! / addcontact! / regularize! / StandingController, reference
// src/lowlevel/momentum.jl:15-33,99-148 and src/highlevel/standing.jl:18-56).  One copy lives in global memory per
// device and is read by every CTA.
#pragma once
#include <stdint.h>

#define QPC_MAXB 48     // bodies
#define QPC_MAXV 64     // velocity dimension
#define QPC_MAXQ 72     // configuration dimension
#define QPC_MAXT 64     // tasks
#define QPC_MAXC 16     // contact points
#define QPC_MAXN 8      // friction-cone generators per contact
#define QPC_MAXPATH 512 // total path entries over all tasks
#define QPC_MAXDES 160  // total desired dimension
#define QPC_MAXW 256    // matrix-weight storage (doubles)
#define QPC_MAXANC 1200 // total entries of the per-body ancestor / descendant lists (sum of depths)
#define QPC_MAXSE3 4    // SE3PDControllers evaluated on the device
#define QPC_MAXSEG 6    // Interpolated pieces of one (Piecewise) trajectory

namespace qpc {

struct Settings {
  double rho, sigma, alpha, eps_abs, eps_rel, eps_prim_inf, eps_dual_inf, adaptive_rho_tolerance;
  int max_iter, scaling, adaptive_rho, adaptive_rho_interval, check_termination;
};

struct DevTask {
  int kind, mode, dim, source, target, frame, joint, des_off;
  int row0;                // first row in G (hard: task_error == 0; weighted: e == task_error) or -1 when eliminated
  int scol0;               // first slack column in x (weighted tasks, momentum.jl:119-126) or -1
  int path_ptr, path_len;  // entries of path_body / path_sign
  int w_off;               // matrix weight offset in Wbuf
  int eliminated;          // hard joint task: its velocities are substituted, no rows
  double weight;
  double point[3];
};

struct DevContact {
  int body, col0;       // first rho column in x
  double Rz[9], pos[3]; // z_up_transform(position, normal)  (contacts.jl:8-14)
  double B[3 * QPC_MAXN];          // forcebasis(mu, N), 3 x N row-major (contacts.jl:16-23)
  double BtB[QPC_MAXN * QPC_MAXN]; // B'B
  double maxrho_factor;            // 1 / (N sqrt(mu^2 + 1))  (contacts.jl:57)
};

// One `Interpolated` piece (reference src/trajectories/interpolated.jl:1-60) of a `Piecewise` trajectory
// (piecewise.jl:1-40): active from break `brk`, evaluated at x - brk.  Rotation: y0 = quaternion (w, x, y, z), dy = unit
// axis of y0 \ yf, angle its rotation angle (interpolated.jl:75-82); vector: y0[0..2], dy = yf - y0.  c = ascending
// coefficients of the polynomial interpolator alpha(theta), nc = 0 for the identity.
struct DevInterp {
  double brk, x0, xf, y0[4], dy[3], angle, c[6];
  int nc, pad;
};
struct DevTraj {
  int nseg, piecewise;  // not piecewise: seg[0] evaluated at x itself
  double brk_end;       // last break of the Piecewise (clamp range [seg[0].brk, brk_end])
  DevInterp seg[QPC_MAXSEG];
};
// SE3PDController (reference src/lowlevel/se3pdcontroller.jl:1-18) driving the desired of one SpatialAccelerationTask:
// SE3Trajectory (src/trajectories/se3.jl:1-27) of `body` in `base` + SE3PDGains (3x3 row-major: K_ang, D_ang, K_lin, D_lin)
struct DevSE3PD {
  int body, base, des_off, task;
  double K[36];
  DevTraj ang, lin;
};

struct DevProgram {
  // ---- mechanism -------------------------------------------------------------------------------------------
  int nb, nq, nv;
  int parent[QPC_MAXB], jtype[QPC_MAXB], qoff[QPC_MAXB], voff[QPC_MAXB], nvj[QPC_MAXB];
  int vbody[QPC_MAXV];
  int nlevels, level_ptr[QPC_MAXB + 1], level_body[QPC_MAXB];  // bodies grouped by depth
  int child_ptr[QPC_MAXB + 1], child_idx[QPC_MAXB];
  // chain-walk sweeps (kin_forward / kin_composite): ancestors of body b including b itself, root first; and its proper
  // descendants, deepest level first -- each body folds its own chain, one barrier per sweep instead of one per tree level
  int anc_ptr[QPC_MAXB + 1], anc_idx[QPC_MAXANC];
  int desc_ptr[QPC_MAXB + 1], desc_idx[QPC_MAXANC];
  double axis[QPC_MAXB * 3], XR[QPC_MAXB * 9], Xp[QPC_MAXB * 3];
  double inertia[QPC_MAXB * 10];  // body-frame SI (see qpc_common.h)
  double gravity[3], total_mass;
  // ---- program ---------------------------------------------------------------------------------------------
  int N, floating;  // floating = successor body of the floating joint or -1
  int ntasks, ncontacts, ndes;
  int n, nvf, ne, mg, nbx;  // QP the device solves: x = (free vd [nvf], task-error slacks e [ne], rho [nbx]);
                            // mg general (equality) rows, nbx box rows 0 <= rho <= maxrho
  int balance_row0;
  int nel;  // leading variables of x with a diagonal, strictly positive cost block (regularised free vd, scalar-weight
            // slacks) that the ADMM fast path may eliminate; 0 when the program does not qualify
  int vcol[QPC_MAXV];      // column of velocity i in x, or -1 when fixed by a hard JointAccelerationTask
  int vfix_des[QPC_MAXV];  // desired offset providing the value of a fixed velocity
  double reg[QPC_MAXV];
  DevTask tasks[QPC_MAXT];
  int path_body[QPC_MAXPATH], path_sign[QPC_MAXPATH];
  // column-parallel task rows (kin_task_rows): sign of velocity column c in path task t's Jacobian (0: its body is not on
  // the path), and the velocities fixed by hard JointAccelerationTasks, ascending
  signed char tsign[QPC_MAXT][QPC_MAXV];
  int nfixv, fixv[QPC_MAXV];
  int nactive, active[QPC_MAXT];  // tasks that own rows (not eliminated), addtask! order
  double Wbuf[QPC_MAXW];
  int nwmat;  // doubles of Wbuf in use (sum of dim^2 over the matrix-weighted tasks)
  DevContact contacts[QPC_MAXC];
  double def_desired[QPC_MAXDES], def_cweight[QPC_MAXC], def_cmaxnf[QPC_MAXC];
  // ---- StandingController constants (standing.jl:1-16) ---------------------------------------------------------
  int standing, st_linmom_des, st_pelvis_des, st_pelvis_body, st_nj;
  int st_jq[QPC_MAXV], st_jv[QPC_MAXV], st_jdes[QPC_MAXV];
  double st_kp[QPC_MAXV], st_kd[QPC_MAXV], st_ref[QPC_MAXV];
  double st_com_kp, st_com_kd, st_pelvis_kp, st_pelvis_kd, st_comref[3];
  // ---- SE3PDControllers evaluated in the assembly prologue (se3pdcontroller.jl:13-18) ------------------------------
  int nse3;
  DevSE3PD se3[QPC_MAXSE3];
  Settings settings;
};

// per-batch I/O of the tick (device pointers)
struct BatchIO {
  const double *q, *v, *desired, *cweight, *cmaxnf;
  long long desired_stride, contact_stride;
  // per-tick Parameters of the reference beyond the above, both optional:
  const double* tweight = nullptr;  // [B][tweight_stride] scalar weight of every task, addtask! order (momentum.jl:107-110)
  const double* cgeom = nullptr;    // [B][cgeom_stride] per contact position[3], normal[3], mu (contacts.jl:39,53-61)
  long long tweight_stride = 0, cgeom_stride = 0;
  const double* twmat = nullptr;    // [B][twmat_stride] matrix weights of the matrix-weighted tasks, Wbuf layout (momentum.jl:113-117)
  long long twmat_stride = 0;
  const double* time = nullptr;     // controller time t of the tick: [B] (time_stride 1) or one value (time_stride 0)
  long long time_stride = 0;
  double time_offset = 0.0;         // added to the time: k * dt at tick k of the closed loops (qpc_step_batch / qpc_simulate_batch)
};

// the condensed QP of every instance, as written by the assembly kernel and consumed by the ADMM kernel
struct QpBuffers {
  double *P, *qv, *G, *lg, *ug, *lb, *ub;  // [B][n*n], [B][n], [B][mg*n], [B][mg], [B][mg], [B][nbx], [B][nbx]
  double* des;                              // [B][ndes] desireds actually used (standing laws applied)
  double *x, *y;                            // [B][n], [B][mg+nbx]
  int *status, *iters;
  double* res;  // [B][2]
  int* nfac;    // [B] number of factorisations (1 + rho updates), optional
  double* rho;  // [B] final rho of the slot's last accepted solve (<= 0: none); read when `warm`
  double* ksave = nullptr;  // [B][kin_save_doubles] kinematic state handed from the assembly to the inverse-dynamics kernel
  int prezeroed = 0;  // P and G were zeroed when the workspace was allocated and only the assembly kernel writes them: the
                      // structurally zero entries (a static pattern of the program) need no per-tick clearing
  int warm;     // OSQP's implicit warm start: start from x, y, rho of the previous tick of the same slot
  // list mode (the instances the one-warp kernel handed back): solve instances list[0 .. *list_count) only, on a
  // bounded persistent grid of list_grid CTAs
  const int* list = nullptr;
  const int* list_count = nullptr;
  int list_grid = 0;
};

}  // namespace qpc
