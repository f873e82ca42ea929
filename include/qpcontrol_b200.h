/* qpcontrol_b200.h -- C ABI of libqpcontrol_b200.so
 *
 * Drop-in boundary for QPControl.jl's per-timestep control loop, batched over independent robot instances.
 * The reference (tkoolen/QPControl.jl) is pure Julia and has no FFI of its own; the seam this library replaces is the
 * controller functor `(controller)(tau, t, x)` (reference src/lowlevel/momentum.jl:41-81, src/highlevel/standing.jl:58-89)
 * and everything it reaches: Parametron `solve!` (momentum.jl:58), the OSQP optimizer plugged in as type parameter `O`
 * (momentum.jl:1,15-16,27) and the RigidBodyDynamics state queries of SURVEY.md 8(a) row a15.  The setup-time calls
 * mirror the reference constructors one to one so a Julia shim (qpcontrol.jl_b200/julia/QPControlB200.jl) can `ccall`
 * them from the same user-facing API.  Plain pointers and sizes only.
 *
 * Conventions: all reals are IEEE fp64; matrices are row-major; spatial vectors are (angular; linear)
 * (reference src/tasks.jl:41-43, src/contacts.jl:84-88); bodies are indexed 0..nb-1 in topological order with -1 = world;
 * a joint is identified with its successor body; floating joint q = (w,x,y,z,px,py,pz), v = (omega; v) in body frame.
 * Every function returning int returns 0 on success and a negative error code otherwise; qpc_last_error() describes
 * the last failure of the calling thread.  CUDA errors never cross the ABI as exceptions.
 */
#ifndef QPCONTROL_B200_H
#define QPCONTROL_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct qpc_mechanism qpc_mechanism;   /* replaces RigidBodyDynamics.Mechanism (momentum.jl:15-17) */
typedef struct qpc_controller qpc_controller; /* replaces MomentumBasedController{N,O,S} (momentum.jl:1-34) */

/* joint types (RigidBodyDynamics: Revolute, Prismatic, QuaternionFloating, Fixed) */
enum { QPC_REVOLUTE = 0, QPC_PRISMATIC = 1, QPC_QUAT_FLOATING = 2, QPC_FIXED = 3 };

/* task kinds: the seven types of reference src/tasks.jl */
enum {
  QPC_TASK_SPATIAL = 0,        /* SpatialAccelerationTask   tasks.jl:3-44    */
  QPC_TASK_ANGULAR = 1,        /* AngularAccelerationTask   tasks.jl:47-84   */
  QPC_TASK_LINEAR = 2,         /* LinearAccelerationTask    tasks.jl:86-123  */
  QPC_TASK_POINT = 3,          /* PointAccelerationTask     tasks.jl:125-171 */
  QPC_TASK_JOINT = 4,          /* JointAccelerationTask     tasks.jl:173-189 */
  QPC_TASK_MOMENTUM_RATE = 5,  /* MomentumRateTask          tasks.jl:192-236 */
  QPC_TASK_LINEAR_MOMENTUM_RATE = 6 /* LinearMomentumRateTask tasks.jl:239-262 */
};
/* addtask! flavours (momentum.jl:99-117) */
enum { QPC_MODE_HARD = 0, QPC_MODE_SCALAR_WEIGHT = 1, QPC_MODE_MATRIX_WEIGHT = 2 };

/* per-instance solver status: OSQP's codes; checkstatus (momentum.jl:83-91) accepts SOLVED and SOLVED_INACCURATE */
enum {
  QPC_SOLVED = 1,
  QPC_SOLVED_INACCURATE = 2,
  QPC_PRIMAL_INFEASIBLE_INACCURATE = 3,
  QPC_DUAL_INFEASIBLE_INACCURATE = 4,
  QPC_MAX_ITER_REACHED = -2,
  QPC_PRIMAL_INFEASIBLE = -3,
  QPC_DUAL_INFEASIBLE = -4,
  QPC_NON_FINITE = -8,
  QPC_UNSOLVED = -10
};

/* error codes */
enum { QPC_OK = 0, QPC_ERR_ARG = -1, QPC_ERR_CUDA = -2, QPC_ERR_STATE = -3, QPC_ERR_LIMIT = -4 };

/* OSQPSettings.* attributes the reference sets (test/runtests.jl:35-43, notebooks/Standing controller.ipynb:66-71);
 * defaults are OSQP 0.5.x's (qpc_default_settings). */
typedef struct {
  double rho, sigma, alpha;
  double eps_abs, eps_rel, eps_prim_inf, eps_dual_inf;
  double adaptive_rho_tolerance;
  int32_t max_iter, scaling, adaptive_rho, adaptive_rho_interval, check_termination;
  int32_t reserved[3];
} qpc_settings;

/* solve_batch flags */
enum {
  QPC_HOST_PTRS = 0,   /* all batch pointers are host memory: the library copies H2D into its device staging buffers
                          (chunk by chunk, each chunk's copies on that chunk's stream), runs, copies D2H and synchronises
                          before returning.  Page-locked caller buffers (qpc_pin_host_buffer, or memory that already is
                          pinned) make those copies asynchronous DMA that overlaps the other chunks' kernels; PAGEABLE
                          buffers (a plain Julia Matrix{Float64}) are accepted -- the CUDA driver then stages them and
                          each copy blocks the calling thread, so the PCIe time is no longer hidden */
  QPC_DEVICE_PTRS = 1  /* all batch pointers are device memory on the controller's device: fully asynchronous on
                          `stream` */
};

/* Inputs of one batched tick.  Replaces `x` of the functor (momentum.jl:41,57: x = [q; v]) plus the mutable per-tick
 * fields the reference reads through Parameters: task `desired` (tasks.jl:40,82,121,165,186,232,260) and
 * ContactPoint.weight / .maxnormalforce (contacts.jl:35-36,54-57,76). */
typedef struct {
  const double* q;          /* [B][nq] */
  const double* v;          /* [B][nv] */
  const double* desired;    /* [B][desired_stride] task desireds concatenated in addtask! order, or NULL = the
                               values of qpc_set_task_desired; entries owned by the standing controller are ignored */
  int64_t desired_stride;   /* >= ndes, or 0 to broadcast one row */
  const double* contact_weight;         /* [B][contact_stride] or NULL = qpc_set_contact_params values */
  const double* contact_maxnormalforce; /* [B][contact_stride] or NULL */
  int64_t contact_stride;   /* >= ncontacts, or 0 to broadcast one row */
  /* Remaining per-tick Parameters of the reference; both may be NULL (= setup-time values): */
  const double* task_weight;      /* [B][task_weight_stride] scalar weight of every task in addtask! order -- a
                                     Parameter-valued `weight` of addtask!(controller, task, weight), momentum.jl:107-110;
                                     entries of hard and matrix-weighted tasks are ignored (matrix weights: below) */
  int64_t task_weight_stride;     /* >= ntasks, or 0 to broadcast one row */
  const double* contact_geometry; /* [B][contact_geometry_stride]: per contact position[3], normal[3] (body frame) and
                                     mu -- ContactPoint.position / .normal / .mu as Parameters, contacts.jl:39,53-61;
                                     exercised by test/controller.jl:42-47,110-118 */
  int64_t contact_geometry_stride; /* >= 7 * ncontacts, or 0 to broadcast one row */
  const double* task_weight_matrix; /* [B][task_weight_matrix_stride]: the dim x dim weights (row-major) of the
                                     matrix-weighted tasks, concatenated in addtask! order -- a Parameter-valued matrix
                                     `weight` of addtask!(controller, task, weight), momentum.jl:113-117, exercised by
                                     test/controller.jl:232-285; NULL = the matrices given at setup.  The Hessian block of
                                     the task's slack is W + W' (an unsymmetric W is allowed, as in the reference) */
  int64_t task_weight_matrix_stride; /* >= qpc_controller_weight_matrix_doubles(), or 0 to broadcast one row */
  const double* time;       /* controller time t of the tick -- the first argument of the functors,
                               (controller::SE3PDController)(t, state), se3pdcontroller.jl:13: [B] (time_stride 1) or
                               one value for the whole batch (time_stride 0); NULL = 0.  Read only by controllers
                               with device-side SE3PDControllers (qpc_add_se3pd) */
  int64_t time_stride;
} qpc_batch_in;

/* Outputs of one batched tick: what the reference leaves in tau (momentum.jl:75-80), controller.result.vd (:62-64)
 * and the per-point world-frame wrenches it sums into controller.contactwrenches (:65-72). Any pointer may be NULL. */
typedef struct {
  double* tau;       /* [B][nv]  floating-joint entries exactly 0 (momentum.jl:93-97) */
  double* vdot;      /* [B][nv] */
  double* wrench;    /* [B][ncontacts][6] world frame (angular; linear) */
  int32_t* status;   /* [B] */
  int32_t* iters;    /* [B] ADMM iterations */
  double* residuals; /* [B][2] unscaled primal / dual residual (OSQP definition) of the QP the device solved */
  int32_t* factorizations; /* [B] KKT factorisations = 1 + rho updates (OSQP info.rho_updates + 1) */
} qpc_batch_out;

int qpc_version(void);
const char* qpc_last_error(void);
int qpc_device_count(void);
void qpc_default_settings(qpc_settings* s);

/* ---- mechanism: RigidBodyDynamics.Mechanism as a flat tree --------------------------------------------------- */
qpc_mechanism* qpc_mechanism_create(int32_t nb, const int32_t* parent, const int32_t* jtype,
                                    const double* axis /*[nb][3]*/, const double* X_R /*[nb][9]*/,
                                    const double* X_p /*[nb][3]*/, const double* mass /*[nb]*/,
                                    const double* com /*[nb][3]*/, const double* inertia_origin /*[nb][9]*/,
                                    const double gravity[3]);
void qpc_mechanism_destroy(qpc_mechanism*);
int qpc_mechanism_dims(const qpc_mechanism*, int32_t* nb, int32_t* nq, int32_t* nv);

/* ---- controller setup: MomentumBasedController{N}(mechanism, optimizer; floatingjoint) (momentum.jl:15-33) ------- */
qpc_controller* qpc_controller_create(qpc_mechanism*, int32_t N, int32_t floating_body /* -1 = fixed base */,
                                      const qpc_settings*);
void qpc_controller_destroy(qpc_controller*);
/* addcontact!(controller, body, position, normal, mu) (momentum.jl:142-148, contacts.jl:38-69); returns index >= 0 */
int qpc_add_contact(qpc_controller*, int32_t body, const double position[3], const double normal[3], double mu);
/* contact.weight[] / contact.maxnormalforce[] defaults (contacts.jl:35-36; both start at 0 = disabled, :50) */
int qpc_set_contact_params(qpc_controller*, int32_t contact, double weight, double maxnormalforce);
/* addtask!(controller, task[, weight]) (momentum.jl:99-117); W is dim x dim for QPC_MODE_MATRIX_WEIGHT; index >= 0 */
int qpc_add_task(qpc_controller*, int32_t kind, int32_t source_body, int32_t target_body, int32_t frame_body,
                 const double point[3], int32_t joint, int32_t mode, double weight, const double* W);
/* setdesired!(task, desired): default used when qpc_batch_in.desired is NULL */
int qpc_set_task_desired(qpc_controller*, int32_t task, const double* desired);
/* regularize!(controller, joint, weight) (momentum.jl:128-131) */
int qpc_regularize(qpc_controller*, int32_t joint, double weight);
/* StandingController's per-tick PD laws (standing.jl:58-85), evaluated on the device before the low-level tick */
int qpc_standing_setup(qpc_controller*, int32_t linmom_task, int32_t pelvis_task, int32_t pelvis_body, int32_t njoints,
                       const int32_t* joint_tasks, const int32_t* joints, const double* kp, const double* kd,
                       const double* qref, double com_kp, double com_kd, double pelvis_kp, double pelvis_kd,
                       const double comref[3]);
/* ---- SE3PDController + SE3Trajectory on the device ----------------------------------------
 * One `Interpolated` piece (src/trajectories/interpolated.jl:1-60) of a trajectory; a `Piecewise` (piecewise.jl:1-40)
 * is a list of pieces, piece i active from break_start and evaluated at x - break_start.  Rotations: y0 = unit
 * quaternion (w, x, y, z) of the start, dy = unit axis and angle = rotation angle of y0 \ yf (interpolated.jl:75-82);
 * vectors: y0[0..2], dy = yf - y0.  coeffs = ascending coefficients of the polynomial interpolator alpha(theta)
 * (fit_polynomial.jl), ncoeffs = 0 for the identity.  A `Constant` (constant.jl) is a piece with dy = 0. */
typedef struct {
  double break_start, x0, xf;
  double y0[4], dy[3], angle;
  double coeffs[6];
  int32_t ncoeffs, reserved;
} qpc_interp_piece;
#define QPC_MAX_SE3PD 4
#define QPC_MAX_PIECES 6
/* SE3PDController(base, body, trajectory, weight, gains) (se3pdcontroller.jl:1-11) whose output
 * Tdref + pd(gains, H, Href, T, Tref) (:13-18) becomes the desired of SpatialAccelerationTask `task` every tick, evaluated
 * in the assembly kernel at qpc_batch_in.time (the entries of qpc_batch_in.desired for that task are then ignored).
 * gains: four 3 x 3 row-major matrices K_angular, D_angular, K_linear, D_linear (SE3PDGains in the body frame).
 * angular / linear: the two components of the SE3Trajectory (src/trajectories/se3.jl:1-27); *_piecewise = 0 evaluates
 * piece 0 at t itself, 1 clamps t to [pieces[0].break_start, *_break_end] first.  Trajectories are clamped to their
 * range (Interpolated's clamp = true); returns the controller's index >= 0.  Before qpc_finalize. */
int qpc_add_se3pd(qpc_controller*, int32_t task, int32_t base_body, int32_t body, const double gains[36],
                  int32_t n_angular, const qpc_interp_piece* angular, int32_t angular_piecewise, double angular_break_end,
                  int32_t n_linear, const qpc_interp_piece* linear, int32_t linear_piecewise, double linear_break_end);
/* controller.trajectory[] = ... / controller.gains[] = ... (both are Refs in the reference, se3pdcontroller.jl:4-6):
 * replace them between ticks; NULL leaves that part unchanged */
int qpc_se3pd_update(qpc_controller*, int32_t se3pd, const double* gains,
                     int32_t n_angular, const qpc_interp_piece* angular, int32_t angular_piecewise, double angular_break_end,
                     int32_t n_linear, const qpc_interp_piece* linear, int32_t linear_piecewise, double linear_break_end);
int qpc_set_settings(qpc_controller*, const qpc_settings*);
/* initialize! (momentum.jl:150-156): freezes the program, builds the device tables on `device` */
int qpc_finalize(qpc_controller*, int32_t device);
/* sizes: nq, nv, ndes, ncontacts, and the dims (n, m_general, n_box) of the condensed QP the device solves */
int qpc_controller_dims(const qpc_controller*, int32_t* nq, int32_t* nv, int32_t* ndes, int32_t* ncontacts,
                        int32_t* n, int32_t* mg, int32_t* nbox);
/* nwmat: doubles of one row of qpc_batch_in.task_weight_matrix (sum of dim^2 over the matrix-weighted tasks) */
int qpc_controller_weight_matrix_doubles(const qpc_controller*);

/* ---- the control tick for B instances: (controller)(tau, t, x) (momentum.jl:41-81 / standing.jl:58-89) ------------ */
int qpc_solve_batch(qpc_controller*, int64_t B, const qpc_batch_in*, const qpc_batch_out*, int32_t flags,
                    void* stream /* cudaStream_t, QPC_DEVICE_PTRS only; NULL = default stream */);
/* The same tick for one batch spread over several devices from ONE process (SURVEY.md 8(b) threading row, 8(e)): `ctrls`
 * are replicas of one program finalized on different devices (qpc_finalize(ctrl_k, device_k)); the batch is cut into
 * nctrl contiguous shards (the first B % nctrl shards one instance longer), one host thread per controller runs
 * qpc_solve_batch(QPC_HOST_PTRS) on its shard, nothing is exchanged between devices.  Results do not depend on nctrl. */
int qpc_solve_batch_multi(qpc_controller* const* ctrls, int32_t nctrl, int64_t B, const qpc_batch_in*,
                          const qpc_batch_out*);
/* page-lock / release a caller-owned host buffer used with QPC_HOST_PTRS (see the flag's description) */
int qpc_pin_host_buffer(void* ptr, int64_t bytes);
int qpc_unpin_host_buffer(void* ptr);
/* ensure workspaces for batches up to B exist (no allocation happens inside qpc_solve_batch afterwards) */
int qpc_reserve(qpc_controller*, int64_t B);
/* number of kernels launched by this controller so far */
int64_t qpc_launch_count(const qpc_controller*);
/* measurement hooks: when profiling is on, every tick records CUDA events around its three kernels on the launching
 * stream; qpc_stage_times waits for the last tick and returns {assembly, ADMM, inverse dynamics} milliseconds */
int qpc_set_profiling(qpc_controller*, int32_t on);
int qpc_stage_times(qpc_controller*, double ms[3]);
/* ADMM fast path: when every free acceleration is regularised and every weighted task has a positive scalar weight
 * (the StandingController's program, standing.jl:35-49), those variables have a diagonal cost block and every general
 * row is an equality; the solver then eliminates them from the KKT system (same OSQP iterates, half the flops) as long
 * as eps_abs >= 1e-6 and no per-tick task weights are passed.  qpc_set_admm_elimination(ctrl, 0) forces the full
 * system; qpc_admm_eliminated returns how many variables the next tick eliminates (0 = full system). */
int qpc_set_admm_elimination(qpc_controller*, int32_t on);
int qpc_admm_eliminated(const qpc_controller*);
/* One-warp-per-QP ADMM (csrc/admm_warp.cuh): programs whose general rows are all equalities and determine the unboxed
 * variables (the StandingController's program: 21 of them, 24 rows) are first reduced to a QP in the <= 32 friction-cone
 * multipliers alone (Householder QR of the equality block, no KKT system), then iterated with a 32 x 32 operator held in
 * one warp's registers; OSQP's termination test is evaluated on the full problem.  Instances the reduction cannot
 * handle are solved by the register-tile kernel in the same tick.  On by default where the program qualifies;
 * qpc_set_admm_warp(ctrl, 0) forces the KKT-system kernels, qpc_admm_warp reports whether the next tick uses it. */
int qpc_set_admm_warp(qpc_controller*, int32_t on);
int qpc_admm_warp(const qpc_controller*);
/* fp64 FMA throughput of the device in TFLOP/s (dependent-chain-free DFMA loop), the roofline denominator */
int qpc_measure_fp64_peak(int32_t device, double* tflops);

/* ---- sequential ticks: warm start and the closed loop ------------------------------------------
 * The reference solves every tick in the same OSQP workspace, so each solve starts from the previous tick's primal /
 * dual iterates and the rho it had adapted to (OSQP's implicit warm start behind `solve!`, momentum.jl:58; SURVEY.md
 * 8(a) a12).  qpc_set_warm_start(ctrl, 1) gives every batch slot that behaviour: the ADMM solve of slot i starts from
 * the (x, y, rho) slot i ended its previous accepted solve with.  Off by default: every tick is a cold start. */
int qpc_set_warm_start(qpc_controller*, int32_t on);
int qpc_reset_warm_start(qpc_controller*); /* forget the stored iterates: the next tick starts cold */
/* `nsteps` control ticks in closed loop without leaving the device: after every tick the states advance in place by
 * dt with the commanded accelerations (semi-implicit Euler; the quaternion of a floating joint through the exponential
 * map) -- the batched counterpart of `simulate(state, T, PeriodicController(tau, dt, controller))` in
 * notebooks/Standing controller.ipynb:202-214 under the model the controller itself assumes (its contact wrenches are
 * the ones applied, so forward dynamics returns the commanded vd: test/controller.jl:92-96).  q [B][nq] and v [B][nv]
 * are read and overwritten (host or device pointers per flags); `in` supplies desired / contact arrays as in
 * qpc_solve_batch (its q and v are ignored; may be NULL); `out` receives the last tick's outputs (may be NULL). */
int qpc_step_batch(qpc_controller*, int64_t B, double* q, double* v, const qpc_batch_in* in, const qpc_batch_out* out,
                   double dt, int32_t nsteps, int32_t flags, void* stream);

/* The same loop with a PLANT (notebooks/Standing controller.ipynb:202-214: simulate(state, T, PeriodicController(tau, dt,
 * controller)) -- the simulator applies the commanded torques, the environment answers with contact forces): after every
 * control tick of period dt the state is advanced by `substeps` steps of dt / substeps of the forward dynamics
 * vd = M(q)^-1 (tau - c(q, v) + sum J'f) (composite-rigid-body mass matrix, Cholesky, RNEA bias) with tau held
 * (PeriodicController's zero-order hold) and every ContactPoint of the controller pressed against the half-space
 * z >= ground_z by a spring-damper normal force max(0, -k phi - d phidot) with regularised Coulomb friction
 * -mu f_n v_t / max(|v_t|, v_eps).  Semi-implicit Euler on the configuration manifold, as qpc_step_batch. */
typedef struct {
  double stiffness; /* k [N/m] per contact point */
  double damping;   /* d [N s/m] per contact point */
  double mu;        /* friction coefficient of the ground */
  double v_eps;     /* [m/s] sliding speed below which friction is proportional to it */
  double ground_z;  /* height of the ground plane in the world frame */
} qpc_contact_model;
int qpc_simulate_batch(qpc_controller*, int64_t B, double* q, double* v, const qpc_batch_in* in, const qpc_batch_out* out,
                       const qpc_contact_model* plant, double dt, int32_t substeps, int32_t nticks, int32_t flags,
                       void* stream);

/* ---- stage-level entry points (parity tests; the tick above is their composition) -------------------------------- */
/* kinematics + QP assembly only: writes the condensed QP of every instance (device or host pointers per flags):
 * P [B][n*n], qv [B][n], G [B][mg*n], lg/ug [B][mg], lb/ub [B][nbox], desired_out [B][ndes] */
int qpc_assemble_batch(qpc_controller*, int64_t B, const qpc_batch_in*, double* P, double* qv, double* G, double* lg,
                       double* ug, double* lb, double* ub, double* desired_out, int32_t flags, void* stream);

/* ---- raw batched dense QPs: min 1/2 x'Px + q'x  s.t. lg <= G x <= ug, lb <= x_tail <= ub
 * where the box applies to the last nbox variables.  Pointers per flags. */
int qpc_solve_qp_batch(int32_t device, int64_t B, int32_t n, int32_t mg, int32_t nbox, const double* P,
                       const double* qv, const double* G, const double* lg, const double* ug, const double* lb,
                       const double* ub, const qpc_settings*, double* x, double* y /*[B][mg+nbox]*/, int32_t* status,
                       int32_t* iters, double* residuals, int32_t flags, void* stream);

#ifdef __cplusplus
}
#endif
#endif
