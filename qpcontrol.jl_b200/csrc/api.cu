// api.cu -- CUDA backend of the C ABI (include/qpcontrol_b200.h): kernels, workspaces, launches.
// Built for sm_100a only:  nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo ...   (see Makefile)
#include <cuda_runtime.h>

#include <array>
#include <atomic>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <mutex>
#include <string>
#include <thread>
#include <vector>

struct DeviceBuffers {
  // QP workspace (both pointer modes)
  double *P = nullptr, *qv = nullptr, *G = nullptr, *lg = nullptr, *ug = nullptr, *lb = nullptr, *ub = nullptr;
  double *des = nullptr, *x = nullptr, *y = nullptr;
  double* rho = nullptr;  // [capB] rho every slot ended its last accepted solve with (<= 0: none) -- warm start
  double* ksave = nullptr;  // [capB][kin_save_doubles] kinematic state handed from assembly to inverse dynamics
  double* anchor = nullptr; // [capB][ncontacts][3] tangential-spring anchors of the plant's contact model (qpc_simulate_batch)
  // staging for QPC_HOST_PTRS
  double *q = nullptr, *v = nullptr, *desired = nullptr, *cw = nullptr, *cm = nullptr, *tw = nullptr, *cg = nullptr;
  double *tau = nullptr, *vdot = nullptr, *wrench = nullptr, *res = nullptr;
  int *status = nullptr, *iters = nullptr;
  double* twm = nullptr;  // staging of per-tick matrix weights
  double* time = nullptr; // staging of the controller times (SE3PDControllers)
  long long capB = 0, cap_desired = 0, cap_contact = 0, cap_tw = 0, cap_cg = 0, cap_twm = 0, cap_time = 0;
};
struct Backend {
  bool dirty = false;
  int device = 0;
  void* d_prog = nullptr;
  cudaStream_t stream = nullptr;
  DeviceBuffers buf;
  std::atomic<long long> launches{0};
  std::mutex mu;
  bool profiling = false;
  bool warm_start = false;  // qpc_set_warm_start
  bool elimination = true;  // qpc_set_admm_elimination: allow the fast path with eliminated free variables
  cudaEvent_t ev[4] = {nullptr, nullptr, nullptr, nullptr};
  int* d_nfac = nullptr;  // [capB] factorisations per instance
  int* d_fb_list = nullptr;   // [capB] instances the one-warp ADMM kernel handed back (per chunk: its own range of the list)
  int* d_fb_count = nullptr;  // [NSIDE + 1] their number, per chunk
  bool warp_admm = true;      // qpc_set_admm_warp: allow the one-warp-per-QP kernel (admm_warp.cuh)
  // chunked tick: sub-batches go to side streams so that one chunk's assembly / inverse dynamics and the tail of its
  // ADMM kernel overlap the next chunk's ADMM kernel
#ifndef QPC_NSIDE
#define QPC_NSIDE 4
#endif
  static constexpr int NSIDE = QPC_NSIDE;
  cudaStream_t side[NSIDE] = {};
  cudaEvent_t fork = nullptr, join[NSIDE] = {};
};

#include "admm.cuh"
#include "admm_reg.cuh"
#include "admm_warp.cuh"
#include "kin.cuh"
#include "kin_warp.h"
#include "tiny_thread.h"
#include "setup_api.h"

using namespace qpc;

static thread_local std::string g_launch_note;  // resource figures of the last failed kernel launch
#define CUDA_TRY(expr)                                                                                  \
  do {                                                                                                  \
    cudaError_t e__ = (expr);                                                                           \
    if (e__ != cudaSuccess)                                                                             \
      return qpc_fail(QPC_ERR_CUDA, std::string(#expr) + ": " + cudaGetErrorString(e__) + g_launch_note); \
  } while (0)

// ---- kernels: one robot instance / one QP per CTA ----------------------------------------------------------------------
#ifndef QPC_KIN_THREADS
#define QPC_KIN_THREADS 64
#endif
constexpr int ASM_THREADS = QPC_KIN_THREADS;
constexpr int ADMM_THREADS = 128;
constexpr int ADMM_BIG_THREADS = 512;
constexpr int ADMM_BIG_SMEM = 100 * 1024;  // above this at most two CTAs fit an SM: run them with ADMM_BIG_THREADS
constexpr int ID_THREADS = QPC_KIN_THREADS;

// SE3: programs with device-side SE3PDControllers (kin_se3pd); a separate instantiation so that the evaluator's registers
// stay out of the kernel every other program runs (one shared kernel cost the Atlas standing tick 0.18 ms in spills)
#ifndef QPC_ASM_MIN_CTAS
#define QPC_ASM_MIN_CTAS 10  // register budget of the plain instantiation: 96 registers (shared memory admits nine CTAs per SM)
#endif
template <bool SE3>
__global__ void __launch_bounds__(ASM_THREADS, SE3 ? 1 : QPC_ASM_MIN_CTAS)
qpc_assemble_kernel(const DevProgram* __restrict__ pg, BatchIO io, QpBuffers qb, long long base, long long B) {
  extern __shared__ double smem[];
  for (long long inst = base + blockIdx.x; inst < B; inst += gridDim.x) {
    KinSmem s = kin_layout(smem, pg->nb, pg->nq, pg->nv, pg->ndes, pg->ncontacts, pg->N);
    kin_load(pg, io, inst, s);
    kin_forward(pg, s);
    kin_composite(pg, s);
    kin_standing(pg, s);
    if (SE3) kin_se3pd(pg, io, inst, s);
    kin_contacts(pg, s);
    const int n = pg->n, mg = pg->mg, nbx = pg->nbx;
    kin_assemble(pg, s, qb.P + inst * n * n, qb.qv + inst * n, qb.G + inst * mg * n, qb.lg + inst * mg,
                 qb.ug + inst * mg, qb.lb + inst * nbx, qb.ub + inst * nbx, qb.prezeroed != 0);
    for (int i = threadIdx.x; i < pg->ndes; i += blockDim.x) qb.des[inst * pg->ndes + i] = s.des[i];
    if (qb.ksave) kin_save(pg, s, qb.ksave + inst * kin_save_doubles(pg->nb, pg->nv, pg->ncontacts, pg->N));
    __syncthreads();
  }
}

// NT = ADMM_THREADS for QPs of which several fit one SM's shared memory; ADMM_BIG_THREADS (two CTAs of 16 warps per
// SM) for the large ones, whose products otherwise leave the SM with 4-8 warps to cover shared-memory / L2 latency
template <int NT>
__global__ void __launch_bounds__(NT, NT >= 512 ? 2 : 1)
qpc_admm_kernel(Settings st, QpBuffers qb, int n, int mg, int nbx, long long base, long long B, double* gscratch) {
  extern __shared__ double smem[];
  double* gmat = gscratch ? gscratch + (size_t)blockIdx.x * admm_matrix_doubles(n, mg) : nullptr;
  const long long cnt = qb.list ? (long long)*qb.list_count : B;
  for (long long k = base + blockIdx.x; k < cnt; k += gridDim.x) {
    const long long inst = qb.list ? (long long)qb.list[k] : k;
    AdmmProblem pb;
    pb.P = qb.P + inst * n * n;
    pb.qv = qb.qv + inst * n;
    pb.G = qb.G + inst * mg * n;
    pb.lg = qb.lg + inst * mg;
    pb.ug = qb.ug + inst * mg;
    pb.lb = qb.lb + inst * nbx;
    pb.ub = qb.ub + inst * nbx;
    pb.x = qb.x + inst * n;
    pb.y = qb.y ? qb.y + inst * (mg + nbx) : nullptr;
    pb.status = qb.status + inst;
    pb.iters = qb.iters ? qb.iters + inst : nullptr;
    pb.res = qb.res ? qb.res + 2 * inst : nullptr;
    pb.nfac = qb.nfac ? qb.nfac + inst : nullptr;
    if (qb.warm) {
      pb.x0 = pb.x;
      pb.y0 = pb.y;
      pb.rho_io = qb.rho + inst;
    }
    admm_solve(st, pb, n, mg, nbx, smem, gmat);
  }
}

// register-resident variant (admm_reg.cuh): 4 x TC register tiles, NB column blocks, NB TC >= n + mg positions
template <int TC, int NB>
struct RegTraits {
  static constexpr int MAXT = (NB * TC / 4) * NB;
  // Registers per thread, chosen against the 16K registers of each SM sub-partition (warps of all resident CTAs are
  // spread over the four of them): NB = 16, TC = 5 -> 320 threads, 2 CTAs/SM = 5 warps per sub-partition -> 96;
  // NB = 8: 128 -> 3 CTAs of 160 threads (TC = 10), 4 of 128 (TC = 8); larger tiles run 1 CTA per SM.
#ifndef QPC_REG_SMALL_TILE
#define QPC_REG_SMALL_TILE 128
#endif
  // TC <= 4 (KKT systems of up to 32 positions, CTAs of one or two warps): 96 -> 20 instead of 16 one-warp CTAs per SM;
  // measured on 2^20 Acrobot QPs 21.1 -> 19.5 ms (64 registers: the same there, slower on 4 x 4 tiles, it spills)
#ifndef QPC_REG_TINY_TILE
#define QPC_REG_TINY_TILE 96
#endif
  static constexpr int MAXREG =
      NB == 16 ? 96 : (TC <= 4 ? QPC_REG_TINY_TILE : (TC <= 10 ? QPC_REG_SMALL_TILE : (TC <= 14 ? 255 : 240)));
};
// ELIM: the leading `nel` variables are eliminated inside the solver (admm_reg.cuh); n stays the caller's dimension
template <int TC, int NB, bool ELIM = false>
__global__ void __launch_bounds__((RegTraits<TC, NB>::MAXT)) __maxnreg__((RegTraits<TC, NB>::MAXREG))
qpc_admm_reg_kernel(Settings st, QpBuffers qb, int n, int mg, int nbx, long long base, long long B, int nel = 0) {
  extern __shared__ double smem[];
  const long long cnt = qb.list ? (long long)*qb.list_count : B;
  for (long long k = base + blockIdx.x; k < cnt; k += gridDim.x) {
    const long long inst = qb.list ? (long long)qb.list[k] : k;
    AdmmProblem pb;
    pb.P = qb.P + inst * n * n;
    pb.qv = qb.qv + inst * n;
    pb.G = qb.G + inst * mg * n;
    pb.lg = qb.lg + inst * mg;
    pb.ug = qb.ug + inst * mg;
    pb.lb = qb.lb + inst * nbx;
    pb.ub = qb.ub + inst * nbx;
    pb.x = qb.x + inst * n;
    pb.y = qb.y ? qb.y + inst * (mg + nbx) : nullptr;
    pb.status = qb.status + inst;
    pb.iters = qb.iters ? qb.iters + inst : nullptr;
    pb.res = qb.res ? qb.res + 2 * inst : nullptr;
    pb.nfac = qb.nfac ? qb.nfac + inst : nullptr;
    if (qb.warm) {
      pb.x0 = pb.x;
      pb.y0 = pb.y;
      pb.rho_io = qb.rho + inst;
    }
    RegSolver<TC, NB, ELIM> s;
    s.n = n - (ELIM ? nel : 0);
    s.nel = ELIM ? nel : 0;
    s.nfull = n;
    s.mg = mg;
    s.nbx = nbx;
    s.solve(st, pb, smem);
  }
}

// one warp per QP on the reduced problem (admm_warp.cuh); instances it cannot reduce are appended to fb_list
#ifndef QPC_WARP_CTAS
#define QPC_WARP_CTAS 12
#endif
template <int MG, int NA>
__global__ void __launch_bounds__(32, QPC_WARP_CTAS)
qpc_admm_warp_kernel(Settings st, WarpParams wp, QpBuffers qb, int n, int nbx, long long base, long long B, int paa_diag,
                     int* fb_list, int* fb_count, double* dbg) {
  extern __shared__ double smem[];
  const long long inst = base + blockIdx.x;
  if (inst >= B) return;
  const int mg = MG;
  AdmmProblem pb;
  pb.P = qb.P + inst * n * n;
  pb.qv = qb.qv + inst * n;
  pb.G = qb.G + inst * mg * n;
  pb.lg = qb.lg + inst * mg;
  pb.ug = qb.ug + inst * mg;
  pb.lb = qb.lb + inst * nbx;
  pb.ub = qb.ub + inst * nbx;
  pb.x = qb.x + inst * n;
  pb.y = qb.y ? qb.y + inst * (mg + nbx) : nullptr;
  pb.status = qb.status + inst;
  pb.iters = qb.iters ? qb.iters + inst : nullptr;
  pb.res = qb.res ? qb.res + 2 * inst : nullptr;
  pb.nfac = qb.nfac ? qb.nfac + inst : nullptr;
  if (qb.warm) {
    pb.x0 = pb.x;
    pb.y0 = pb.y;
    pb.rho_io = qb.rho + inst;
  }
  int fallback = 0;
  WarpSolver<MG, NA>::solve(st, wp, pb, n, nbx, paa_diag, smem, fallback, inst == base ? dbg : nullptr);
  if (fallback && threadIdx.x == 0) {
    *pb.status = QPC_WARP_FALLBACK;
    if (pb.iters) *pb.iters = -fallback;  // reason code, visible when the hand-back launch is disabled (QPC_WARP_NOFALLBACK)
    fb_list[atomicAdd(fb_count, 1)] = (int)inst;
  }
}

// status for programs whose QP has no free variable (everything fixed by hard joint tasks)
__global__ void qpc_trivial_status_kernel(int* status, int* iters, double* res, long long base, long long B) {
  long long i = base + blockIdx.x * (long long)blockDim.x + threadIdx.x;
  if (i < B) {
    status[i] = 1;
    if (iters) iters[i] = 0;
    if (res) res[2 * i] = res[2 * i + 1] = 0.0;
  }
}

__global__ void __launch_bounds__(ID_THREADS)
qpc_inverse_dynamics_kernel(const DevProgram* __restrict__ pg, BatchIO io, QpBuffers qb, double* tau, double* vdot,
                            double* wrench, long long base, long long B) {
  extern __shared__ double smem[];
  for (long long inst = base + blockIdx.x; inst < B; inst += gridDim.x) {
    KinSmem s = kin_layout(smem, pg->nb, pg->nq, pg->nv, pg->ndes, pg->ncontacts, pg->N);
    kin_load(pg, io, inst, s);
    for (int i = threadIdx.x; i < pg->ndes; i += blockDim.x) s.des[i] = qb.des[inst * pg->ndes + i];
    __syncthreads();
    kin_forward(pg, s);
    kin_contacts(pg, s);
    double* tdst = tau ? tau + inst * pg->nv : s.q;  // tau is always computed; discard into dead scratch if unwanted
    kin_inverse_dynamics(pg, s, qb.x + inst * pg->n, vdot ? vdot + inst * pg->nv : nullptr,
                         wrench ? wrench + inst * pg->ncontacts * 6 : nullptr, tdst);
  }
}

// ---- the plant of the closed loop: forward dynamics under a soft ground contact (kin.cuh: kin_forward_dynamics) ------------
__global__ void __launch_bounds__(ASM_THREADS)
qpc_forward_dynamics_kernel(const DevProgram* __restrict__ pg, const double* __restrict__ q, const double* __restrict__ v,
                            const double* __restrict__ tau, ContactModel cm, double* vd_out, double* anchor, long long B) {
  extern __shared__ double smem[];
  const int ks = kin_smem_doubles(pg->nb, pg->nq, pg->nv, pg->ndes, pg->ncontacts, pg->N);
  for (long long inst = blockIdx.x; inst < B; inst += gridDim.x) {
    KinSmem s = kin_layout(smem, pg->nb, pg->nq, pg->nv, pg->ndes, pg->ncontacts, pg->N);
    double* M = smem + ks;
    BatchIO io;
    io.q = q;
    io.v = v;
    io.desired = nullptr;
    io.cweight = io.cmaxnf = nullptr;
    io.desired_stride = io.contact_stride = 0;
    kin_load(pg, io, inst, s);
    kin_forward(pg, s);
    kin_composite(pg, s);
    kin_forward_dynamics(pg, s, cm, tau + inst * pg->nv, M, M + pg->nv * pg->nv, M + pg->nv * pg->nv + pg->nv,
                         vd_out + inst * pg->nv, nullptr, anchor ? anchor + inst * 3 * pg->ncontacts : nullptr);
  }
}

// ---- host side ---------------------------------------------------------------------------------------------------------
template <class T>
static cudaError_t grow(T*& p, long long count) {
  if (p) cudaFree(p);
  p = nullptr;
  return cudaMalloc((void**)&p, sizeof(T) * (size_t)(count > 0 ? count : 1));
}

static int ensure_capacity(qpc_controller* c, long long B, long long dstride, long long cstride, long long twstride = 0,
                           long long cgstride = 0) {
  DeviceBuffers& b = c->be.buf;
  const DevProgram& p = c->prog;
  if (B > b.capB) {
    const long long n = p.n, mg = p.mg, nbx = p.nbx;
    CUDA_TRY(grow(b.P, B * n * n));
    CUDA_TRY(grow(b.qv, B * n));
    CUDA_TRY(grow(b.G, B * mg * n));
    // zeroed once: the assembly kernel is their only writer and writes a static pattern (QpBuffers::prezeroed)
    if (n > 0) CUDA_TRY(cudaMemset(b.P, 0, sizeof(double) * (size_t)(B * n * n)));
    if (n > 0 && mg > 0) CUDA_TRY(cudaMemset(b.G, 0, sizeof(double) * (size_t)(B * mg * n)));
    CUDA_TRY(grow(b.lg, B * mg));
    CUDA_TRY(grow(b.ug, B * mg));
    CUDA_TRY(grow(b.lb, B * nbx));
    CUDA_TRY(grow(b.ub, B * nbx));
    CUDA_TRY(grow(b.des, B * p.ndes));
    CUDA_TRY(grow(b.x, B * n));
    CUDA_TRY(grow(b.y, B * (mg + nbx)));
    CUDA_TRY(grow(b.rho, B));
    CUDA_TRY(grow(b.ksave, B * kin_save_doubles(p.nb, p.nv, p.ncontacts, p.N)));
    CUDA_TRY(grow(b.anchor, B * 3 * (p.ncontacts > 0 ? p.ncontacts : 1)));
    CUDA_TRY(cudaMemset(b.anchor, 0, sizeof(double) * (size_t)B * 3 * (p.ncontacts > 0 ? p.ncontacts : 1)));
    CUDA_TRY(cudaMemset(b.rho, 0, sizeof(double) * (size_t)B));
    CUDA_TRY(grow(b.q, B * p.nq));
    CUDA_TRY(grow(b.v, B * p.nv));
    CUDA_TRY(grow(b.tau, B * p.nv));
    CUDA_TRY(grow(b.vdot, B * p.nv));
    CUDA_TRY(grow(b.wrench, B * p.ncontacts * 6));
    CUDA_TRY(grow(b.res, B * 2));
    CUDA_TRY(grow(b.status, B));
    CUDA_TRY(grow(b.iters, B));
    CUDA_TRY(grow(c->be.d_nfac, B));
    CUDA_TRY(grow(c->be.d_fb_list, B));
    if (!c->be.d_fb_count) CUDA_TRY(grow(c->be.d_fb_count, Backend::NSIDE + 1));
    b.capB = B;
    b.cap_desired = 0;
    b.cap_contact = 0;
    b.cap_tw = 0;
    b.cap_twm = 0;
    b.cap_cg = 0;
  }
  if (B * twstride > b.cap_tw) {
    CUDA_TRY(grow(b.tw, B * twstride));
    b.cap_tw = B * twstride;
  }
  if (B * cgstride > b.cap_cg) {
    CUDA_TRY(grow(b.cg, B * cgstride));
    b.cap_cg = B * cgstride;
  }
  if (B * dstride > b.cap_desired) {
    CUDA_TRY(grow(b.desired, B * dstride));
    b.cap_desired = B * dstride;
  }
  if (B * cstride > b.cap_contact) {
    CUDA_TRY(grow(b.cw, B * cstride));
    CUDA_TRY(grow(b.cm, B * cstride));
    b.cap_contact = B * cstride;
  }
  return QPC_OK;
}

static int upload_program(qpc_controller* c) {
  if (!c->be.d_prog) CUDA_TRY(cudaMalloc(&c->be.d_prog, sizeof(DevProgram)));
  CUDA_TRY(cudaMemcpyAsync(c->be.d_prog, &c->prog, sizeof(DevProgram), cudaMemcpyHostToDevice, c->be.stream));
  CUDA_TRY(cudaStreamSynchronize(c->be.stream));
  c->be.dirty = false;
  return QPC_OK;
}

static int launch_grid(long long B) { return (int)(B < (1ll << 30) ? B : (1ll << 30)); }
// grid of an ADMM launch: one CTA per instance, or the bounded persistent grid of list mode
static int admm_grid(const QpBuffers& qb, long long base, long long B) {
  const int g = launch_grid(B - base);
  return (qb.list && qb.list_grid > 0 && qb.list_grid < g) ? qb.list_grid : g;
}

// cudaFuncAttributeMaxDynamicSharedMemorySize is per function and per device, not per controller: it is only ever RAISED
// (per-device high-water mark), so that finalising a controller with a smaller mechanism cannot break the launches of an
// earlier, larger one on the same device
static std::mutex g_attr_mu;
template <class K>
static cudaError_t raise_dyn_smem(K kernel, int bytes, int (&mark)[64]) {
  int dev = 0;
  cudaGetDevice(&dev);
  if (dev < 0 || dev >= 64) return cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes);
  std::lock_guard<std::mutex> lock(g_attr_mu);
  if (bytes <= mark[dev]) return cudaSuccess;
  cudaError_t e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes);
  if (e == cudaSuccess) mark[dev] = bytes;
  return e;
}
static int g_mark_fd[64];
static int g_mark_asm[64], g_mark_asm_se3[64], g_mark_id[64], g_mark_admm128[64], g_mark_admm512[64], g_mark_kinwarp[64], g_mark_idsaved[64];
// ---- ADMM dispatch: register-resident kernel when the KKT matrix fits the register file, shared-memory kernel otherwise
// returns a code TC * 100 + NB, or 0 for the shared-memory kernel
static int reg_tile(int NK) {
  const char* e = getenv("QPC_ADMM_SMEM");
  if (e && e[0] == '1') return 0;
  // QPC_ADMM_TILE=5x16: the 16-lane tiling (320 threads, 2 CTAs/SM, no spills) for 64 < NK <= 80.  Measured on the Atlas
  // workload it is 20 % slower than 10x8 (3 CTAs/SM): twice the warps pay the reduction and the row update.
  const char* t = getenv("QPC_ADMM_TILE");
  if (t && t[0] == '5' && NK > 64 && NK <= 80) return 5 * 100 + 16;
  // TC = 18 (9 warps) cannot launch: 3 warps x 32 x 144+ registers exceed a 16K sub-partition
  static const int sizes[] = {2, 4, 6, 8, 10, 12, 14, 16};
  for (int c : sizes)
    if (admm_reg_positions(c, 8) >= NK) return c * 100 + 8;
  // 128 < NK <= 144 (the dense (68, 71) pair of SURVEY.md 8(d)): 16 lanes per row group, 9 columns per thread,
  // 576 threads, one CTA per SM
  if (NK <= admm_reg_positions(9, 16)) return 9 * 100 + 16;
  return 0;
}
template <int TC, int NB, bool ELIM = false>
static cudaError_t launch_reg(const Settings& st, const QpBuffers& qb, int n, int mg, int nbx, long long base,
                              long long B, cudaStream_t stream, int nel = 0) {
  const int NT = admm_reg_threads(TC, NB);
  // QPC_ADMM_SMEM_PAD=<bytes>: development knob, inflates the dynamic shared memory to cap the CTAs per SM
  static const int pad = [] { const char* e = getenv("QPC_ADMM_SMEM_PAD"); return e ? atoi(e) : 0; }();
  const int bytes = (admm_reg_smem_doubles(TC, NB) + (ELIM ? mg * nel + 4 * nel + NB * (TC + (TC % 4 == 0 ? 2 : 0)) : 0)) * 8 + pad;
  static int configured[64] = {0};  // per device (and per instantiation: this is a function template)
  int dev = 0;
  cudaGetDevice(&dev);
  if (dev < 64 && configured[dev] < bytes) {
    cudaError_t e = cudaFuncSetAttribute(qpc_admm_reg_kernel<TC, NB, ELIM>, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes);
    if (e != cudaSuccess) return e;
    e = cudaFuncSetAttribute(qpc_admm_reg_kernel<TC, NB, ELIM>, cudaFuncAttributePreferredSharedMemoryCarveout,
                             cudaSharedmemCarveoutMaxShared);
    if (e != cudaSuccess) return e;
    configured[dev] = bytes;
  }
  qpc_admm_reg_kernel<TC, NB, ELIM><<<admm_grid(qb, base, B), NT, bytes, stream>>>(st, qb, n, mg, nbx, base, B, nel);
  cudaError_t le = cudaGetLastError();
  if (le != cudaSuccess) {
    cudaFuncAttributes fa;
    if (cudaFuncGetAttributes(&fa, qpc_admm_reg_kernel<TC, NB, ELIM>) == cudaSuccess)
      g_launch_note = " [admm_reg TC=" + std::to_string(TC) + " NB=" + std::to_string(NB) + " threads=" +
                      std::to_string(NT) + " regs=" + std::to_string(fa.numRegs) + " dyn_smem=" + std::to_string(bytes) +
                      " static_smem=" + std::to_string(fa.sharedSizeBytes) + " local=" +
                      std::to_string(fa.localSizeBytes) + "]";
  }
  return le;
}
// Global scratch of the matrices-in-L2 fallback (QPs whose matrices exceed shared memory).  One buffer per device, shared
// by every caller on that device and kept (it only grows, outside steady state: no allocation once a size has been seen).
// Launches that use it are serialised on the buffer through an event -- wait before, record after, both under the mutex
// -- so two streams never work in it at once.
template <class Launch>
static cudaError_t with_admm_global_scratch(int dev, size_t doubles, cudaStream_t stream, Launch launch) {
  static std::mutex mu;
  static double* buf[64] = {nullptr};
  static size_t cap[64] = {0};
  static cudaEvent_t last[64] = {nullptr};
  if (dev < 0 || dev >= 64) return cudaErrorInvalidDevice;
  std::lock_guard<std::mutex> lock(mu);
  cudaError_t e;
  if (!last[dev] && (e = cudaEventCreateWithFlags(&last[dev], cudaEventDisableTiming)) != cudaSuccess) return e;
  if (doubles > cap[dev]) {
    if ((e = cudaDeviceSynchronize()) != cudaSuccess) return e;
    if (buf[dev]) cudaFree(buf[dev]);
    buf[dev] = nullptr;
    cap[dev] = 0;
    if ((e = cudaMalloc((void**)&buf[dev], sizeof(double) * doubles)) != cudaSuccess) return e;
    cap[dev] = doubles;
  } else if ((e = cudaStreamWaitEvent(stream, last[dev], 0)) != cudaSuccess) {
    return e;
  }
  if ((e = launch(buf[dev])) != cudaSuccess) return e;
  return cudaEventRecord(last[dev], stream);
}
// returns cudaSuccess or the launch error; `smem_configured` = the v1 kernel's attribute was already set for this size
// instances [base, B)
static cudaError_t launch_admm(const Settings& st, const QpBuffers& qb, int n, int mg, int nbx, long long base,
                               long long B, cudaStream_t stream, int nel = 0) {
  // fast path: eliminated free variables, tile chosen for the reduced system (see run_tick for the gate); four lanes
  // serve each eliminated column, so 4 nel must not exceed the CTA size
  const int etile = nel > 0 ? reg_tile(n - nel + mg) : 0;
  if (etile && 4 * nel <= admm_reg_threads(etile / 100, etile % 100)) {
    switch (etile) {
      case 608: return launch_reg<6, 8, true>(st, qb, n, mg, nbx, base, B, stream, nel);
      case 808: return launch_reg<8, 8, true>(st, qb, n, mg, nbx, base, B, stream, nel);
      case 1008: return launch_reg<10, 8, true>(st, qb, n, mg, nbx, base, B, stream, nel);
      default: break;  // other sizes: the full system below
    }
  }
  switch (reg_tile(n + mg)) {
    case 516: return launch_reg<5, 16>(st, qb, n, mg, nbx, base, B, stream);
    case 916: return launch_reg<9, 16>(st, qb, n, mg, nbx, base, B, stream);
    case 1008: return launch_reg<10, 8>(st, qb, n, mg, nbx, base, B, stream);
#ifndef QPC_ONLY_ATLAS  /* development builds (register-liveness dumps) instantiate the Atlas tiles only */
    case 208: return launch_reg<2, 8>(st, qb, n, mg, nbx, base, B, stream);
    case 408: return launch_reg<4, 8>(st, qb, n, mg, nbx, base, B, stream);
    case 608: return launch_reg<6, 8>(st, qb, n, mg, nbx, base, B, stream);
    case 808: return launch_reg<8, 8>(st, qb, n, mg, nbx, base, B, stream);
    case 1208: return launch_reg<12, 8>(st, qb, n, mg, nbx, base, B, stream);
    case 1408: return launch_reg<14, 8>(st, qb, n, mg, nbx, base, B, stream);
    case 1608: return launch_reg<16, 8>(st, qb, n, mg, nbx, base, B, stream);
#endif
    default: break;
  }
  const int asmem = admm_smem_doubles(n, mg, nbx) * 8;
  if (asmem <= ADMM_BIG_SMEM) {
    qpc_admm_kernel<ADMM_THREADS><<<admm_grid(qb, base, B), ADMM_THREADS, asmem, stream>>>(st, qb, n, mg, nbx, base, B, nullptr);
    return cudaGetLastError();
  }
  if (asmem <= 227 * 1024) {
    cudaError_t e = raise_dyn_smem(qpc_admm_kernel<ADMM_BIG_THREADS>, asmem, g_mark_admm512);
    if (e != cudaSuccess) return e;
    qpc_admm_kernel<ADMM_BIG_THREADS><<<admm_grid(qb, base, B), ADMM_BIG_THREADS, asmem, stream>>>(st, qb, n, mg, nbx, base, B, nullptr);
    return cudaGetLastError();
  }
  // QPs that fit neither the register file nor shared memory: matrices in a per-CTA global scratch (L2 resident for
  // the grid sizes used), vectors in shared memory; a bounded persistent grid strides over the batch
  const int vsmem = admm_vector_doubles(n, mg, nbx) * 8;
  if (vsmem > 227 * 1024) return cudaErrorInvalidConfiguration;
  int dev = 0, sms = 148;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  const long long grid = B - base < 2ll * sms ? B - base : 2ll * sms;
  cudaError_t e = raise_dyn_smem(qpc_admm_kernel<ADMM_BIG_THREADS>, vsmem, g_mark_admm512);
  if (e != cudaSuccess) return e;
  return with_admm_global_scratch(dev, (size_t)grid * admm_matrix_doubles(n, mg), stream, [&](double* scratch) {
    qpc_admm_kernel<ADMM_BIG_THREADS><<<(int)grid, ADMM_BIG_THREADS, vsmem, stream>>>(st, qb, n, mg, nbx, base, B, scratch);
    return cudaGetLastError();
  });
}

// ---- one-warp-per-QP ADMM (admm_warp.cuh): programs whose unboxed variables are determined by the equality rows ----------
static WarpParams warp_params() {
  static const WarpParams wp = [] {
    WarpParams w;
    const char* e;
    w.kappa = (e = getenv("QPC_WARP_KAPPA")) ? atof(e) : 30.0;
    w.growth = (e = getenv("QPC_WARP_GROWTH")) ? atof(e) : 1.35;
    w.first = (e = getenv("QPC_WARP_FIRST")) ? atoi(e) : 25;
    w.check = (e = getenv("QPC_WARP_CHECK")) ? atoi(e) : 25;  // OSQP's check_termination default; denser checks lose
    w.aitken = (e = getenv("QPC_WARP_AITKEN")) ? atoi(e) : 25;
    w.gather_eps = (e = getenv("QPC_WARP_GATHER_EPS")) ? atof(e) : 1e-6;
    if (!(w.kappa >= 1.0)) w.kappa = 1.0;
    if (!(w.growth > 1.0)) w.growth = 1.35;
    return w;
  }();
  return wp;
}
// (MG, NA) pairs with an instantiation: code MG * 100 + NA, or 0
static int warp_shape(const DevProgram& p) {
  static const bool on = [] { const char* e = getenv("QPC_ADMM_WARP"); return !e || e[0] != '0'; }();
  if (!on || p.nbx < 1 || p.nbx > 32) return 0;
  const int na = p.n - p.nbx;
  if (p.mg == 24 && na == 21) return 2421;  // StandingController on a humanoid with two feet (standing.jl:31-50)
  if (p.mg == 30 && na == 27) return 3027;  // ... plus one weighted 6-row task (e.g. a hand driven by an SE3PDController)
  return 0;
}
static bool paa_is_diagonal(const DevProgram& p) {
  for (int i = 0; i < p.ntasks; i++)
    if (!p.tasks[i].eliminated && p.tasks[i].mode == 2) return false;
  return true;
}
// structure flags of the assembled P for the one-warp kernel (admm_warp.cuh: solve): P_aa diagonal; P_bb block diagonal
// with one N x N block per contact point (contacts.jl:75-79: the cost couples only the multipliers of one point)
static int warp_pflags(const DevProgram& p) {
  bool blocks = p.ncontacts > 0 && p.nbx == p.ncontacts * p.N;
  for (int c = 0; c < p.ncontacts && blocks; c++) blocks = p.contacts[c].col0 == (p.n - p.nbx) + c * p.N;
  return (paa_is_diagonal(p) ? 1 : 0) | ((blocks ? p.N : 0) << 8);
}
template <int MG, int NA>
static cudaError_t launch_warp_t(const Settings& st, const QpBuffers& qb, int n, int nbx, long long base, long long B,
                                 int paa_diag, int* fb_list, int* fb_count, double* dbg, cudaStream_t stream) {
  // QPC_WARP_PAD_SMEM=<bytes>: development knob, pads the dynamic shared memory to lower the resident CTAs per SM
  // (occupancy scan: DESIGN.md 2.4)
  static const int pad = [] { const char* e = getenv("QPC_WARP_PAD_SMEM"); return e ? atoi(e) : 0; }();
  const int bytes = WarpSolver<MG, NA>::SMEM_DOUBLES * 8 + pad;
  static bool configured[64] = {false};
  int dev = 0;
  cudaGetDevice(&dev);
  if (dev < 64 && !configured[dev]) {
    cudaError_t e = cudaFuncSetAttribute(qpc_admm_warp_kernel<MG, NA>, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes);
    if (e != cudaSuccess) return e;
    e = cudaFuncSetAttribute(qpc_admm_warp_kernel<MG, NA>, cudaFuncAttributePreferredSharedMemoryCarveout,
                             cudaSharedmemCarveoutMaxShared);
    if (e != cudaSuccess) return e;
    configured[dev] = true;
  }
  qpc_admm_warp_kernel<MG, NA><<<launch_grid(B - base), 32, bytes, stream>>>(st, warp_params(), qb, n, nbx, base, B, paa_diag,
                                                                            fb_list, fb_count, dbg);
  return cudaGetLastError();
}
static cudaError_t launch_warp(int shape, const Settings& st, const QpBuffers& qb, int n, int nbx, long long base,
                               long long B, int paa_diag, int* fb_list, int* fb_count, double* dbg, cudaStream_t stream) {
  switch (shape) {
    case 2421: return launch_warp_t<24, 21>(st, qb, n, nbx, base, B, paa_diag, fb_list, fb_count, dbg, stream);
    case 3027: return launch_warp_t<30, 27>(st, qb, n, nbx, base, B, paa_diag, fb_list, fb_count, dbg, stream);
    default: return cudaErrorInvalidValue;
  }
}

// tiny mechanisms run the kinematics kernels one warp per instance (kin_warp.cu)
static bool kin_warp_per_instance(const DevProgram& p, int ksm) {
  static const bool on = [] { const char* e = getenv("QPC_KIN_WARP"); return !e || e[0] != '0'; }();
  // QPC_KIN_WARP_BODIES=<n>: development knob, raises the body limit (and lifts the 48 KB limit) of the warp-per-instance form
  static const int maxb = [] { const char* e = getenv("QPC_KIN_WARP_BODIES"); return e ? atoi(e) : KIN_WARP_MAX_BODIES; }();
  if (maxb != KIN_WARP_MAX_BODIES) return on && p.nb <= maxb && KIN_WPC * ksm <= 227 * 1024;
  return on && p.nb <= KIN_WARP_MAX_BODIES && KIN_WPC * ksm <= KIN_WARP_MAX_SMEM;
}
static int configure_kernels(const DevProgram& p) {
  const int ksm = kin_smem_doubles(p.nb, p.nq, p.nv, p.ndes, p.ncontacts, p.N) * 8;
  const int asmem = admm_smem_doubles(p.n, p.mg, p.nbx) * 8;
  if (ksm > 227 * 1024) return qpc_fail(QPC_ERR_LIMIT, "mechanism does not fit the 227 KB shared memory of one CTA");
  CUDA_TRY(raise_dyn_smem(qpc_assemble_kernel<false>, ksm, g_mark_asm));
  CUDA_TRY(raise_dyn_smem(qpc_assemble_kernel<true>, ksm, g_mark_asm_se3));
  CUDA_TRY(raise_dyn_smem(qpc_inverse_dynamics_kernel, ksm, g_mark_id));
  if (kin_warp_per_instance(p, ksm)) {
    int dev = 0;
    cudaGetDevice(&dev);
    std::lock_guard<std::mutex> lock(g_attr_mu);
    if (dev < 0 || dev >= 64 || ksm > g_mark_kinwarp[dev]) {
      CUDA_TRY(kin_warp_configure(ksm));
      if (dev >= 0 && dev < 64) g_mark_kinwarp[dev] = ksm;
    }
  }
  if (asmem <= ADMM_BIG_SMEM) CUDA_TRY(raise_dyn_smem(qpc_admm_kernel<ADMM_THREADS>, asmem, g_mark_admm128));
  if (const int tiny = tiny_thread_class(p); tiny >= 0) {
    std::lock_guard<std::mutex> lock(g_attr_mu);
    CUDA_TRY(tiny_thread_configure(tiny, p));
  }
  {
    const int idsm = kin_id_smem_doubles(p.nb, p.nv, p.ndes, p.ncontacts, p.N) * 8;
    int dev = 0;
    cudaGetDevice(&dev);
    std::lock_guard<std::mutex> lock(g_attr_mu);
    if (KIN_ID_WPC * idsm <= 227 * 1024 && (dev < 0 || dev >= 64 || idsm > g_mark_idsaved[dev])) {
      CUDA_TRY(kin_warp_id_saved_configure(idsm));
      if (dev >= 0 && dev < 64) g_mark_idsaved[dev] = idsm;
    }
  }
  return QPC_OK;
}

// staging buffer of the per-tick matrix weights (QPC_HOST_PTRS); grown outside steady state like the others
static int ensure_twm(qpc_controller* c, long long B, long long stride) {
  DeviceBuffers& b = c->be.buf;
  const long long need = stride ? B * stride : c->prog.nwmat;
  if (need > b.cap_twm) {
    CUDA_TRY(grow(b.twm, need));
    b.cap_twm = need;
  }
  return QPC_OK;
}

static int ensure_time(qpc_controller* c, long long B) {
  DeviceBuffers& b = c->be.buf;
  if (B > b.cap_time) {
    CUDA_TRY(grow(b.time, B));
    b.cap_time = B;
  }
  return QPC_OK;
}

static QpBuffers qp_view(const DeviceBuffers& b) {
  QpBuffers q;
  q.P = b.P;
  q.prezeroed = 1;
  q.qv = b.qv;
  q.G = b.G;
  q.lg = b.lg;
  q.ug = b.ug;
  q.lb = b.lb;
  q.ub = b.ub;
  q.des = b.des;
  q.x = b.x;
  q.y = b.y;
  q.status = b.status;
  q.iters = b.iters;
  q.res = b.res;
  q.rho = b.rho;
  q.ksave = b.ksave;
  q.warm = 0;
  return q;
}

// host buffers of a QPC_HOST_PTRS call: copied chunk by chunk on the chunk's own stream so that PCIe transfers overlap
// the other chunks' kernels
struct HostXfer {
  const qpc_batch_in* in;
  const qpc_batch_out* out;
  long long dstride, cstride;  // 0 = one broadcast row (copied once, before the fork)
  long long twstride = 0, cgstride = 0, twmstride = 0;
  // outputs the epilogue kernel writes straight into the caller's page-locked buffer (no staging, no D2H copy)
  bool direct_tau = false, direct_vdot = false, direct_wrench = false;
};

// Device alias of a page-locked host buffer (cudaHostAlloc / cudaHostRegister / qpc_pin_host_buffer: mapped under unified
// addressing), or nullptr for pageable memory.  QPC_DIRECT_OUT=0 disables the direct output path (A/B).
static double* mapped_alias(const void* host) {
  static const bool on = [] { const char* e = getenv("QPC_DIRECT_OUT"); return !e || e[0] != '0'; }();
  if (!on || !host) return nullptr;
  cudaPointerAttributes a;
  if (cudaPointerGetAttributes(&a, host) != cudaSuccess) {
    cudaGetLastError();
    return nullptr;
  }
  return a.type == cudaMemoryTypeHost ? (double*)a.devicePointer : nullptr;
}

// the tick on device pointers; asynchronous on `stream`
static int run_tick(qpc_controller* c, long long B, const BatchIO& io, double* tau, double* vdot, double* wrench,
                    int* status, int* iters, double* res, int* nfac, cudaStream_t stream,
                    const HostXfer* hx = nullptr) {
  const DevProgram& p = c->prog;
  const DevProgram* dp = (const DevProgram*)c->be.d_prog;
  QpBuffers qb = qp_view(c->be.buf);
  if (status) qb.status = status;
  if (iters) qb.iters = iters;
  if (res) qb.res = res;
  qb.nfac = nfac ? nfac : c->be.d_nfac;
  qb.warm = c->be.warm_start ? 1 : 0;
  const bool prof = c->be.profiling;
  const int ksm = kin_smem_doubles(p.nb, p.nq, p.nv, p.ndes, p.ncontacts, p.N) * 8;
  DeviceBuffers& hb = c->be.buf;
  auto tick_range = [&](long long lo, long long hi, cudaStream_t s, bool timed, int chunk) -> int {
    const int grid = launch_grid(hi - lo);
    const size_t cnt = (size_t)(hi - lo);
    if (hx) {  // inputs of this chunk, host -> device staging
      CUDA_TRY(cudaMemcpyAsync(hb.q + lo * p.nq, hx->in->q + lo * p.nq, sizeof(double) * cnt * p.nq, cudaMemcpyHostToDevice, s));
      CUDA_TRY(cudaMemcpyAsync(hb.v + lo * p.nv, hx->in->v + lo * p.nv, sizeof(double) * cnt * p.nv, cudaMemcpyHostToDevice, s));
      if (hx->in->desired && hx->dstride)
        CUDA_TRY(cudaMemcpyAsync(hb.desired + lo * hx->dstride, hx->in->desired + lo * hx->dstride,
                                 sizeof(double) * cnt * hx->dstride, cudaMemcpyHostToDevice, s));
      if (hx->in->contact_weight && hx->cstride) {
        CUDA_TRY(cudaMemcpyAsync(hb.cw + lo * hx->cstride, hx->in->contact_weight + lo * hx->cstride,
                                 sizeof(double) * cnt * hx->cstride, cudaMemcpyHostToDevice, s));
        CUDA_TRY(cudaMemcpyAsync(hb.cm + lo * hx->cstride, hx->in->contact_maxnormalforce + lo * hx->cstride,
                                 sizeof(double) * cnt * hx->cstride, cudaMemcpyHostToDevice, s));
      }
      if (hx->in->task_weight && hx->twstride)
        CUDA_TRY(cudaMemcpyAsync(hb.tw + lo * hx->twstride, hx->in->task_weight + lo * hx->twstride,
                                 sizeof(double) * cnt * hx->twstride, cudaMemcpyHostToDevice, s));
      if (hx->in->task_weight_matrix && hx->twmstride)
        CUDA_TRY(cudaMemcpyAsync(hb.twm + lo * hx->twmstride, hx->in->task_weight_matrix + lo * hx->twmstride,
                                 sizeof(double) * cnt * hx->twmstride, cudaMemcpyHostToDevice, s));
      if (hx->in->contact_geometry && hx->cgstride)
        CUDA_TRY(cudaMemcpyAsync(hb.cg + lo * hx->cgstride, hx->in->contact_geometry + lo * hx->cgstride,
                                 sizeof(double) * cnt * hx->cgstride, cudaMemcpyHostToDevice, s));
    }
    if (timed) cudaEventRecord(c->be.ev[0], s);
    // Tiny mechanisms: the whole tick in ONE kernel, one thread per instance (tiny_thread.cu; QPC_TINY_THREAD=0 takes the
    // three-kernel path instead).  The stage events then bracket the single launch: all of its time is reported as `admm`.
    static const bool tiny_on = [] { const char* e = getenv("QPC_TINY_THREAD"); return !e || e[0] != '0'; }();
    const int tiny = tiny_on ? tiny_thread_class(p) : -1;
    if (tiny >= 0) {
      if (timed) cudaEventRecord(c->be.ev[1], s);
      CUDA_TRY(tiny_thread_tick(tiny, p, dp, io, qb, tau, vdot, wrench, lo, hi, s));
      if (timed) {
        cudaEventRecord(c->be.ev[2], s);
        cudaEventRecord(c->be.ev[3], s);
      }
      c->be.launches += 1;
    } else {
    const bool kwarp = kin_warp_per_instance(p, ksm);
    if (kwarp) CUDA_TRY(kin_warp_assemble(dp, io, qb, lo, hi, ksm, s, p.nse3 > 0));
    else if (p.nse3) qpc_assemble_kernel<true><<<grid, ASM_THREADS, ksm, s>>>(dp, io, qb, lo, hi);
    else {
      // QPC_ASM_PAD_SMEM=<bytes>: development knob (occupancy scan of the assembly kernel)
      static const int pad = [] { const char* e = getenv("QPC_ASM_PAD_SMEM"); return e ? atoi(e) : 0; }();
      if (pad) cudaFuncSetAttribute(qpc_assemble_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, ksm + pad);
      qpc_assemble_kernel<false><<<grid, ASM_THREADS, ksm + pad, s>>>(dp, io, qb, lo, hi);
    }
    if (timed) cudaEventRecord(c->be.ev[1], s);
    if (p.n > 0) {
      // Fast path with the diagonal-cost free variables eliminated: only for tolerances above the floor that form puts
      // on the primal residual, static task weights, and unless QPC_ADMM_ELIM=0
      static const bool elim_on = [] { const char* e = getenv("QPC_ADMM_ELIM"); return !e || e[0] != '0'; }();
      const int nel = (elim_on && c->be.elimination && p.nel > 0 && p.settings.eps_abs >= 1e-6 && !io.tweight) ? p.nel : 0;
      const int wshape = c->be.warp_admm ? warp_shape(p) : 0;
      if (wshape) {
        // one warp per QP on the reduced problem; what it hands back (rank-deficient reductions, ...) goes through the
        // register-tile kernel in list mode
        int* cnt = c->be.d_fb_count + chunk;
        CUDA_TRY(cudaMemsetAsync(cnt, 0, sizeof(int), s));
        // development knobs: QPC_WARP_DEBUG=<file> dumps the reduced problem of the chunk's first instance,
        // QPC_WARP_NOFALLBACK=1 leaves handed-back instances unsolved (status -99, iters = -reason)
        static const char* dbg_path = getenv("QPC_WARP_DEBUG");
        static const bool no_fb = [] { const char* e = getenv("QPC_WARP_NOFALLBACK"); return e && e[0] == '1'; }();
        double* dbg = nullptr;
        const int ndbg = 4096;
        if (dbg_path) {
          CUDA_TRY(cudaMalloc((void**)&dbg, sizeof(double) * ndbg));
          CUDA_TRY(cudaMemset(dbg, 0, sizeof(double) * ndbg));
        }
        CUDA_TRY(launch_warp(wshape, p.settings, qb, p.n, p.nbx, lo, hi, warp_pflags(p), c->be.d_fb_list + lo,
                             cnt, dbg, s));
        if (dbg) {
          std::vector<double> hd(ndbg);
          CUDA_TRY(cudaStreamSynchronize(s));
          CUDA_TRY(cudaMemcpy(hd.data(), dbg, sizeof(double) * ndbg, cudaMemcpyDeviceToHost));
          cudaFree(dbg);
          if (FILE* f = fopen(dbg_path, "wb")) {
            fwrite(hd.data(), sizeof(double), ndbg, f);
            fclose(f);
          }
        }
        if (!no_fb) {
          QpBuffers ql = qb;
          ql.list = c->be.d_fb_list + lo;
          ql.list_count = cnt;
          ql.list_grid = 148;
          CUDA_TRY(launch_admm(p.settings, ql, p.n, p.mg, p.nbx, 0, hi - lo, s, nel));
          c->be.launches += 1;
        }
      } else {
        CUDA_TRY(launch_admm(p.settings, qb, p.n, p.mg, p.nbx, lo, hi, s, nel));
      }
    }
    else
      qpc_trivial_status_kernel<<<(unsigned)((hi - lo + 255) / 256), 256, 0, s>>>(qb.status, qb.iters, qb.res, lo, hi);
    if (timed) cudaEventRecord(c->be.ev[2], s);
    // inverse dynamics from the state the assembly kernel saved (QPC_ID_SAVED=0: recompute the forward sweep instead)
    static const bool id_saved = [] { const char* e = getenv("QPC_ID_SAVED"); return !e || e[0] != '0'; }();
    const int idsm = kin_id_smem_doubles(p.nb, p.nv, p.ndes, p.ncontacts, p.N) * 8;
    if (id_saved && qb.ksave && KIN_ID_WPC * idsm <= 227 * 1024) {
      CUDA_TRY(kin_warp_id_saved(dp, qb, tau, vdot, wrench, lo, hi, idsm, s));
    } else if (kwarp) {
      CUDA_TRY(kin_warp_inverse_dynamics(dp, io, qb, tau, vdot, wrench, lo, hi, ksm, s));
    } else {
      qpc_inverse_dynamics_kernel<<<grid, ID_THREADS, ksm, s>>>(dp, io, qb, tau, vdot, wrench, lo, hi);
    }
    if (timed) cudaEventRecord(c->be.ev[3], s);
    c->be.launches += 3;
    }
    if (hx) {  // results of this chunk, device staging -> host
      const qpc_batch_out* o = hx->out;
      const int nc6 = p.ncontacts * 6;
      if (o->tau && !hx->direct_tau)
        CUDA_TRY(cudaMemcpyAsync(o->tau + lo * p.nv, tau + lo * p.nv, sizeof(double) * cnt * p.nv, cudaMemcpyDeviceToHost, s));
      if (o->vdot && !hx->direct_vdot)
        CUDA_TRY(cudaMemcpyAsync(o->vdot + lo * p.nv, vdot + lo * p.nv, sizeof(double) * cnt * p.nv, cudaMemcpyDeviceToHost, s));
      if (o->wrench && !hx->direct_wrench)
        CUDA_TRY(cudaMemcpyAsync(o->wrench + lo * nc6, wrench + lo * nc6, sizeof(double) * cnt * nc6, cudaMemcpyDeviceToHost, s));
      if (o->status) CUDA_TRY(cudaMemcpyAsync(o->status + lo, qb.status + lo, sizeof(int) * cnt, cudaMemcpyDeviceToHost, s));
      if (o->iters) CUDA_TRY(cudaMemcpyAsync(o->iters + lo, qb.iters + lo, sizeof(int) * cnt, cudaMemcpyDeviceToHost, s));
      if (o->residuals) CUDA_TRY(cudaMemcpyAsync(o->residuals + 2 * lo, qb.res + 2 * lo, sizeof(double) * 2 * cnt, cudaMemcpyDeviceToHost, s));
      if (o->factorizations) CUDA_TRY(cudaMemcpyAsync(o->factorizations + lo, qb.nfac + lo, sizeof(int) * cnt, cudaMemcpyDeviceToHost, s));
    }
    return QPC_OK;
  };
  // Large batches are cut into NSIDE contiguous chunks on side streams (forked from / joined to the caller's stream):
  // the GPU then overlaps chunk k's inverse dynamics, chunk k+1's assembly and the tail of chunk k's ADMM kernel with
  // the bulk of the next ADMM kernel.  Results do not depend on the chunking (instances are independent).
  // QPC_NCHUNK=<n>: development knob (1 .. NSIDE)
  static const int nchunk_env = [] { const char* e = getenv("QPC_NCHUNK"); return e ? atoi(e) : 0; }();
  int nchunk = (!prof && B >= 4096 && c->be.side[0]) ? Backend::NSIDE : 1;
  if (nchunk > 1 && nchunk_env >= 1 && nchunk_env <= Backend::NSIDE) nchunk = nchunk_env;
  if (nchunk == 1) {
    int rc = tick_range(0, B, stream, prof, Backend::NSIDE);
    if (rc) return rc;
  } else {
    CUDA_TRY(cudaEventRecord(c->be.fork, stream));
    // host buffers: the first chunk's H2D copies and the last chunk's D2H copies are the only ones no kernel hides, so
    // those two chunks are smaller (15 % of the batch each with four chunks)
    auto cut = [&](int k) -> long long {
      if (!hx || nchunk != 4) return B * k / nchunk;
      // QPC_CHUNK_FRACS="a,b,c": development knob for the three inner cut points
      static const std::array<double, 5> frac = [] {
        std::array<double, 5> f = {0.0, 0.15, 0.5, 0.85, 1.0};
        if (const char* e = getenv("QPC_CHUNK_FRACS")) sscanf(e, "%lf,%lf,%lf", &f[1], &f[2], &f[3]);
        return f;
      }();
      return k == nchunk ? B : (long long)(B * frac[k]);
    };
    for (int k = 0; k < nchunk; k++) {
      const long long lo = cut(k), hi = cut(k + 1);
      CUDA_TRY(cudaStreamWaitEvent(c->be.side[k], c->be.fork, 0));
      int rc = tick_range(lo, hi, c->be.side[k], false, k);
      if (rc) return rc;
      CUDA_TRY(cudaEventRecord(c->be.join[k], c->be.side[k]));
      CUDA_TRY(cudaStreamWaitEvent(stream, c->be.join[k], 0));
    }
  }
  CUDA_TRY(cudaGetLastError());
  return QPC_OK;
}

extern "C" {

int qpc_device_count(void) {
  int n = 0;
  if (cudaGetDeviceCount(&n) != cudaSuccess) {
    cudaGetLastError();
    return 0;
  }
  return n;
}

int qpc_finalize(qpc_controller* c, int32_t device) {
  if (!c) return qpc_fail(QPC_ERR_ARG, "null controller");
  if (c->finalized) return QPC_OK;
  std::string err = compile_program(c->hc, c->prog);
  if (!err.empty()) return qpc_fail(QPC_ERR_LIMIT, err);
  c->be.device = device;
  CUDA_TRY(cudaSetDevice(device));
  CUDA_TRY(cudaStreamCreateWithFlags(&c->be.stream, cudaStreamNonBlocking));
  for (int k = 0; k < Backend::NSIDE; k++) {
    CUDA_TRY(cudaStreamCreateWithFlags(&c->be.side[k], cudaStreamNonBlocking));
    CUDA_TRY(cudaEventCreateWithFlags(&c->be.join[k], cudaEventDisableTiming));
  }
  CUDA_TRY(cudaEventCreateWithFlags(&c->be.fork, cudaEventDisableTiming));
  int rc = configure_kernels(c->prog);
  if (rc) return rc;
  rc = upload_program(c);
  if (rc) return rc;
  c->finalized = true;
  return QPC_OK;
}

void qpc_controller_destroy(qpc_controller* c) {
  if (!c) return;
  if (c->finalized) {
    cudaSetDevice(c->be.device);
    DeviceBuffers& b = c->be.buf;
    void* ptrs[] = {b.P, b.qv, b.G, b.lg, b.ug, b.lb, b.ub, b.des, b.x, b.y, b.q, b.v, b.desired, b.cw,
                    b.cm, b.tau, b.vdot, b.wrench, b.res, b.status, b.iters, c->be.d_prog, c->be.d_nfac, b.rho, b.ksave, b.anchor, b.tw, b.cg, b.twm, b.time, c->be.d_fb_list, c->be.d_fb_count};
    for (int i = 0; i < 4; i++)
      if (c->be.ev[i]) cudaEventDestroy(c->be.ev[i]);
    for (void* p : ptrs)
      if (p) cudaFree(p);
    if (c->be.stream) cudaStreamDestroy(c->be.stream);
    for (int k = 0; k < Backend::NSIDE; k++) {
      if (c->be.side[k]) cudaStreamDestroy(c->be.side[k]);
      if (c->be.join[k]) cudaEventDestroy(c->be.join[k]);
    }
    if (c->be.fork) cudaEventDestroy(c->be.fork);
  }
  delete c;
}

int qpc_reserve(qpc_controller* c, int64_t B) {
  if (!c || !c->finalized) return qpc_fail(QPC_ERR_STATE, "controller not finalized");
  std::lock_guard<std::mutex> lock(c->be.mu);
  CUDA_TRY(cudaSetDevice(c->be.device));
  return ensure_capacity(c, B, c->prog.ndes, c->prog.ncontacts);
}

int64_t qpc_launch_count(const qpc_controller* c) { return c ? (int64_t)c->be.launches.load() : 0; }

// validation of the optional per-tick Parameter arrays (task weights, contact geometry)
static int check_tick_parameters(const DevProgram& p, const qpc_batch_in* in) {
  if (!in) return QPC_OK;
  if (in->task_weight && in->task_weight_stride != 0 && in->task_weight_stride < p.ntasks)
    return qpc_fail(QPC_ERR_ARG, "task_weight_stride smaller than the number of tasks");
  if (in->contact_geometry && in->contact_geometry_stride != 0 && in->contact_geometry_stride < 7 * p.ncontacts)
    return qpc_fail(QPC_ERR_ARG, "contact_geometry_stride smaller than 7 x the number of contacts");
  if (in->task_weight_matrix && in->task_weight_matrix_stride != 0 && in->task_weight_matrix_stride < p.nwmat)
    return qpc_fail(QPC_ERR_ARG, "task_weight_matrix_stride smaller than the matrix-weight storage (sum of dim^2)");
  return QPC_OK;
}
// device pointers: use the caller's arrays; host pointers: broadcast rows are copied here, per-instance rows by the
// chunk that owns them (HostXfer) or, for qpc_step_batch, as one copy
static int stage_tick_parameters(qpc_controller* c, long long B, const qpc_batch_in* in, bool host, bool copy_all,
                                 BatchIO& io, cudaStream_t s) {
  const DevProgram& p = c->prog;
  DeviceBuffers& b = c->be.buf;
  io.tweight = io.cgeom = io.twmat = io.time = nullptr;
  io.tweight_stride = io.cgeom_stride = io.twmat_stride = io.time_stride = 0;
  if (!in) return QPC_OK;
  if (in->time && p.nse3 > 0) {  // 8 bytes per instance: copied whole, ahead of the chunks
    io.time_stride = in->time_stride ? 1 : 0;
    io.time = in->time;
    if (host) {
      const long long cnt = io.time_stride ? B : 1;
      int rc = ensure_time(c, cnt);
      if (rc) return rc;
      CUDA_TRY(cudaMemcpyAsync(b.time, in->time, sizeof(double) * (size_t)cnt, cudaMemcpyHostToDevice, s));
      io.time = b.time;
    }
  }
  if (in->task_weight_matrix && p.nwmat > 0) {
    io.twmat_stride = in->task_weight_matrix_stride;
    io.twmat = in->task_weight_matrix;
    if (host) {
      int rc = ensure_twm(c, B, io.twmat_stride);
      if (rc) return rc;
      const long long cnt = io.twmat_stride ? (copy_all ? B * io.twmat_stride : 0) : p.nwmat;
      if (cnt) CUDA_TRY(cudaMemcpyAsync(b.twm, in->task_weight_matrix, sizeof(double) * (size_t)cnt, cudaMemcpyHostToDevice, s));
      io.twmat = b.twm;
    }
  }
  if (in->task_weight) {
    io.tweight_stride = in->task_weight_stride;
    io.tweight = in->task_weight;
    if (host) {
      const long long cnt = io.tweight_stride ? (copy_all ? B * io.tweight_stride : 0) : p.ntasks;
      if (cnt) CUDA_TRY(cudaMemcpyAsync(b.tw, in->task_weight, sizeof(double) * (size_t)cnt, cudaMemcpyHostToDevice, s));
      io.tweight = b.tw;
    }
  }
  if (in->contact_geometry) {
    io.cgeom_stride = in->contact_geometry_stride;
    io.cgeom = in->contact_geometry;
    if (host) {
      const long long cnt = io.cgeom_stride ? (copy_all ? B * io.cgeom_stride : 0) : 7 * p.ncontacts;
      if (cnt) CUDA_TRY(cudaMemcpyAsync(b.cg, in->contact_geometry, sizeof(double) * (size_t)cnt, cudaMemcpyHostToDevice, s));
      io.cgeom = b.cg;
    }
  }
  return QPC_OK;
}

int qpc_solve_batch(qpc_controller* c, int64_t B, const qpc_batch_in* in, const qpc_batch_out* out, int32_t flags,
                    void* stream_) {
  if (!c || !c->finalized) return qpc_fail(QPC_ERR_STATE, "controller not finalized");
  if (!in || !out || !in->q || !in->v || B < 0) return qpc_fail(QPC_ERR_ARG, "qpc_solve_batch: bad arguments");
  if (B == 0) return QPC_OK;
  const DevProgram& p = c->prog;
  if (in->desired && in->desired_stride != 0 && in->desired_stride < p.ndes)
    return qpc_fail(QPC_ERR_ARG, "desired_stride smaller than the number of desired values");
  if ((in->contact_weight == nullptr) != (in->contact_maxnormalforce == nullptr))
    return qpc_fail(QPC_ERR_ARG, "contact_weight and contact_maxnormalforce must be given together");
  if (in->contact_weight && in->contact_stride != 0 && in->contact_stride < p.ncontacts)
    return qpc_fail(QPC_ERR_ARG, "contact_stride smaller than the number of contacts");
  std::lock_guard<std::mutex> lock(c->be.mu);
  CUDA_TRY(cudaSetDevice(c->be.device));
  if (c->be.dirty) {
    int rc = upload_program(c);
    if (rc) return rc;
  }
  const long long dstride = in->desired ? in->desired_stride : 0, cstride = in->contact_weight ? in->contact_stride : 0;
  const long long drows = dstride ? B : 1, crows = cstride ? B : 1;
  int rc = check_tick_parameters(p, in);
  if (rc) return rc;
  const long long twstride = in->task_weight ? in->task_weight_stride : 0;
  const long long cgstride = in->contact_geometry ? in->contact_geometry_stride : 0;
  rc = ensure_capacity(c, B, in->desired ? (dstride ? dstride : p.ndes) : 0,
                       in->contact_weight ? (cstride ? cstride : p.ncontacts) : 0,
                       in->task_weight ? (twstride ? twstride : p.ntasks) : 0,
                       in->contact_geometry ? (cgstride ? cgstride : 7 * p.ncontacts) : 0);
  if (rc) return rc;
  DeviceBuffers& b = c->be.buf;
  BatchIO io;
  io.desired_stride = dstride;
  io.contact_stride = cstride;
  if (flags == QPC_DEVICE_PTRS) {
    cudaStream_t stream = (cudaStream_t)stream_;
    io.q = in->q;
    io.v = in->v;
    io.desired = in->desired;
    io.cweight = in->contact_weight;
    io.cmaxnf = in->contact_maxnormalforce;
    rc = stage_tick_parameters(c, B, in, false, false, io, stream);
    if (rc) return rc;
    return run_tick(c, B, io, out->tau, out->vdot, out->wrench, out->status, out->iters, out->residuals,
                    out->factorizations, stream);
  }
  cudaStream_t s = c->be.stream;
  io.q = b.q;
  io.v = b.v;
  io.desired = nullptr;
  io.cweight = io.cmaxnf = nullptr;
  if (in->desired) {
    if (!dstride) CUDA_TRY(cudaMemcpyAsync(b.desired, in->desired, sizeof(double) * p.ndes, cudaMemcpyHostToDevice, s));
    io.desired = b.desired;
  }
  if (in->contact_weight) {
    if (!cstride) {
      CUDA_TRY(cudaMemcpyAsync(b.cw, in->contact_weight, sizeof(double) * p.ncontacts, cudaMemcpyHostToDevice, s));
      CUDA_TRY(cudaMemcpyAsync(b.cm, in->contact_maxnormalforce, sizeof(double) * p.ncontacts, cudaMemcpyHostToDevice, s));
    }
    io.cweight = b.cw;
    io.cmaxnf = b.cm;
  }
  (void)drows;
  (void)crows;
  rc = stage_tick_parameters(c, B, in, true, false, io, s);
  if (rc) return rc;
  HostXfer hx{in, out, dstride, cstride, twstride, cgstride,
              (in->task_weight_matrix && p.nwmat > 0) ? in->task_weight_matrix_stride : 0};
  // Page-locked output buffers: the inverse-dynamics epilogue writes tau / vdot / wrenches straight into them (posted
  // PCIe writes spread over the tick) instead of a staging buffer plus a D2H copy per chunk -- the copies of the chunks
  // that finish together at the end of the tick were 0.17 ms of exposed time.  Pageable buffers keep the staged path.
  double *tau_d = b.tau, *vdot_d = b.vdot, *wrench_d = b.wrench;
  if (double* m = mapped_alias(out->tau)) tau_d = m, hx.direct_tau = true;
  if (double* m = mapped_alias(out->vdot)) vdot_d = m, hx.direct_vdot = true;
  if (double* m = mapped_alias(out->wrench)) wrench_d = m, hx.direct_wrench = true;
  rc = run_tick(c, B, io, tau_d, vdot_d, wrench_d, b.status, b.iters, b.res, nullptr, s, &hx);
  if (rc) return rc;
  CUDA_TRY(cudaStreamSynchronize(s));
  return QPC_OK;
}

// ---- single-process multi-GPU entry point: contiguous shards, one host thread per controller / device, no collective -----
int qpc_solve_batch_multi(qpc_controller* const* ctrls, int32_t nctrl, int64_t B, const qpc_batch_in* in,
                          const qpc_batch_out* out) {
  if (!ctrls || nctrl < 1 || !in || !out || !in->q || !in->v || B < 0)
    return qpc_fail(QPC_ERR_ARG, "qpc_solve_batch_multi: bad arguments");
  for (int k = 0; k < nctrl; k++) {
    if (!ctrls[k] || !ctrls[k]->finalized) return qpc_fail(QPC_ERR_STATE, "qpc_solve_batch_multi: controller not finalized");
    const DevProgram &a = ctrls[0]->prog, &b = ctrls[k]->prog;
    if (a.nq != b.nq || a.nv != b.nv || a.ndes != b.ndes || a.ncontacts != b.ncontacts || a.ntasks != b.ntasks ||
        a.n != b.n || a.mg != b.mg || a.nbx != b.nbx || a.nwmat != b.nwmat)
      return qpc_fail(QPC_ERR_ARG, "qpc_solve_batch_multi: the controllers must be replicas of one program");
    for (int j = 0; j < k; j++)
      if (ctrls[j] == ctrls[k]) return qpc_fail(QPC_ERR_ARG, "qpc_solve_batch_multi: a controller handle appears twice");
  }
  if (B == 0) return QPC_OK;
  const DevProgram& p = ctrls[0]->prog;
  std::vector<int> rc(nctrl, QPC_OK);
  std::vector<std::string> msg(nctrl);
  std::vector<std::thread> th;
  auto shard = [&](int k) {
    const long long base = B / nctrl, extra = B % nctrl;
    const long long lo = k * base + (k < extra ? k : extra), cnt = base + (k < extra ? 1 : 0);
    if (cnt == 0) return;
    qpc_batch_in i2 = *in;
    qpc_batch_out o2 = *out;
    i2.q += lo * p.nq;
    i2.v += lo * p.nv;
    if (i2.desired) i2.desired += lo * i2.desired_stride;
    if (i2.contact_weight) i2.contact_weight += lo * i2.contact_stride;
    if (i2.contact_maxnormalforce) i2.contact_maxnormalforce += lo * i2.contact_stride;
    if (i2.task_weight) i2.task_weight += lo * i2.task_weight_stride;
    if (i2.contact_geometry) i2.contact_geometry += lo * i2.contact_geometry_stride;
    if (i2.task_weight_matrix) i2.task_weight_matrix += lo * i2.task_weight_matrix_stride;
    if (i2.time && i2.time_stride) i2.time += lo;
    if (o2.tau) o2.tau += lo * p.nv;
    if (o2.vdot) o2.vdot += lo * p.nv;
    if (o2.wrench) o2.wrench += lo * p.ncontacts * 6;
    if (o2.status) o2.status += lo;
    if (o2.iters) o2.iters += lo;
    if (o2.residuals) o2.residuals += 2 * lo;
    if (o2.factorizations) o2.factorizations += lo;
    rc[k] = qpc_solve_batch(ctrls[k], cnt, &i2, &o2, QPC_HOST_PTRS, nullptr);
    if (rc[k] != QPC_OK) msg[k] = qpc_last_error();  // the error string is thread-local: carry it to the caller's thread
  };
  for (int k = 1; k < nctrl; k++) th.emplace_back(shard, k);
  shard(0);
  for (auto& t : th) t.join();
  for (int k = 0; k < nctrl; k++)
    if (rc[k] != QPC_OK) return qpc_fail(rc[k], "qpc_solve_batch_multi: shard " + std::to_string(k) + ": " + msg[k]);
  return QPC_OK;
}

// Page-locks a caller-owned host buffer (cudaHostRegister, portable across devices) so that the QPC_HOST_PTRS copies are
// asynchronous DMA transfers that overlap the kernels of other chunks; pageable buffers work too, the driver then stages
// them and every copy blocks the calling thread.  The caller unpins before freeing the buffer.
int qpc_pin_host_buffer(void* ptr, int64_t bytes) {
  if (!ptr || bytes <= 0) return qpc_fail(QPC_ERR_ARG, "qpc_pin_host_buffer: bad arguments");
  cudaError_t e = cudaHostRegister(ptr, (size_t)bytes, cudaHostRegisterPortable);
  if (e == cudaErrorHostMemoryAlreadyRegistered) {
    cudaGetLastError();
    return QPC_OK;
  }
  CUDA_TRY(e);
  return QPC_OK;
}
int qpc_unpin_host_buffer(void* ptr) {
  if (!ptr) return qpc_fail(QPC_ERR_ARG, "qpc_unpin_host_buffer: null pointer");
  cudaError_t e = cudaHostUnregister(ptr);
  if (e == cudaErrorHostMemoryNotRegistered) {
    cudaGetLastError();
    return QPC_OK;
  }
  CUDA_TRY(e);
  return QPC_OK;
}

int qpc_set_warm_start(qpc_controller* c, int32_t on) {
  if (!c || !c->finalized) return qpc_fail(QPC_ERR_STATE, "controller not finalized");
  std::lock_guard<std::mutex> lock(c->be.mu);
  c->be.warm_start = on != 0;
  return QPC_OK;
}

int qpc_set_admm_elimination(qpc_controller* c, int32_t on) {
  if (!c || !c->finalized) return qpc_fail(QPC_ERR_STATE, "controller not finalized");
  std::lock_guard<std::mutex> lock(c->be.mu);
  c->be.elimination = on != 0;
  return QPC_OK;
}

int qpc_set_admm_warp(qpc_controller* c, int32_t on) {
  if (!c || !c->finalized) return qpc_fail(QPC_ERR_STATE, "controller not finalized");
  std::lock_guard<std::mutex> lock(c->be.mu);
  c->be.warp_admm = on != 0;
  return QPC_OK;
}

int qpc_admm_warp(const qpc_controller* c) {
  if (!c || !c->finalized) return 0;
  return (c->be.warp_admm && warp_shape(c->prog)) ? 1 : 0;
}

int qpc_admm_eliminated(const qpc_controller* c) {
  if (!c || !c->finalized) return 0;
  const DevProgram& p = c->prog;
  if (!c->be.elimination || p.nel <= 0 || p.settings.eps_abs < 1e-6) return 0;
  const int etile = reg_tile(p.n - p.nel + p.mg);
  return (etile && 4 * p.nel <= admm_reg_threads(etile / 100, etile % 100) &&
          (etile == 608 || etile == 808 || etile == 1008)) ? p.nel : 0;
}

int qpc_reset_warm_start(qpc_controller* c) {
  if (!c || !c->finalized) return qpc_fail(QPC_ERR_STATE, "controller not finalized");
  std::lock_guard<std::mutex> lock(c->be.mu);
  CUDA_TRY(cudaSetDevice(c->be.device));
  if (c->be.buf.rho && c->be.buf.capB > 0) {
    if (c->be.buf.anchor)
      CUDA_TRY(cudaMemsetAsync(c->be.buf.anchor, 0, sizeof(double) * (size_t)c->be.buf.capB * 3 *
                               (c->prog.ncontacts > 0 ? c->prog.ncontacts : 1), c->be.stream));
    CUDA_TRY(cudaMemsetAsync(c->be.buf.rho, 0, sizeof(double) * (size_t)c->be.buf.capB, c->be.stream));
    CUDA_TRY(cudaStreamSynchronize(c->be.stream));
  }
  return QPC_OK;
}

// ---- closed loop: tick, then advance (q, v) in place with the commanded accelerations -------------------------------------
// Semi-implicit Euler on the configuration manifold: v+ = v + dt vd; revolute / prismatic q+ = q + dt v+; quaternion
// floating joint (q = (w,x,y,z,p), v = (omega, v) in body frame): quat+ = quat * exp(dt omega+ / 2), p+ = p + dt R v+.
// Instances whose solve was rejected (checkstatus, momentum.jl:83-91) are left where they are.
__global__ void qpc_integrate_kernel(const DevProgram* __restrict__ pg, double* q, double* v, const double* __restrict__ vdot,
                                     const int* __restrict__ status, double dt, long long B) {
  const long long inst = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  if (inst >= B) return;
  if (status && status[inst] != 1 && status[inst] != 2) return;
  const int nq = pg->nq, nv = pg->nv;
  double* qi = q + inst * nq;
  double* vi = v + inst * nv;
  const double* vd = vdot + inst * nv;
  for (int i = 0; i < nv; i++) vi[i] += dt * vd[i];
  for (int b = 0; b < pg->nb; b++) {
    const int jt = pg->jtype[b], qo = pg->qoff[b], vo = pg->voff[b];
    if (jt == 0 || jt == 1) {
      qi[qo] += dt * vi[vo];
    } else if (jt == 2) {
      double R[9];
      quat_to_rot(qi[qo], qi[qo + 1], qi[qo + 2], qi[qo + 3], R);
      const V3 om = ld3(vi + vo), vl = rot(R, ld3(vi + vo + 3));
      const double th = sqrt(dot(om, om)) * dt;
      const double hs = th > 1e-12 ? sin(0.5 * th) / th * dt : 0.5 * dt, hc = cos(0.5 * th);
      const double ew = hc, ex = hs * om.x, ey = hs * om.y, ez = hs * om.z;
      const double w0 = qi[qo], x0 = qi[qo + 1], y0 = qi[qo + 2], z0 = qi[qo + 3];
      double w1 = w0 * ew - x0 * ex - y0 * ey - z0 * ez;
      double x1 = w0 * ex + x0 * ew + y0 * ez - z0 * ey;
      double y1 = w0 * ey - x0 * ez + y0 * ew + z0 * ex;
      double z1 = w0 * ez + x0 * ey - y0 * ex + z0 * ew;
      const double nrm = 1.0 / sqrt(w1 * w1 + x1 * x1 + y1 * y1 + z1 * z1);
      qi[qo] = w1 * nrm;
      qi[qo + 1] = x1 * nrm;
      qi[qo + 2] = y1 * nrm;
      qi[qo + 3] = z1 * nrm;
      qi[qo + 4] += dt * vl.x;
      qi[qo + 5] += dt * vl.y;
      qi[qo + 6] += dt * vl.z;
    }
  }
}

static int step_impl(qpc_controller* c, int64_t B, double* q, double* v, const qpc_batch_in* in, const qpc_batch_out* out,
                     double dt, int32_t nsteps, int32_t flags, void* stream_, const qpc_contact_model* plant, int32_t substeps);

int qpc_step_batch(qpc_controller* c, int64_t B, double* q, double* v, const qpc_batch_in* in, const qpc_batch_out* out,
                   double dt, int32_t nsteps, int32_t flags, void* stream_) {
  return step_impl(c, B, q, v, in, out, dt, nsteps, flags, stream_, nullptr, 1);
}

// Closed loop with a PLANT: every control tick (period dt, zero-order hold on tau: RigidBodySim's PeriodicController) is
// followed by `substeps` integration steps of dt / substeps of the forward dynamics under the soft ground contact.
int qpc_simulate_batch(qpc_controller* c, int64_t B, double* q, double* v, const qpc_batch_in* in, const qpc_batch_out* out,
                       const qpc_contact_model* plant, double dt, int32_t substeps, int32_t nticks, int32_t flags,
                       void* stream_) {
  if (!plant || substeps < 1 || !(plant->stiffness >= 0.0) || !(plant->damping >= 0.0) || !(plant->v_eps > 0.0))
    return qpc_fail(QPC_ERR_ARG, "qpc_simulate_batch: bad contact model / substeps");
  return step_impl(c, B, q, v, in, out, dt, nticks, flags, stream_, plant, substeps);
}

static int step_impl(qpc_controller* c, int64_t B, double* q, double* v, const qpc_batch_in* in, const qpc_batch_out* out,
                     double dt, int32_t nsteps, int32_t flags, void* stream_, const qpc_contact_model* plant, int32_t substeps) {
  if (!c || !c->finalized) return qpc_fail(QPC_ERR_STATE, "controller not finalized");
  if (!q || !v || B < 0 || nsteps < 0 || !(dt > 0.0)) return qpc_fail(QPC_ERR_ARG, "qpc_step_batch: bad arguments");
  if (B == 0 || nsteps == 0) return QPC_OK;
  const DevProgram& p = c->prog;
  const double* desired = in ? in->desired : nullptr;
  const double *cwt = in ? in->contact_weight : nullptr, *cmx = in ? in->contact_maxnormalforce : nullptr;
  const long long dstride = desired ? in->desired_stride : 0, cstride = cwt ? in->contact_stride : 0;
  if (desired && dstride != 0 && dstride < p.ndes)
    return qpc_fail(QPC_ERR_ARG, "desired_stride smaller than the number of desired values");
  if ((cwt == nullptr) != (cmx == nullptr))
    return qpc_fail(QPC_ERR_ARG, "contact_weight and contact_maxnormalforce must be given together");
  if (cwt && cstride != 0 && cstride < p.ncontacts)
    return qpc_fail(QPC_ERR_ARG, "contact_stride smaller than the number of contacts");
  std::lock_guard<std::mutex> lock(c->be.mu);
  CUDA_TRY(cudaSetDevice(c->be.device));
  if (c->be.dirty) {
    int rc = upload_program(c);
    if (rc) return rc;
  }
  int rc = check_tick_parameters(p, in);
  if (rc) return rc;
  rc = ensure_capacity(c, B, desired ? (dstride ? dstride : p.ndes) : 0, cwt ? (cstride ? cstride : p.ncontacts) : 0,
                       in && in->task_weight ? (in->task_weight_stride ? in->task_weight_stride : p.ntasks) : 0,
                       in && in->contact_geometry
                           ? (in->contact_geometry_stride ? in->contact_geometry_stride : 7 * p.ncontacts) : 0);
  if (rc) return rc;
  DeviceBuffers& b = c->be.buf;
  const bool host = flags != QPC_DEVICE_PTRS;
  cudaStream_t s = host ? c->be.stream : (cudaStream_t)stream_;
  BatchIO io;
  io.desired_stride = dstride;
  io.contact_stride = cstride;
  double *dq = q, *dv = v;
  io.desired = desired;
  io.cweight = cwt;
  io.cmaxnf = cmx;
  qpc_batch_out o;
  memset(&o, 0, sizeof(o));
  if (out) o = *out;
  if (host) {
    dq = b.q;
    dv = b.v;
    CUDA_TRY(cudaMemcpyAsync(dq, q, sizeof(double) * (size_t)B * p.nq, cudaMemcpyHostToDevice, s));
    CUDA_TRY(cudaMemcpyAsync(dv, v, sizeof(double) * (size_t)B * p.nv, cudaMemcpyHostToDevice, s));
    if (desired) {
      CUDA_TRY(cudaMemcpyAsync(b.desired, desired, sizeof(double) * (size_t)(dstride ? B * dstride : p.ndes),
                               cudaMemcpyHostToDevice, s));
      io.desired = b.desired;
    }
    if (cwt) {
      const size_t cnt = (size_t)(cstride ? B * cstride : p.ncontacts);
      CUDA_TRY(cudaMemcpyAsync(b.cw, cwt, sizeof(double) * cnt, cudaMemcpyHostToDevice, s));
      CUDA_TRY(cudaMemcpyAsync(b.cm, cmx, sizeof(double) * cnt, cudaMemcpyHostToDevice, s));
      io.cweight = b.cw;
      io.cmaxnf = b.cm;
    }
    o.tau = b.tau;
    o.vdot = b.vdot;
    o.wrench = b.wrench;
    o.status = b.status;
    o.iters = b.iters;
    o.residuals = b.res;
    o.factorizations = nullptr;
  }
  io.q = dq;
  io.v = dv;
  rc = stage_tick_parameters(c, B, in, host, true, io, s);
  if (rc) return rc;
  double* vdot = o.vdot ? o.vdot : b.vdot;
  int* status = o.status ? o.status : b.status;
  const DevProgram* dp = (const DevProgram*)c->be.d_prog;
  double* tau_dev = o.tau ? o.tau : b.tau;  // the plant needs the torques even when the caller does not ask for them
  int fdsm = 0;
  ContactModel cmodel{0, 0, 0, 1, 0};
  if (plant) {
    fdsm = (kin_smem_doubles(p.nb, p.nq, p.nv, p.ndes, p.ncontacts, p.N) + kin_fd_extra_doubles(p.nv)) * 8;
    if (fdsm > 227 * 1024) return qpc_fail(QPC_ERR_LIMIT, "mechanism too large for the forward-dynamics kernel");
    CUDA_TRY(raise_dyn_smem(qpc_forward_dynamics_kernel, fdsm, g_mark_fd));
    cmodel = ContactModel{plant->stiffness, plant->damping, plant->mu, plant->v_eps, plant->ground_z};
  }
  double* anchor = plant ? b.anchor : nullptr;  // persists across calls; qpc_reset_warm_start clears it
  for (int k = 0; k < nsteps; k++) {
    io.time_offset = k * dt;  // the functor's t (se3pdcontroller.jl:13): qpc_batch_in.time is the time of the first tick
    rc = run_tick(c, B, io, plant ? tau_dev : o.tau, vdot, o.wrench, status, o.iters, o.residuals, o.factorizations, s);
    if (rc) return rc;
    if (!plant) {
      qpc_integrate_kernel<<<(unsigned)((B + 127) / 128), 128, 0, s>>>(dp, dq, dv, vdot, status, dt, B);
      c->be.launches += 1;
    } else {
      // the commanded accelerations stay in `vdot` (what the caller reads); the plant's go through the x workspace's
      // neighbour b.vdot when the caller supplied its own buffer, else a scratch view of b.tau's twin
      double* vd_sim = b.ksave;  // [B][>= nv] scratch: the saved kinematic state is dead once the tick has finished
      for (int ss = 0; ss < substeps; ss++) {
        qpc_forward_dynamics_kernel<<<launch_grid(B), ASM_THREADS, fdsm, s>>>(dp, dq, dv, tau_dev, cmodel, vd_sim, anchor, B);
        qpc_integrate_kernel<<<(unsigned)((B + 127) / 128), 128, 0, s>>>(dp, dq, dv, vd_sim, status, dt / substeps, B);
        c->be.launches += 2;
      }
    }
  }
  CUDA_TRY(cudaGetLastError());
  if (host) {
    const size_t nb_ = (size_t)B;
    CUDA_TRY(cudaMemcpyAsync(q, dq, sizeof(double) * nb_ * p.nq, cudaMemcpyDeviceToHost, s));
    CUDA_TRY(cudaMemcpyAsync(v, dv, sizeof(double) * nb_ * p.nv, cudaMemcpyDeviceToHost, s));
    if (out) {
      const int nc6 = p.ncontacts * 6;
      if (out->tau) CUDA_TRY(cudaMemcpyAsync(out->tau, b.tau, sizeof(double) * nb_ * p.nv, cudaMemcpyDeviceToHost, s));
      if (out->vdot) CUDA_TRY(cudaMemcpyAsync(out->vdot, b.vdot, sizeof(double) * nb_ * p.nv, cudaMemcpyDeviceToHost, s));
      if (out->wrench) CUDA_TRY(cudaMemcpyAsync(out->wrench, b.wrench, sizeof(double) * nb_ * nc6, cudaMemcpyDeviceToHost, s));
      if (out->status) CUDA_TRY(cudaMemcpyAsync(out->status, b.status, sizeof(int) * nb_, cudaMemcpyDeviceToHost, s));
      if (out->iters) CUDA_TRY(cudaMemcpyAsync(out->iters, b.iters, sizeof(int) * nb_, cudaMemcpyDeviceToHost, s));
      if (out->residuals) CUDA_TRY(cudaMemcpyAsync(out->residuals, b.res, sizeof(double) * 2 * nb_, cudaMemcpyDeviceToHost, s));
      if (out->factorizations)
        CUDA_TRY(cudaMemcpyAsync(out->factorizations, c->be.d_nfac, sizeof(int) * nb_, cudaMemcpyDeviceToHost, s));
    }
    CUDA_TRY(cudaStreamSynchronize(s));
  }
  return QPC_OK;
}

int qpc_set_profiling(qpc_controller* c, int32_t on) {
  if (!c || !c->finalized) return qpc_fail(QPC_ERR_STATE, "controller not finalized");
  CUDA_TRY(cudaSetDevice(c->be.device));
  if (on && !c->be.ev[0])
    for (int i = 0; i < 4; i++) CUDA_TRY(cudaEventCreate(&c->be.ev[i]));
  c->be.profiling = on != 0;
  return QPC_OK;
}

int qpc_stage_times(qpc_controller* c, double ms[3]) {
  if (!c || !c->finalized || !c->be.ev[0]) return qpc_fail(QPC_ERR_STATE, "profiling was not enabled");
  CUDA_TRY(cudaSetDevice(c->be.device));
  CUDA_TRY(cudaEventSynchronize(c->be.ev[3]));
  for (int i = 0; i < 3; i++) {
    float t = 0;
    CUDA_TRY(cudaEventElapsedTime(&t, c->be.ev[i], c->be.ev[i + 1]));
    ms[i] = t;
  }
  return QPC_OK;
}

// ---- fp64 FMA peak of the device (the roofline denominator MEASURED_PEAKS.json does not carry) ----------------------
__global__ void qpc_dfma_peak_kernel(double* out, int iters) {
  double a0 = threadIdx.x * 1e-9, a1 = a0 + 1, a2 = a0 + 2, a3 = a0 + 3, a4 = a0 + 4, a5 = a0 + 5, a6 = a0 + 6, a7 = a0 + 7;
  const double b = 1.0000001, c = 1e-9;
  for (int i = 0; i < iters; i++) {
    a0 = fma(a0, b, c); a1 = fma(a1, b, c); a2 = fma(a2, b, c); a3 = fma(a3, b, c);
    a4 = fma(a4, b, c); a5 = fma(a5, b, c); a6 = fma(a6, b, c); a7 = fma(a7, b, c);
  }
  out[blockIdx.x * blockDim.x + threadIdx.x] = a0 + a1 + a2 + a3 + a4 + a5 + a6 + a7;
}

int qpc_measure_fp64_peak(int32_t device, double* tflops) {
  CUDA_TRY(cudaSetDevice(device));
  cudaDeviceProp prop;
  CUDA_TRY(cudaGetDeviceProperties(&prop, device));
  const int blocks = prop.multiProcessorCount * 8, threads = 256, iters = 1 << 16;
  double* d = nullptr;
  CUDA_TRY(cudaMalloc((void**)&d, sizeof(double) * blocks * threads));
  cudaEvent_t e0, e1;
  CUDA_TRY(cudaEventCreate(&e0));
  CUDA_TRY(cudaEventCreate(&e1));
  double best = 0;
  for (int rep = 0; rep < 5; rep++) {
    CUDA_TRY(cudaEventRecord(e0));
    qpc_dfma_peak_kernel<<<blocks, threads>>>(d, iters);
    CUDA_TRY(cudaEventRecord(e1));
    CUDA_TRY(cudaEventSynchronize(e1));
    float ms = 0;
    CUDA_TRY(cudaEventElapsedTime(&ms, e0, e1));
    const double tf = 2.0 * 8.0 * iters * (double)blocks * threads / (ms * 1e-3) / 1e12;
    if (rep > 0 && tf > best) best = tf;
  }
  cudaEventDestroy(e0);
  cudaEventDestroy(e1);
  cudaFree(d);
  *tflops = best;
  return QPC_OK;
}

int qpc_assemble_batch(qpc_controller* c, int64_t B, const qpc_batch_in* in, double* P, double* qv, double* G,
                       double* lg, double* ug, double* lb, double* ub, double* desired_out, int32_t flags,
                       void* stream_) {
  if (!c || !c->finalized) return qpc_fail(QPC_ERR_STATE, "controller not finalized");
  if (!in || !in->q || !in->v || B <= 0) return qpc_fail(QPC_ERR_ARG, "qpc_assemble_batch: bad arguments");
  const DevProgram& p = c->prog;
  std::lock_guard<std::mutex> lock(c->be.mu);
  CUDA_TRY(cudaSetDevice(c->be.device));
  if (c->be.dirty) {
    int rc = upload_program(c);
    if (rc) return rc;
  }
  const long long dstride = in->desired ? in->desired_stride : 0, cstride = in->contact_weight ? in->contact_stride : 0;
  int rc = check_tick_parameters(p, in);
  if (rc) return rc;
  rc = ensure_capacity(c, B, in->desired ? (dstride ? dstride : p.ndes) : 0,
                       in->contact_weight ? (cstride ? cstride : p.ncontacts) : 0,
                       in->task_weight ? (in->task_weight_stride ? in->task_weight_stride : p.ntasks) : 0,
                       in->contact_geometry
                           ? (in->contact_geometry_stride ? in->contact_geometry_stride : 7 * p.ncontacts) : 0);
  if (rc) return rc;
  DeviceBuffers& b = c->be.buf;
  const int ksm = kin_smem_doubles(p.nb, p.nq, p.nv, p.ndes, p.ncontacts, p.N) * 8;
  BatchIO io;
  io.desired_stride = dstride;
  io.contact_stride = cstride;
  const DevProgram* dp = (const DevProgram*)c->be.d_prog;
  if (flags == QPC_DEVICE_PTRS) {
    cudaStream_t s = (cudaStream_t)stream_;
    io.q = in->q;
    io.v = in->v;
    io.desired = in->desired;
    io.cweight = in->contact_weight;
    io.cmaxnf = in->contact_maxnormalforce;
    rc = stage_tick_parameters(c, B, in, false, false, io, s);
    if (rc) return rc;
    QpBuffers qb = qp_view(b);
    qb.prezeroed = 0;  // the caller's arrays
    qb.P = P;
    qb.qv = qv;
    qb.G = G;
    qb.lg = lg;
    qb.ug = ug;
    qb.lb = lb;
    qb.ub = ub;
    if (desired_out) qb.des = desired_out;
    if (p.nse3) qpc_assemble_kernel<true><<<launch_grid(B), ASM_THREADS, ksm, s>>>(dp, io, qb, 0, B);
    else qpc_assemble_kernel<false><<<launch_grid(B), ASM_THREADS, ksm, s>>>(dp, io, qb, 0, B);
    c->be.launches += 1;
    CUDA_TRY(cudaGetLastError());
    return QPC_OK;
  }
  cudaStream_t s = c->be.stream;
  CUDA_TRY(cudaMemcpyAsync(b.q, in->q, sizeof(double) * B * p.nq, cudaMemcpyHostToDevice, s));
  CUDA_TRY(cudaMemcpyAsync(b.v, in->v, sizeof(double) * B * p.nv, cudaMemcpyHostToDevice, s));
  io.q = b.q;
  io.v = b.v;
  io.desired = nullptr;
  io.cweight = io.cmaxnf = nullptr;
  if (in->desired) {
    const long long cnt = dstride ? B * dstride : p.ndes;
    CUDA_TRY(cudaMemcpyAsync(b.desired, in->desired, sizeof(double) * cnt, cudaMemcpyHostToDevice, s));
    io.desired = b.desired;
  }
  if (in->contact_weight) {
    const long long cnt = cstride ? B * cstride : p.ncontacts;
    CUDA_TRY(cudaMemcpyAsync(b.cw, in->contact_weight, sizeof(double) * cnt, cudaMemcpyHostToDevice, s));
    CUDA_TRY(cudaMemcpyAsync(b.cm, in->contact_maxnormalforce, sizeof(double) * cnt, cudaMemcpyHostToDevice, s));
    io.cweight = b.cw;
    io.cmaxnf = b.cm;
  }
  rc = stage_tick_parameters(c, B, in, true, true, io, s);
  if (rc) return rc;
  if (p.nse3) qpc_assemble_kernel<true><<<launch_grid(B), ASM_THREADS, ksm, s>>>(dp, io, qp_view(b), 0, B);
  else qpc_assemble_kernel<false><<<launch_grid(B), ASM_THREADS, ksm, s>>>(dp, io, qp_view(b), 0, B);
  c->be.launches += 1;
  CUDA_TRY(cudaGetLastError());
  const long long n = p.n, mg = p.mg, nbx = p.nbx;
  if (P) CUDA_TRY(cudaMemcpyAsync(P, b.P, sizeof(double) * B * n * n, cudaMemcpyDeviceToHost, s));
  if (qv) CUDA_TRY(cudaMemcpyAsync(qv, b.qv, sizeof(double) * B * n, cudaMemcpyDeviceToHost, s));
  if (G) CUDA_TRY(cudaMemcpyAsync(G, b.G, sizeof(double) * B * mg * n, cudaMemcpyDeviceToHost, s));
  if (lg) CUDA_TRY(cudaMemcpyAsync(lg, b.lg, sizeof(double) * B * mg, cudaMemcpyDeviceToHost, s));
  if (ug) CUDA_TRY(cudaMemcpyAsync(ug, b.ug, sizeof(double) * B * mg, cudaMemcpyDeviceToHost, s));
  if (lb) CUDA_TRY(cudaMemcpyAsync(lb, b.lb, sizeof(double) * B * nbx, cudaMemcpyDeviceToHost, s));
  if (ub) CUDA_TRY(cudaMemcpyAsync(ub, b.ub, sizeof(double) * B * nbx, cudaMemcpyDeviceToHost, s));
  if (desired_out)
    CUDA_TRY(cudaMemcpyAsync(desired_out, b.des, sizeof(double) * B * p.ndes, cudaMemcpyDeviceToHost, s));
  CUDA_TRY(cudaStreamSynchronize(s));
  return QPC_OK;
}

int qpc_solve_qp_batch(int32_t device, int64_t B, int32_t n, int32_t mg, int32_t nbox, const double* P,
                       const double* qv, const double* G, const double* lg, const double* ug, const double* lb,
                       const double* ub, const qpc_settings* st, double* x, double* y, int32_t* status,
                       int32_t* iters, double* residuals, int32_t flags, void* stream_) {
  if (B <= 0 || n <= 0 || mg < 0 || nbox < 0 || nbox > n || !P || !qv || !x || !status || !st)
    return qpc_fail(QPC_ERR_ARG, "qpc_solve_qp_batch: bad arguments");
  Settings s;
  qpc_copy_settings(st, s);
  const int asmem = admm_smem_doubles(n, mg, nbox) * 8;
  if (admm_vector_doubles(n, mg, nbox) * 8 > 227 * 1024)
    return qpc_fail(QPC_ERR_LIMIT, "QP vectors do not fit the 227 KB shared memory of one CTA");
  CUDA_TRY(cudaSetDevice(device));
  if (asmem <= ADMM_BIG_SMEM) CUDA_TRY(raise_dyn_smem(qpc_admm_kernel<ADMM_THREADS>, asmem, g_mark_admm128));
  QpBuffers qb;
  qb = QpBuffers();
  if (flags == QPC_DEVICE_PTRS) {
    qb.P = (double*)P;
    qb.qv = (double*)qv;
    qb.G = (double*)G;
    qb.lg = (double*)lg;
    qb.ug = (double*)ug;
    qb.lb = (double*)lb;
    qb.ub = (double*)ub;
    qb.x = x;
    qb.y = y;
    qb.status = status;
    qb.iters = iters;
    qb.res = residuals;
    CUDA_TRY(launch_admm(s, qb, n, mg, nbox, 0, B, (cudaStream_t)stream_));
    return QPC_OK;
  }
  // host pointers: temporary device copies (this entry point serves the synthetic-QP sweep, not the control tick)
  const long long m = mg + nbox;
  std::vector<void*> tmp;
  auto up = [&](const double* h, long long cnt, double*& d) -> cudaError_t {
    d = nullptr;
    cudaError_t e = cudaMalloc((void**)&d, sizeof(double) * (size_t)(cnt > 0 ? cnt : 1));
    if (e != cudaSuccess) return e;
    tmp.push_back(d);
    if (h && cnt > 0) e = cudaMemcpy(d, h, sizeof(double) * (size_t)cnt, cudaMemcpyHostToDevice);
    return e;
  };
  int rc = QPC_OK;
  int *dstat = nullptr, *diter = nullptr;
  do {
    cudaError_t e;
    if ((e = up(P, B * n * n, qb.P)) || (e = up(qv, B * n, qb.qv)) || (e = up(G, B * mg * n, qb.G)) ||
        (e = up(lg, B * mg, qb.lg)) || (e = up(ug, B * mg, qb.ug)) || (e = up(lb, B * nbox, qb.lb)) ||
        (e = up(ub, B * nbox, qb.ub)) || (e = up(nullptr, B * n, qb.x)) || (e = up(nullptr, B * m, qb.y)) ||
        (e = up(nullptr, B * 2, qb.res)) || (e = cudaMalloc((void**)&dstat, sizeof(int) * B)) ||
        (e = cudaMalloc((void**)&diter, sizeof(int) * B))) {
      rc = qpc_fail(QPC_ERR_CUDA, std::string(cudaGetErrorString(e)) + g_launch_note);
      break;
    }
    qb.status = dstat;
    qb.iters = diter;
    if ((e = launch_admm(s, qb, n, mg, nbox, 0, B, nullptr)) || (e = cudaDeviceSynchronize()) ||
        (e = cudaMemcpy(x, qb.x, sizeof(double) * B * n, cudaMemcpyDeviceToHost)) ||
        (y && (e = cudaMemcpy(y, qb.y, sizeof(double) * B * m, cudaMemcpyDeviceToHost))) ||
        (e = cudaMemcpy(status, dstat, sizeof(int) * B, cudaMemcpyDeviceToHost)) ||
        (iters && (e = cudaMemcpy(iters, diter, sizeof(int) * B, cudaMemcpyDeviceToHost))) ||
        (residuals && (e = cudaMemcpy(residuals, qb.res, sizeof(double) * 2 * B, cudaMemcpyDeviceToHost)))) {
      rc = qpc_fail(QPC_ERR_CUDA, std::string(cudaGetErrorString(e)) + g_launch_note);
      break;
    }
  } while (0);
  for (void* p : tmp) cudaFree(p);
  if (dstat) cudaFree(dstat);
  if (diter) cudaFree(diter);
  return rc;
}

}  // extern "C"
