// tiny_thread.cu -- the whole control tick of a TINY mechanism in one kernel, one THREAD per robot instance.
//
// BASELINE config 2 (reference notebooks/PointAccelerationTask Demo.ipynb:135-202: the Acrobot, 2 bodies, a QP of 2
// variables and 3 rows) leaves a warp -- let alone a CTA -- idle: two or three bodies of parallelism in the sweeps, a
// 5 x 5 KKT system in the solver.  Here the very same kernel bodies (kin.cuh, admm.cuh) are compiled in a third
// execution model, QPC_THREAD_PER_INSTANCE: QPC_TID = 0, QPC_NT = 1, QPC_SYNC() = nothing, the "shared memory" of an
// instance is a per-thread local array.  Local memory is interleaved across the lanes of a warp by the hardware, and all
// instances run the same program, so lane i's access to workspace[k] coalesces with its neighbours'.
// state -> kinematics -> QP assembly -> ADMM -> inverse dynamics never leave the thread: one launch per tick, the QP
// never touches HBM (only q, v, desireds in; tau, vdot, wrenches, x, y, status out).
#define QPC_THREAD_PER_INSTANCE 1
#include <cuda_runtime.h>

#include "admm.cuh"
#include "kin.cuh"
#include "tiny_thread.h"

namespace qpc {

// CN / CMG / CNBX: the QP's dimensions at compile time (every solver loop unrolls), or -1 = read from the program
template <int KWS, int AWS, int CN = -1, int CMG = -1, int CNBX = -1>
__global__ void __launch_bounds__(TINY_THREADS, QPC_TINY_MINBLOCKS)
qpc_tiny_tick_kernel(const DevProgram* __restrict__ pg, Settings st, BatchIO io, QpBuffers qb, double* tau, double* vdot,
                     double* wrench, long long base, long long B) {
  const long long inst = base + (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (inst >= B) return;
  double kws[KWS], aws[AWS];
  KinSmem s = kin_layout(kws, pg->nb, pg->nq, pg->nv, pg->ndes, pg->ncontacts, pg->N);
  kin_load(pg, io, inst, s);
  kin_forward(pg, s);
  kin_composite(pg, s);
  kin_standing(pg, s);
  kin_se3pd(pg, io, inst, s);
  kin_contacts(pg, s);
  const int n = CN >= 0 ? CN : pg->n, mg = CMG >= 0 ? CMG : pg->mg, nbx = CNBX >= 0 ? CNBX : pg->nbx;
  double* P = aws;
  double* qv = P + n * n;
  double* G = qv + n;
  double* lg = G + mg * n;
  double* ug = lg + mg;
  double* lb = ug + mg;
  double* ub = lb + nbx;
  double* ws = ub + nbx;
  kin_assemble(pg, s, P, qv, G, lg, ug, lb, ub);
  for (int i = 0; i < pg->ndes; i++) qb.des[inst * pg->ndes + i] = s.des[i];
  if (n > 0) {
    AdmmProblem pb;
    pb.P = P;
    pb.qv = qv;
    pb.G = G;
    pb.lg = lg;
    pb.ug = ug;
    pb.lb = lb;
    pb.ub = ub;
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
    admm_solve<CN, CMG, CNBX>(st, pb, n, mg, nbx, ws);
  } else {
    qb.status[inst] = 1;
    if (qb.iters) qb.iters[inst] = 0;
    if (qb.res) qb.res[2 * inst] = qb.res[2 * inst + 1] = 0.0;
  }
  double* tdst = tau ? tau + inst * pg->nv : s.q;  // tau is always computed; discard into dead scratch if unwanted
  kin_inverse_dynamics(pg, s, qb.x + inst * n, vdot ? vdot + inst * pg->nv : nullptr,
                       wrench ? wrench + inst * pg->ncontacts * 6 : nullptr, tdst);
}

int tiny_thread_class(const DevProgram& p) {
  if (p.nb > TINY_MAX_BODIES) return -1;
  const int kneed = kin_smem_doubles(p.nb, p.nq, p.nv, p.ndes, p.ncontacts, p.N);
  const int aneed = p.n * p.n + p.n + p.mg * p.n + 2 * p.mg + 2 * p.nbx + admm_smem_doubles(p.n, p.mg, p.nbx);
  for (int c = 0; c < 2; c++)
    if (kneed <= TINY_KWS[c] && aneed <= TINY_AWS[c]) return c;
  return -1;
}

// kernel of a size class; the Acrobot demo's QP (2 variables, 3 equality rows, no box rows) has its own instantiation
typedef void (*TinyKernel)(const DevProgram*, Settings, BatchIO, QpBuffers, double*, double*, double*, long long, long long);
static TinyKernel tiny_kernel(int cls, const DevProgram& p) {
  if (cls == 0 && p.n == 2 && p.mg == 3 && p.nbx == 0) return qpc_tiny_tick_kernel<TINY_KWS[0], TINY_AWS[0], 2, 3, 0>;
  return cls == 0 ? qpc_tiny_tick_kernel<TINY_KWS[0], TINY_AWS[0]> : qpc_tiny_tick_kernel<TINY_KWS[1], TINY_AWS[1]>;
}

// The per-thread workspace is a local array: make sure the device's per-thread stack limit covers the kernel's frame
// (set once at finalize, outside steady state -- changing the limit reallocates the device's local-memory pool).
cudaError_t tiny_thread_configure(int cls, const DevProgram& p) {
  cudaFuncAttributes fa;
  cudaError_t e = cudaFuncGetAttributes(&fa, tiny_kernel(cls, p));
  if (e != cudaSuccess) return e;
  size_t cur = 0;
  e = cudaDeviceGetLimit(&cur, cudaLimitStackSize);
  if (e != cudaSuccess) return e;
  const size_t need = fa.localSizeBytes + 512;
  return cur >= need ? cudaSuccess : cudaDeviceSetLimit(cudaLimitStackSize, need);
}

cudaError_t tiny_thread_tick(int cls, const DevProgram& p, const DevProgram* dp, const BatchIO& io, const QpBuffers& qb,
                             double* tau, double* vdot, double* wrench, long long lo, long long hi, cudaStream_t s) {
  const long long g = (hi - lo + TINY_THREADS - 1) / TINY_THREADS;
  if (g <= 0) return cudaSuccess;
  tiny_kernel(cls, p)<<<(unsigned)g, TINY_THREADS, 0, s>>>(dp, p.settings, io, qb, tau, vdot, wrench, lo, hi);
  return cudaGetLastError();
}

}  // namespace qpc
