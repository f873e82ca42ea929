// kin_warp.cu -- the kinematics kernels of api.cu (assembly, inverse dynamics) in warp-per-instance form, for tiny
// mechanisms (the Acrobot of the PointAccelerationTask demo, BASELINE config 2: two bodies).
//
// A 64-thread CTA idles on two or three bodies; what bounds those kernels is the latency of one instance's serial
// phases times the instances resident per SM (8 CTAs at 128 registers per thread).  Here one WARP owns an instance:
// blockDim = (32, KIN_WPC), each warp has its own slice of the dynamic shared memory, and every QPC_SYNC() of kin.cuh
// compiles to __syncwarp() (QPC_WARP_PER_INSTANCE) -- twice the resident instances and no CTA barriers.  This is a
// separate translation unit because QPC_SYNC is a macro of the shared kinematics code; api.cu calls the launchers below.
#define QPC_WARP_PER_INSTANCE 1
#include <cuda_runtime.h>

#include "kin.cuh"
#include "kin_warp.h"

namespace qpc {

template <bool SE3>  // see qpc_assemble_kernel (api.cu)
__global__ void __launch_bounds__(32 * KIN_WPC, KIN_WARP_MIN_CTAS)
qpc_assemble_warp_kernel(const DevProgram* __restrict__ pg, BatchIO io, QpBuffers qb, long long base, long long B) {
  extern __shared__ double smem_all[];
  double* smem = smem_all + threadIdx.y * kin_smem_doubles(pg->nb, pg->nq, pg->nv, pg->ndes, pg->ncontacts, pg->N);
  for (long long inst = base + (long long)blockIdx.x * blockDim.y + threadIdx.y; inst < B;
       inst += (long long)gridDim.x * blockDim.y) {
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
    QPC_SYNC();
  }
}

// inverse dynamics from the kinematic state the assembly kernel saved (kin.cuh: kin_save / kin_id_load): no forward sweep,
// half the shared memory, one warp per instance
__global__ void __launch_bounds__(32 * KIN_ID_WPC, KIN_ID_MIN_CTAS)
qpc_id_saved_warp_kernel(const DevProgram* __restrict__ pg, QpBuffers qb, double* tau, double* vdot, double* wrench,
                         long long base, long long B) {
  extern __shared__ double smem_all[];
  const int per = kin_id_smem_doubles(pg->nb, pg->nv, pg->ndes, pg->ncontacts, pg->N);
  double* smem = smem_all + threadIdx.y * per;
  const int ks = kin_save_doubles(pg->nb, pg->nv, pg->ncontacts, pg->N);
  for (long long inst = base + (long long)blockIdx.x * blockDim.y + threadIdx.y; inst < B;
       inst += (long long)gridDim.x * blockDim.y) {
    KinSmem s = kin_id_layout(smem, pg->nb, pg->nv, pg->ndes, pg->ncontacts, pg->N);
    kin_id_load(pg, s, qb.ksave + inst * ks, qb.des + inst * pg->ndes);
    double* tdst = tau ? tau + inst * pg->nv : s.Jt;  // tau is always computed; discard into scratch if unwanted
    kin_inverse_dynamics(pg, s, qb.x + inst * pg->n, vdot ? vdot + inst * pg->nv : nullptr,
                         wrench ? wrench + inst * pg->ncontacts * 6 : nullptr, tdst);
  }
}

__global__ void __launch_bounds__(32 * KIN_WPC, KIN_WARP_MIN_CTAS)
qpc_inverse_dynamics_warp_kernel(const DevProgram* __restrict__ pg, BatchIO io, QpBuffers qb, double* tau, double* vdot,
                                 double* wrench, long long base, long long B) {
  extern __shared__ double smem_all[];
  double* smem = smem_all + threadIdx.y * kin_smem_doubles(pg->nb, pg->nq, pg->nv, pg->ndes, pg->ncontacts, pg->N);
  for (long long inst = base + (long long)blockIdx.x * blockDim.y + threadIdx.y; inst < B;
       inst += (long long)gridDim.x * blockDim.y) {
    KinSmem s = kin_layout(smem, pg->nb, pg->nq, pg->nv, pg->ndes, pg->ncontacts, pg->N);
    kin_load(pg, io, inst, s);
    for (int i = threadIdx.x; i < pg->ndes; i += blockDim.x) s.des[i] = qb.des[inst * pg->ndes + i];
    QPC_SYNC();
    kin_forward(pg, s);
    kin_contacts(pg, s);
    double* tdst = tau ? tau + inst * pg->nv : s.q;  // tau is always computed; discard into dead scratch if unwanted
    kin_inverse_dynamics(pg, s, qb.x + inst * pg->n, vdot ? vdot + inst * pg->nv : nullptr,
                         wrench ? wrench + inst * pg->ncontacts * 6 : nullptr, tdst);
  }
}

cudaError_t kin_warp_id_saved_configure(int bytes_per_instance) {
  return cudaFuncSetAttribute(qpc_id_saved_warp_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                              KIN_ID_WPC * bytes_per_instance);
}
cudaError_t kin_warp_id_saved(const DevProgram* dp, const QpBuffers& qb, double* tau, double* vdot, double* wrench,
                              long long lo, long long hi, int bytes_per_instance, cudaStream_t s) {
  const long long g = (hi - lo + KIN_ID_WPC - 1) / KIN_ID_WPC;
  qpc_id_saved_warp_kernel<<<(unsigned)(g < (1ll << 30) ? g : (1ll << 30)), dim3(32, KIN_ID_WPC),
                             KIN_ID_WPC * bytes_per_instance, s>>>(dp, qb, tau, vdot, wrench, lo, hi);
  return cudaGetLastError();
}

static unsigned warp_grid(long long count) {
  const long long g = (count + KIN_WPC - 1) / KIN_WPC;
  return (unsigned)(g < (1ll << 30) ? g : (1ll << 30));
}

cudaError_t kin_warp_configure(int ksm_bytes) {
  cudaError_t e = cudaFuncSetAttribute(qpc_assemble_warp_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                       KIN_WPC * ksm_bytes);
  if (e != cudaSuccess) return e;
  e = cudaFuncSetAttribute(qpc_assemble_warp_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, KIN_WPC * ksm_bytes);
  if (e != cudaSuccess) return e;
  return cudaFuncSetAttribute(qpc_inverse_dynamics_warp_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                              KIN_WPC * ksm_bytes);
}
cudaError_t kin_warp_assemble(const DevProgram* dp, const BatchIO& io, const QpBuffers& qb, long long lo, long long hi,
                              int ksm_bytes, cudaStream_t s, bool se3) {
  if (se3) qpc_assemble_warp_kernel<true><<<warp_grid(hi - lo), dim3(32, KIN_WPC), KIN_WPC * ksm_bytes, s>>>(dp, io, qb, lo, hi);
  else qpc_assemble_warp_kernel<false><<<warp_grid(hi - lo), dim3(32, KIN_WPC), KIN_WPC * ksm_bytes, s>>>(dp, io, qb, lo, hi);
  return cudaGetLastError();
}
cudaError_t kin_warp_inverse_dynamics(const DevProgram* dp, const BatchIO& io, const QpBuffers& qb, double* tau,
                                      double* vdot, double* wrench, long long lo, long long hi, int ksm_bytes,
                                      cudaStream_t s) {
  qpc_inverse_dynamics_warp_kernel<<<warp_grid(hi - lo), dim3(32, KIN_WPC), KIN_WPC * ksm_bytes, s>>>(dp, io, qb, tau, vdot,
                                                                                                   wrench, lo, hi);
  return cudaGetLastError();
}

}  // namespace qpc
