// kin_warp.h -- launchers of the warp-per-instance kinematics kernels (kin_warp.cu), used by api.cu's tick for tiny
// mechanisms.
#pragma once
#include <cuda_runtime.h>

#include "qpc_program.h"

namespace qpc {
#ifndef QPC_KIN_WPC
#define QPC_KIN_WPC 2
#endif
constexpr int KIN_WPC = QPC_KIN_WPC;        // instances (warps) per CTA
#ifndef QPC_KIN_WARP_MIN_CTAS
#define QPC_KIN_WARP_MIN_CTAS 10
#endif
constexpr int KIN_WARP_MIN_CTAS = QPC_KIN_WARP_MIN_CTAS;  // resident CTAs per SM the register budget is set for
constexpr int KIN_WARP_MAX_BODIES = 4;      // mechanisms up to this many bodies take the warp-per-instance kernels
constexpr int KIN_WARP_MAX_SMEM = 48 * 1024;  // ... if KIN_WPC instances fit the default dynamic shared memory
#ifndef QPC_KIN_ID_WPC
#define QPC_KIN_ID_WPC 2
#endif
constexpr int KIN_ID_WPC = QPC_KIN_ID_WPC;  // instances (warps) per CTA of the saved-state inverse-dynamics kernel
#ifndef QPC_KIN_ID_MIN_CTAS
#define QPC_KIN_ID_MIN_CTAS 8
#endif
constexpr int KIN_ID_MIN_CTAS = QPC_KIN_ID_MIN_CTAS;
cudaError_t kin_warp_id_saved_configure(int bytes_per_instance);
cudaError_t kin_warp_id_saved(const DevProgram* dp, const QpBuffers& qb, double* tau, double* vdot, double* wrench,
                              long long lo, long long hi, int bytes_per_instance, cudaStream_t s);
cudaError_t kin_warp_configure(int ksm_bytes);
cudaError_t kin_warp_assemble(const DevProgram* dp, const BatchIO& io, const QpBuffers& qb, long long lo, long long hi,
                              int ksm_bytes, cudaStream_t s, bool se3);
cudaError_t kin_warp_inverse_dynamics(const DevProgram* dp, const BatchIO& io, const QpBuffers& qb, double* tau,
                                      double* vdot, double* wrench, long long lo, long long hi, int ksm_bytes,
                                      cudaStream_t s);
}  // namespace qpc
