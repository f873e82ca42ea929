// tiny_thread.h -- launcher of the thread-per-instance tick (tiny_thread.cu): ONE kernel per tick for tiny mechanisms
// (the Acrobot of the PointAccelerationTask demo, BASELINE config 2).
#pragma once
#include <cuda_runtime.h>

#include "qpc_program.h"

namespace qpc {
#ifndef QPC_TINY_THREADS
#define QPC_TINY_THREADS 64
#endif
#ifndef QPC_TINY_MINBLOCKS
#define QPC_TINY_MINBLOCKS 6
#endif
constexpr int TINY_THREADS = QPC_TINY_THREADS;   // threads (= robot instances) per CTA
constexpr int TINY_MAX_BODIES = 4;
// doubles of per-thread local memory of the (small, large) size classes: kinematic state, QP + ADMM workspace
constexpr int TINY_KWS[2] = {320, 640};
constexpr int TINY_AWS[2] = {128, 448};
// size class (0, 1) of a program the thread-per-instance tick can run, or -1
int tiny_thread_class(const DevProgram& p);
cudaError_t tiny_thread_configure(int cls, const DevProgram& p);
// p: the host copy of the program (dimensions, settings); dp: its device copy
cudaError_t tiny_thread_tick(int cls, const DevProgram& p, const DevProgram* dp, const BatchIO& io, const QpBuffers& qb,
                             double* tau, double* vdot, double* wrench, long long lo, long long hi, cudaStream_t s);
}  // namespace qpc
