// Development microbenchmark (not product): fp64 FMA issue rate on sm_100a for the operand patterns of the ADMM kernel.
//   peak   : a = fma(a, const, const), 8 independent chains (what qpc_measure_fp64_peak times)
//   tile   : acc[r] = fma(A[r][c], u[c], acc[r]), A 4 x TC register tile, 4 chains  (the matvec of one iteration)
//   tile8  : the same with 8 chains (two accumulator sets)
// Prints DFMA per clock per SM for several CTA sizes / CTAs per SM.
#include <cstdio>
#include <cuda_runtime.h>

template <int CH>
__global__ void k_peak(double* out, int iters) {
  double a[CH];
#pragma unroll
  for (int i = 0; i < CH; i++) a[i] = threadIdx.x * 1e-9 + i;
  const double b = 1.0000001, c = 1e-9;
  for (int it = 0; it < iters; it++) {
#pragma unroll
    for (int i = 0; i < CH; i++) a[i] = fma(a[i], b, c);
  }
  double s = 0;
#pragma unroll
  for (int i = 0; i < CH; i++) s += a[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <int TC, int SETS>
__global__ void __launch_bounds__(256) k_tile(const double* in, double* out, int iters) {
  double A[4][TC], u[TC], acc[SETS][4];
#pragma unroll
  for (int r = 0; r < 4; r++)
#pragma unroll
    for (int c = 0; c < TC; c++) A[r][c] = in[(r * TC + c) * 32 + (threadIdx.x & 31)];
#pragma unroll
  for (int c = 0; c < TC; c++) u[c] = in[2048 + c * 32 + (threadIdx.x & 31)];
#pragma unroll
  for (int s = 0; s < SETS; s++)
#pragma unroll
    for (int r = 0; r < 4; r++) acc[s][r] = 0.0;
#pragma unroll 1
  for (int it = 0; it < iters; it++) {
#pragma unroll
    for (int s = 0; s < SETS; s++)
#pragma unroll
      for (int c = 0; c < TC; c++)
#pragma unroll
        for (int r = 0; r < 4; r++) acc[s][r] = fma(A[r][c], u[c], acc[s][r]);
  }
  double t = 0;
#pragma unroll
  for (int s = 0; s < SETS; s++)
#pragma unroll
    for (int r = 0; r < 4; r++) t += acc[s][r];
  out[blockIdx.x * blockDim.x + threadIdx.x] = t;
}

template <typename F>
static double run(F launch, double dfma_per_thread, int blocks, int threads, int sms, double clock_ghz) {
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0);
  cudaEventCreate(&e1);
  float best = 1e30f;
  for (int rep = 0; rep < 4; rep++) {
    cudaEventRecord(e0);
    launch();
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    float ms;
    cudaEventElapsedTime(&ms, e0, e1);
    if (rep && ms < best) best = ms;
  }
  const double total = dfma_per_thread * blocks * (double)threads;
  return total / (best * 1e-3) / (clock_ghz * 1e9) / sms;  // DFMA per clock per SM at the nominal clock
}

int main() {
  cudaDeviceProp p;
  cudaGetDeviceProperties(&p, 0);
  const int sms = p.multiProcessorCount;
  const double ghz = p.clockRate * 1e-6;
  printf("device %s, %d SMs, %.3f GHz nominal\n", p.name, sms, ghz);
  double *in, *out;
  cudaMalloc(&in, 8 * 4096);
  cudaMemset(in, 0, 8 * 4096);
  cudaMalloc(&out, 8 * 1024 * 1024);
  const int iters = 20000;
  for (int threads : {128, 160, 256}) {
    for (int per_sm : {1, 2, 3, 4}) {
      const int blocks = sms * per_sm;
      if (threads * per_sm > 1024) continue;
      double r8 = run([&] { k_peak<8><<<blocks, threads>>>(out, iters); }, 8.0 * iters, blocks, threads, sms, ghz);
      double r4 = run([&] { k_peak<4><<<blocks, threads>>>(out, iters); }, 4.0 * iters, blocks, threads, sms, ghz);
      double t10 = run([&] { k_tile<10, 1><<<blocks, threads>>>(in, out, iters); }, 40.0 * iters, blocks, threads, sms, ghz);
      double t10b = run([&] { k_tile<10, 2><<<blocks, threads>>>(in, out, iters); }, 80.0 * iters, blocks, threads, sms, ghz);
      double t8 = run([&] { k_tile<8, 1><<<blocks, threads>>>(in, out, iters); }, 32.0 * iters, blocks, threads, sms, ghz);
      printf("threads %3d x %d CTA/SM (%2d warps/SM): peak8 %5.1f  peak4 %5.1f  tile4x10 %5.1f  tile4x10(8 chains) %5.1f  tile4x8 %5.1f  DFMA/clk/SM\n",
             threads, per_sm, threads * per_sm / 32, r8, r4, t10, t10b, t8);
    }
  }
  cudaError_t e = cudaDeviceSynchronize();
  printf("status %s\n", cudaGetErrorString(e));
  return 0;
}
