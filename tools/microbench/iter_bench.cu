// Development microbenchmark (not product): skeleton of one ADMM iteration of admm_reg.cuh (4 x TC register tile times a
// shared-memory vector, transpose-reduction over 8 lanes, row update, one barrier) to study how co-resident CTAs scale.
// Variants switch off one ingredient at a time.  Prints cycles per iteration per CTA and per SM.
#include <cstdio>
#include <cuda_runtime.h>

template <int TC, int VAR>  // VAR bits: 1 = no shuffles, 2 = no barrier, 4 = no vector loads (registers), 8 = no tail
__global__ void __launch_bounds__(8 * TC * 2) k_iter(const double* in, double* out, int iters, long long* cyc) {
  constexpr int NP = 8 * TC, NT = NP / 4 * 8;
  extern __shared__ double sm[];
  double* vec0 = sm;
  double* vec1 = sm + NP + 2;
  double* cst = sm + 2 * (NP + 2);
  const int tid = threadIdx.x, q = tid & 7, g = tid >> 3, c0 = q * TC, row = 4 * g + ((q >> 1) & 3);
  double a[4][TC];
#pragma unroll
  for (int r = 0; r < 4; r++)
#pragma unroll
    for (int c = 0; c < TC; c++) a[r][c] = in[(r * TC + c) * 32 + (tid & 31)] * 1e-3;
  for (int i = tid; i < 2 * (NP + 2) + 6 * NP; i += NT) sm[i] = 1e-3 * i;
  __syncthreads();
  double x = 0.1 * tid, z = 0.2, yr = 0.3;
  const double* ub = vec0;
  double* un = vec1;
  const long long t0 = clock64();
#pragma unroll 1
  for (int it = 0; it < iters; it++) {
    double s0[4] = {0, 0, 0, 0};
    if (VAR & 4) {
#pragma unroll
      for (int c = 0; c < TC; c++)
#pragma unroll
        for (int r = 0; r < 4; r++) s0[r] = fma(a[r][c], x + c, s0[r]);
    } else {
      const double2* v2 = reinterpret_cast<const double2*>(ub + c0);
#pragma unroll
      for (int k = 0; k < TC / 2; k++) {
        const double2 u = v2[k];
#pragma unroll
        for (int r = 0; r < 4; r++) s0[r] = fma(a[r][2 * k], u.x, s0[r]);
#pragma unroll
        for (int r = 0; r < 4; r++) s0[r] = fma(a[r][2 * k + 1], u.y, s0[r]);
      }
    }
    double t;
    if (VAR & 1) {
      t = s0[0] + s0[1] + s0[2] + s0[3];
    } else {
      const bool hiA = q & 4, hiB = q & 2;
      const double k0 = hiA ? s0[2] : s0[0], k1 = hiA ? s0[3] : s0[1];
      const double o0 = hiA ? s0[0] : s0[2], o1 = hiA ? s0[1] : s0[3];
      const double r0 = k0 + __shfl_xor_sync(0xffffffffu, o0, 4);
      const double r1 = k1 + __shfl_xor_sync(0xffffffffu, o1, 4);
      const double k = hiB ? r1 : r0, o = hiB ? r0 : r1;
      t = k + __shfl_xor_sync(0xffffffffu, o, 2);
      t += __shfl_xor_sync(0xffffffffu, t, 1);
    }
    double rhs;
    if (VAR & 8) {
      rhs = t * 1e-3;
    } else {
      const double rinv = cst[row], lo = cst[NP + row], up = cst[2 * NP + row], qs = cst[3 * NP + row];
      x = fma(1.6, -t, -0.6 * x);
      const double zt = fma(-t, rinv, z - yr);
      const double zr = fma(1.6, zt, -0.6 * z);
      const double v = zr + yr;
      double zn = v < lo ? lo : v;
      zn = zn > up ? up : zn;
      yr = v - zn;
      z = zn;
      rhs = (row & 1) ? zn - yr : fma(1e-6, x, -qs);
    }
    if ((q & 1) == 0) un[row] = rhs;
    if (!(VAR & 2)) __syncthreads();
    const double* tmp = ub;
    ub = un;
    un = const_cast<double*>(tmp);
  }
  const long long t1 = clock64();
  out[blockIdx.x * NT + tid] = x + z + yr;
  if (tid == 0) cyc[blockIdx.x] = t1 - t0;
}

template <int TC, int VAR>
static void run(const char* name, const double* in, double* out, long long* cyc, int sms) {
  constexpr int NP = 8 * TC, NT = NP / 4 * 8;
  const int iters = 4000;
  const size_t base = (2 * (NP + 2) + 6 * NP) * 8;
  printf("%-28s TC=%2d threads=%3d:", name, TC, NT);
  for (int per_sm = 1; per_sm <= 4; per_sm++) {
    // pad dynamic shared memory so that exactly per_sm CTAs fit on an SM
    size_t smem = (227 * 1024) / per_sm - 2048;
    if (smem < base) smem = base;
    cudaFuncSetAttribute(k_iter<TC, VAR>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    const int blocks = sms * per_sm;
    k_iter<TC, VAR><<<blocks, NT, smem>>>(in, out, iters, cyc);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf(" [%s]", cudaGetErrorString(e)); continue; }
    static long long h[4096];
    cudaMemcpy(h, cyc, sizeof(long long) * blocks, cudaMemcpyDeviceToHost);
    double s = 0;
    for (int i = 0; i < blocks; i++) s += (double)h[i];
    const double per_cta = s / blocks / iters;
    printf("  %dx: %6.0f/CTA %6.0f/SM", per_sm, per_cta, per_cta / per_sm);
  }
  printf("  cycles per iteration\n");
}

int main() {
  cudaDeviceProp p;
  cudaGetDeviceProperties(&p, 0);
  const int sms = p.multiProcessorCount;
  double *in, *out;
  long long* cyc;
  cudaMalloc(&in, 8 * 8192);
  cudaMemset(in, 0, 8 * 8192);
  cudaMalloc(&out, 8 * 1024 * 1024);
  cudaMalloc(&cyc, 8 * 4096);
  run<10, 0>("full iteration", in, out, cyc, sms);
  run<10, 1>("no shuffles", in, out, cyc, sms);
  run<10, 2>("no barrier", in, out, cyc, sms);
  run<10, 4>("no vector loads", in, out, cyc, sms);
  run<10, 8>("no row update", in, out, cyc, sms);
  run<10, 15>("DFMA only", in, out, cyc, sms);
  run<8, 0>("full iteration", in, out, cyc, sms);
  run<8, 2>("no barrier", in, out, cyc, sms);
  run<8, 15>("DFMA only", in, out, cyc, sms);
  run<12, 0>("full iteration", in, out, cyc, sms);
  return 0;
}
