// admm_reg.cuh -- OSQP-style ADMM with the KKT inverse held in registers (sm_100a), one QP per CTA.
//
//   min 1/2 x'Px + q'x   s.t.  lg <= G x <= ug (mg general rows),  lb <= x[n-nbx..n) <= ub (nbx box rows)
//
// Device-side replacement for `MOI.optimize!(::OSQP.Optimizer)` reached by `solve!(qpmodel)` (reference
// src/lowlevel/momentum.jl:58; solver plugged in at momentum.jl:1,15-16,27).  Same algorithm and constants as admm.cuh
// (OSQP 0.5.x): Ruiz equilibration, per-row rho, relaxed ADMM, unscaled termination residuals every
// `check_termination` iterations, infeasibility certificates, adaptive rho with refactorisation.  What differs is the
// machine mapping:
//
//  * The quasi-definite KKT matrix  Z = [[-diag(1/rho), G], [G', P + sigma I + box terms]]  (NK = mg + n rows, padded
//    with an identity block to NP = 8 TC positions) lives in REGISTERS as 4 x TC tiles: thread t holds rows
//    4(t/8)..4(t/8)+3 and columns (t%8) TC .. (t%8) TC + TC-1.  Every vector element fetched from shared memory feeds
//    4 DFMAs, which keeps the 128 B/clk shared-memory pipe below the fp64 pipe (a 1 x C row layout is bound by it).
//    Box rows are a diagonal and are folded into the x-row that owns the variable.
//  * "Factorisation" = NP symmetric sweep steps turning the registers into -Z^-1 in place: per step the pivot row is
//    broadcast through shared memory and every thread does 4 TC DFMAs on its tile; one barrier per step.  The tile
//    columns rotate by one register per step so the pivot column is always register 0 and the new inverse column is
//    inserted at register TC-1 (no dynamic register index); after NP steps the rotation is the identity again.
//  * An ADMM iteration is ONE matrix-vector product  [nu; x~] = Z^-1 [z - y/rho; sigma x - q + box]  : 4 TC DFMAs per
//    thread against a double-buffered vector in shared memory, a 4-shuffle transpose-reduction over the 8 lanes of a
//    row group that leaves each lane with the entry of the row it owns, the projection / dual update in that lane's
//    registers, one barrier per iteration.
//  * Ruiz equilibration works on the same tiles: by symmetry of the KKT matrix the column norms are the row norms.
//  * The scaled, unswept matrix is kept in shared memory (thread-major, conflict-free) for the residual products of
//    the termination / adaptive-rho checks and for refactorisation after a rho update.
#pragma once
#include "admm.cuh"

namespace qpc {

#ifndef QPC_REG_PARK
#define QPC_REG_PARK 10
#endif
constexpr int REG_MAXW = 20;  // most warps per CTA of any tile (the reduction scratch holds 16 doubles per warp)
constexpr int REG_TR = 4;     // rows per thread

// NB = column blocks = lanes per row group (8 or 16); TC = tile columns per thread; NP = NB TC positions
QPC_HD int admm_reg_positions(int TC, int NB) { return NB * TC; }
QPC_HD int admm_reg_threads(int TC, int NB) { return (NB * TC / REG_TR) * NB; }  // (NP / 4 row groups) x NB blocks
QPC_HD int admm_reg_warps(int TC, int NB) { return (admm_reg_threads(TC, NB) + 31) / 32; }
QPC_HD int admm_reg_smem_doubles(int TC, int NB) {
  const int NP = NB * TC;
  const int NPV = NP + (TC % 4 == 0 ? 2 * NB : 0);  // vectors are padded by 2 per column block when TC % 4 == 0
  // K0 planes are strided by NT + 1 (see RegSolver::KS)
  return REG_TR * TC * (admm_reg_threads(TC, NB) + 1) + 2 * (2 * NPV + 2) + 2 * NPV + 3 * admm_reg_warps(TC, NB) * 16 + 13 * NP + 16;
}

#if defined(__CUDACC__)

// warp-wide maximum of NON-NEGATIVE doubles (NaN sorts above +inf and is therefore propagated): their IEEE bit
// patterns order like unsigned integers, so two 32-bit redux.sync operations do it
__device__ __forceinline__ double warp_max_nonneg(double v) {
  const unsigned hi = (unsigned)__double2hiint(v), lo = (unsigned)__double2loint(v);
  const unsigned mh = __reduce_max_sync(0xffffffffu, hi);
  const unsigned ml = __reduce_max_sync(0xffffffffu, hi == mh ? lo : 0u);
  return __hiloint2double((int)mh, (int)ml);
}
__device__ __forceinline__ double warp_sum(double a) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) a += __shfl_xor_sync(0xffffffffu, a, o);
  return a;
}

// block reduction of KM maxima of non-negative values followed by KS sums; every thread gets the (bitwise identical)
// result.  `red2` alternates between two buffers so no trailing barrier is needed.
template <int KM, int KS>
__device__ __forceinline__ void reg_block_reduce(double (&v)[KM + KS], double* red2, int& sel, int stride) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = blockDim.x >> 5;
  double* red = red2 + sel * stride;  // stride = 16 x warps of the CTA
  sel ^= 1;
#pragma unroll
  for (int k = 0; k < KM + KS; k++) {
    const double a = k < KM ? warp_max_nonneg(v[k]) : warp_sum(v[k]);
    if (lane == 0) red[warp * 16 + k] = a;
  }
  __syncthreads();
#pragma unroll
  for (int k = 0; k < KM + KS; k++) {
    const double t = lane < nw ? red[lane * 16 + k] : 0.0;
    v[k] = k < KM ? warp_max_nonneg(t) : warp_sum(t);
  }
}

// The same reduction split in two so that the 15 residual quantities never have to be live at once: stage 1 reduces
// one value over the warp and parks it, stage 2 (after a barrier) combines the warps' entries on demand.
template <bool IS_MAX>
__device__ __forceinline__ void red_put(double* red, int k, double v) {
  const double a = IS_MAX ? warp_max_nonneg(v) : warp_sum(v);
  if ((threadIdx.x & 31) == 0) red[(threadIdx.x >> 5) * 16 + k] = a;
}
template <bool IS_MAX>
__device__ __forceinline__ double red_get(const double* red, int k) {
  const int lane = threadIdx.x & 31, nw = blockDim.x >> 5;
  const double t = lane < nw ? red[lane * 16 + k] : 0.0;
  return IS_MAX ? warp_max_nonneg(t) : warp_sum(t);
}

// Transpose-reduction over the NB lanes (column blocks) of a row group: every lane enters with partial results for
// its 4 tile rows and leaves with the complete result of the row it owns, row (q / (NB/4)) & 3.  The first two
// exchanges halve the number of values carried (keep one half, send the other), the rest are plain all-reduces.
template <bool IS_MAX, int NB>
__device__ __forceinline__ double group_reduce(const double (&s)[REG_TR], int q) {
  auto comb = [](double a, double b) { return IS_MAX ? fmax(a, b) : a + b; };
  const bool hiA = q & (NB / 2), hiB = q & (NB / 4);
  const double k0 = hiA ? s[2] : s[0], k1 = hiA ? s[3] : s[1];
  const double o0 = hiA ? s[0] : s[2], o1 = hiA ? s[1] : s[3];
  const double r0 = comb(k0, __shfl_xor_sync(0xffffffffu, o0, NB / 2));
  const double r1 = comb(k1, __shfl_xor_sync(0xffffffffu, o1, NB / 2));
  const double k = hiB ? r1 : r0, o = hiB ? r0 : r1;
  double v = comb(k, __shfl_xor_sync(0xffffffffu, o, NB / 4));
#pragma unroll
  for (int d = NB / 8; d > 0; d >>= 1) v = comb(v, __shfl_xor_sync(0xffffffffu, v, d));
  return v;
}

// the same transpose-reduction for unsigned maxima (Ruiz norms on high words)
template <int NB>
__device__ __forceinline__ unsigned group_reduce_umax(const unsigned (&s)[REG_TR], int q) {
  const bool hiA = q & (NB / 2), hiB = q & (NB / 4);
  const unsigned k0 = hiA ? s[2] : s[0], k1 = hiA ? s[3] : s[1];
  const unsigned o0 = hiA ? s[0] : s[2], o1 = hiA ? s[1] : s[3];
  const unsigned r0 = max(k0, __shfl_xor_sync(0xffffffffu, o0, NB / 2));
  const unsigned r1 = max(k1, __shfl_xor_sync(0xffffffffu, o1, NB / 2));
  const unsigned k = hiB ? r1 : r0, o = hiB ? r0 : r1;
  unsigned v = max(k, __shfl_xor_sync(0xffffffffu, o, NB / 4));
#pragma unroll
  for (int d = NB / 8; d > 0; d >>= 1) v = max(v, __shfl_xor_sync(0xffffffffu, v, d));
  return v;
}

// TC doubles at `p` (8-byte aligned; 16-byte aligned when TC is even) -> registers
template <int TC>
__device__ __forceinline__ void load_vec(const double* __restrict__ p, double (&v)[TC]) {
  if constexpr (TC % 2 == 0) {
    const double2* p2 = reinterpret_cast<const double2*>(p);
#pragma unroll
    for (int c = 0; c < TC / 2; c++) {
      const double2 t = p2[c];
      v[2 * c] = t.x;
      v[2 * c + 1] = t.y;
    }
  } else {
#pragma unroll
    for (int c = 0; c < TC; c++) v[c] = p[c];
  }
}
template <int TC>
__device__ __forceinline__ void store_vec(double* __restrict__ p, const double (&v)[TC]) {
  if constexpr (TC % 2 == 0) {
    double2* p2 = reinterpret_cast<double2*>(p);
#pragma unroll
    for (int c = 0; c < TC / 2; c++) p2[c] = make_double2(v[2 * c], v[2 * c + 1]);
  } else {
#pragma unroll
    for (int c = 0; c < TC; c++) p[c] = v[c];
  }
}

template <int TC, int NB, bool ELIM = false>
struct RegSolver {
  static_assert(NB == 8 || NB == 16, "lanes per row group");
  static constexpr int NP = NB * TC;               // row / column positions: general rows (mg), x rows (n), padding
  static constexpr int NT = (NP / REG_TR) * NB;    // threads
  static constexpr int NW = (NT + 31) / 32;        // warps (reduction scratch: 16 doubles per warp and buffer)
  static constexpr int LPR = NB / REG_TR;          // lanes that own the same row (they hold identical row state)
  // K0 (the scaled, unswept matrix) is thread-major: plane (r, c) holds entry (r, c) of every thread's tile.  The plane
  // stride NT + 1 (odd) spreads the scattered 8-byte stores of the load phase (consecutive
  // matrix columns = consecutive planes) over the banks; with stride NT = 160 they all hit one bank.
  static constexpr int KS = NT + 1;
  static constexpr int TR = REG_TR;
  // Position-indexed vectors in shared memory (iteration vectors, published pivot rows, check / scaling vectors) are
  // read by the NB lanes of a row group at a stride of one column block.  When TC % 4 == 0 that stride is a multiple of
  // 32 bytes and the 16-byte loads collide (2- to 8-way bank conflicts), so every block is padded by two doubles:
  // position j lives at VP(j) = j + PAD (j / TC); the stride TC + 2 is conflict-free like the odd multiples of 16 bytes.
  static constexpr int PAD = (TC % 4 == 0) ? 2 : 0;
  static constexpr int NPV = NP + PAD * NB;   // padded vector length
  static constexpr int US = NPV + 2;          // stride of the double-buffered iteration vectors
  static constexpr int PS = 2 * NPV + 2;      // stride of the double-buffered published pivot rows (same storage)
  static __device__ __forceinline__ int VP(int j) { return PAD ? j + PAD * (j / TC) : j; }
  // ---- geometry -------------------------------------------------------------------------------------------------
  // ELIM: the leading `nel` variables of the caller's QP have a diagonal, positive cost block and are eliminated from
  // the KKT system (x_f = -P_ff^-1 (q_f + G_f' nu)): the (g, g) block becomes -(1/rho + M), M = G_f P_ff^-1 G_f', and
  // the iteration runs on NK = mg + n positions with n the number of KEPT variables.  Requires every general row to be
  // an equality (then the dual iterate y equals the relaxed average of nu, so x_f = -P_ff^-1 (q_f + G_f' y) is the
  // primal iterate OSQP would carry with sigma = 0 on those variables).  Ruiz scaling is OSQP's on the FULL problem
  // (G_f takes part in the row / column norms), residuals are the full problem's.  The floor this form puts on the
  // primal residual (~1e-7 on Atlas) restricts it to eps_abs >= 1e-6; the host falls back otherwise.
  int nel = 0, nfull = 0;            // eliminated variables, stride of the caller's P / G / q (= n + nel)
  double *GF, *PF, *QF, *DF, *UF;    // ELIM: G_f [mg x nel] unscaled, diag P_ff, q_f, column scalings D_f, 1 / diag P_ff
  int n, mg, nbx, NK, tid, q, g, row, h, c0;  // q: column block, g: row group, row: the row this lane owns
  int cvo, vrow, vg4;  // VP(c0), VP(row), VP(4 g): where this lane's column block / row / row group sit in a vector
  bool isx, isg, hasbox, hasc;  // x-row / general-constraint row / x-row that also owns a box row / owns any row
  // ---- shared memory -----------------------------------------------------------------------------------------------
  double *K0, *uv, *cv, *red, *SC;
  // SC: per-position constants, 13 arrays of NP: 0 q | 1 l | 2 u | 3 cb | 4 rho | 5 1/rho | 6 D | 7 E(box) | 8 diag |
  //     9 xprev | 10 yprev | 11 scalars | 12 cb rho
  int redsel;
  // ---- registers: the tile and the state of the owned row -----------------------------------------------------------
  double a[TR][TC];
  double x, z, yr;  // yr = y / rho of the owned constraint row: the iteration never needs y itself
  double wv;        // ELIM: relaxed average of nu of the owned general row (= the multiplier x_f is built from)
  double* WV;       // ELIM: wv of every general row, published for the residual products (padded layout)

  __device__ __forceinline__ double& sc(int k, int pos) const { return SC[k * NP + pos]; }

  __device__ __forceinline__ double rho_of(double rho0) const {
    const double l = sc(1, row), u = sc(2, row);
    if (l < -QPC_INFTY * QPC_MIN_SCALING && u > QPC_INFTY * QPC_MIN_SCALING) return QPC_RHO_MIN;
    if (u - l < QPC_RHO_TOL) return QPC_RHO_EQ_FACTOR * rho0;
    return rho0;
  }

  __device__ __forceinline__ void load_tile() {
#pragma unroll
    for (int r = 0; r < TR; r++)
#pragma unroll
      for (int c = 0; c < TC; c++) a[r][c] = K0[(r * TC + c) * KS + tid];
  }
  __device__ __forceinline__ void store_tile() {
#pragma unroll
    for (int r = 0; r < TR; r++)
#pragma unroll
      for (int c = 0; c < TC; c++) K0[(r * TC + c) * KS + tid] = a[r][c];
  }

  // a <- -(Z^-1) by NP symmetric sweep steps, pivot position s at step s (the -1/rho block first, then the by then
  // positive definite x block, then the identity padding).  See the header for the rotation.
  __device__ __forceinline__ void factor(double sigma) {
    if (h == 0) {
      const double cb = sc(3, row), rho = sc(4, row), rinv = sc(5, row);
      sc(8, row) = isx ? sigma + (hasbox ? rho * cb * cb : 0.0) : (isg ? -rinv : 1.0);
    }
    // the row state is dead weight during the sweep: park it in the (idle) check vector and xprev / yprev slots
    // the row state is per row (both lanes of a pair hold the same values)
    cv[vrow] = z;
    sc(9, row) = yr;
    sc(10, row) = x;
    load_tile();
    __syncthreads();
#pragma unroll
    for (int r = 0; r < TR; r++) {
      const int pos = 4 * g + r, ck = pos - c0;
      const double dd = sc(8, pos);
#pragma unroll
      for (int c = 0; c < TC; c++)
        if (c == ck) a[r][c] += dd;
    }
    // Published pivot row, double-buffered, PS doubles each: [0, NP) the raw register dump of the 8 column blocks (what
    // the DFMAs consume, in the current rotation), [NP, 2 NP) the same row in canonical column order (entry j is, by
    // symmetry, the multiplier of row j), [2 NP] 1 / pivot.
    double* pb = uv;
    auto publish = [&](double* pw, const double (&rowv)[TC], int sm1, bool pivot_lane) {
      store_vec<TC>(pw + cvo, rowv);
#pragma unroll
      for (int c = 0; c < TC; c++) {
        int idx = c + sm1;
        idx -= idx >= TC ? TC : 0;
        pw[NPV + cvo + idx] = rowv[c];
      }
      if (pivot_lane) pw[2 * NPV] = 1.0 / rowv[0];
    };
    if (g == 0) publish(pb, a[0], 0, q == 0);
    __syncthreads();
    for (int s4 = 0; s4 < NP / 4; s4++) {
      const bool own = g == s4;
#pragma unroll
      for (int rr = 0; rr < TR; rr++) {
        const int s = 4 * s4 + rr;
        if (s >= NK) {  // identity padding: the step is the bare rotation (multipliers are zero), no broadcast needed
#pragma unroll
          for (int r = 0; r < TR; r++) {
            const double t0 = a[r][0];
#pragma unroll
            for (int c = 0; c < TC - 1; c++) a[r][c] = a[r][c + 1];
            a[r][TC - 1] = t0;
          }
          continue;
        }
        const double* pr = pb + (s & 1) * PS;
        const int b = s / TC;  // pivot column block
        const double dinv = pr[2 * NPV];
        double f[TR];
        {
          const double2* f2 = reinterpret_cast<const double2*>(pr + NPV + vg4);
          const double2 fa = f2[0], fb = f2[1];
          f[0] = fa.x * dinv;
          f[1] = fa.y * dinv;
          f[2] = fb.x * dinv;
          f[3] = fb.y * dinv;
        }
        const bool inb = q == b;
        double t0[TR];
#pragma unroll
        for (int r = 0; r < TR; r++) t0[r] = a[r][0];
        double p[TC];
        load_vec<TC>(pr + cvo, p);
        // a[r][c-1] <- a[r][c] - f[r] p[c]  (the pivot row itself: a[r][c] / pivot)
        // (publishing the next pivot row before the other three rows are updated -- look-ahead -- was measured 5 % slower:
        // it costs the registers that keep the iteration loop spill-free)
#pragma unroll
        for (int c = 1; c < TC; c++) {
#pragma unroll
          for (int r = 0; r < TR; r++) {
            if (r == rr && own) a[r][c - 1] = a[r][c] * dinv;
            else a[r][c - 1] = fma(-f[r], p[c], a[r][c]);
          }
        }
#pragma unroll
        for (int r = 0; r < TR; r++) {
          if (r == rr && own) a[r][TC - 1] = inb ? -dinv : t0[r] * dinv;
          else a[r][TC - 1] = inb ? f[r] : fma(-f[r], p[0], t0[r]);
        }
        // publish the next pivot row (position s + 1: row group (s + 1) / 4, tile row (rr + 1) % 4)
        const int s1 = s + 1;
        constexpr int TRm = TR - 1;
        const int rn = (rr + 1) & TRm;
        if (s1 < NK && g == (s1 >> 2)) {
          const int b1 = s1 / TC;
          publish(pb + (s1 & 1) * PS, a[rn], s1 - b1 * TC, q == b1);
        }
        __syncthreads();
      }
    }
    z = cv[vrow];
    yr = sc(9, row);
    x = sc(10, row);
  }

  // right-hand side entry of the owned row for the next KKT solve: z - y/rho for a general row,
  // sigma x - q + cb (rho z - y) for an x row (box term only if it owns one)
  __device__ __forceinline__ double rhs_entry(double sigma) const {
    const double w = hasc ? z - yr : 0.0;
    if (isg) return w;
    if (isx) return fma(sigma, x, -sc(0, row)) + (hasbox ? sc(12, row) * w : 0.0);
    return 0.0;
  }
  __device__ __forceinline__ void set_rho(double rho0) {  // owner lanes: per-row rho and what depends on it
    const double r = rho_of(rho0);
    sc(4, row) = r;
    sc(5, row) = 1.0 / r;
    sc(12, row) = sc(3, row) * r;
  }

  // One ADMM iteration: KKT solve = tile times the right-hand side in `vec`, then the
  // relaxation / projection / dual update of the owned row, then its entry of the next right-hand side into `nxt`.
  // With yr = y/rho the dependent chain after the solve result t is  zt -> zr -> v -> clip -> rhs:
  //   v = zr + yr, z+ = clip(v), yr+ = v - z+, next rhs (general row) = z+ - yr+.
  __device__ __forceinline__ void iterate(const double* __restrict__ vec, double* __restrict__ nxt, double alpha,
                                          double oma, double sigma, int zf) {
    double s0[TR];
#pragma unroll
    for (int r = 0; r < TR; r++) s0[r] = 0.0;
    if constexpr (TC % 2 == 0 && TC >= 6) {
      // All 16-byte chunks of the vector must be in flight before the first DFMA: ptxas otherwise reuses one register
      // quad and serialises TC / 2 shared-memory round trips (25 % of the iteration's latency at one CTA per SM, ncu),
      // whatever the source order or volatility of the loads.  The first element is therefore made to depend on one
      // word of every later chunk through `zf`, a zero the compiler cannot prove (sign bit of max_iter): three ORs and
      // an AND on the integer pipe buy back four load latencies.
      double u[TC];
      load_vec<TC>(vec + cvo, u);
      unsigned hx = 0u;
#pragma unroll
      for (int k = 1; k < TC / 2; k++) hx |= (unsigned)__double2hiint(u[2 * k]);
      u[0] = __hiloint2double((int)((unsigned)__double2hiint(u[0]) | (hx & (unsigned)zf)), __double2loint(u[0]));
#pragma unroll
      for (int c = 0; c < TC; c++) {
#pragma unroll
        for (int r = 0; r < TR; r++) s0[r] = fma(a[r][c], u[c], s0[r]);
      }
    } else {
      double u[TC];
      load_vec<TC>(vec + cvo, u);
#pragma unroll
      for (int c = 0; c < TC; c++) {
#pragma unroll
        for (int r = 0; r < TR; r++) s0[r] = fma(a[r][c], u[c], s0[r]);
      }
    }
    // Everything that does not depend on the solve result t is fetched / computed before the shuffle reduction, and the
    // update is arranged so that the chain after t is short:  v = alpha zt + (1 - alpha) z + yr  with
    // zt = t / rho + (z - yr) (general row) or cb t (box row)  becomes one FMA  v = t k1 + c1;  then z+ = clip(v),
    // yr+ = v - z+, and the next right-hand side needs  z+ - yr+ = 2 z+ - v.
    const double rinv = sc(5, row), lo = sc(1, row), up = sc(2, row), cb = sc(3, row), qs = sc(0, row),
                 cbrho = sc(12, row);
    const double w = z - yr, oz = oma * z, ox = oma * x;
    double ow = 0.0;
    if constexpr (ELIM) ow = oma * wv;
    const double k1 = alpha * (isg ? rinv : cb);
    const double c1 = (isg ? fma(alpha, w, oz) : oz) + yr;
    asm volatile("" ::"d"(lo), "d"(up), "d"(qs), "d"(cbrho), "d"(k1), "d"(c1), "d"(ox));
    const double t = -group_reduce<false, NB>(s0, q);
    double rhs = 0.0, rw = 0.0;
    if (isx) {
      x = fma(alpha, t, ox);
      rhs = fma(sigma, x, -qs);
    }
    if constexpr (ELIM)
      if (isg) wv = fma(alpha, t, ow);  // t = nu on a general row
    if (hasc) {
      const double v = fma(t, k1, c1);
      double zn = v < lo ? lo : v;
      zn = zn > up ? up : zn;
      yr = v - zn;
      z = zn;
      rw = fma(2.0, zn, -v);
    }
    if (isg) rhs = rw;
    else if (hasbox) rhs = fma(cbrho, rw, rhs);
    if (h == 0) nxt[vrow] = rhs;
  }

  // product of the owned row of the scaled, unswept matrix with a vector in shared memory
  __device__ __forceinline__ double k0_product(const double* v) const {
    double s0[TR];
#pragma unroll
    for (int r = 0; r < TR; r++) s0[r] = 0.0;
#pragma unroll
    for (int c = 0; c < TC; c++) {
      const double e = v[cvo + c];
#pragma unroll
      for (int r = 0; r < TR; r++) s0[r] = fma(K0[(r * TC + c) * KS + tid], e, s0[r]);
    }
    return group_reduce<false, NB>(s0, q);
  }
  // x-rows get (P vx, G' vy), general rows get (G vx, 0)
  __device__ __forceinline__ void k0_products(const double* vx, const double* vy, double& px, double& py) const {
    px = k0_product(vx);
    py = k0_product(vy);
  }
  // ELIM: as k0_product(vy), but general rows multiply their (g, g) block (-M) with vw instead: x rows get G_k' vy,
  // general rows -M vw
  __device__ __forceinline__ double k0_product_yw(const double* vy, const double* vw) const {
    double s0[TR];
#pragma unroll
    for (int r = 0; r < TR; r++) s0[r] = 0.0;
#pragma unroll
    for (int c = 0; c < TC; c++) {
      const double ey = vy[cvo + c], ew = vw[cvo + c];
#pragma unroll
      for (int r = 0; r < TR; r++) s0[r] = fma(K0[(r * TC + c) * KS + tid], (4 * g + r < mg) ? ew : ey, s0[r]);
    }
    return group_reduce<false, NB>(s0, q);
  }

  // ELIM: the eliminated column fl = tid / 4 is served by the four lanes tid % 4 = 0..3 (requires 4 nel <= NT): entry
  // fl of the scaled G_f' v for two vectors of general-row values published at vy, vw (padded layout); E = sc(6, .) of
  // the general rows.  Every lane of the quad returns the complete sums; lanes with fl >= nel return 0.
  __device__ __forceinline__ void gf_transpose_times2(const double* vy, const double* vw, double& uy, double& uw) const {
    const int fl = tid >> 2, fp = tid & 3;
    double ay = 0.0, aw = 0.0;
    if (fl < nel) {
#pragma unroll 1
      for (int r = fp; r < mg; r += 4) {
        const double gs = GF[r * nel + fl] * sc(6, r);
        ay = fma(gs, vy[VP(r)], ay);
        aw = fma(gs, vw[VP(r)], aw);
      }
    }
    ay += __shfl_xor_sync(0xffffffffu, ay, 1);
    aw += __shfl_xor_sync(0xffffffffu, aw, 1);
    ay += __shfl_xor_sync(0xffffffffu, ay, 2);
    aw += __shfl_xor_sync(0xffffffffu, aw, 2);
    const double d = fl < nel ? DF[fl] : 0.0;
    uy = ay * d;
    uw = aw * d;
  }

  __device__ void solve(const Settings& st, const AdmmProblem& pb_, double* smem) {
    tid = threadIdx.x;
    q = tid % NB;
    g = tid / NB;
    c0 = q * TC;
    cvo = q * (TC + PAD);
    row = 4 * g + ((q / LPR) & 3);
    vrow = VP(row);
    vg4 = VP(4 * g);
    h = q % LPR;  // lane 0 of the LPR lanes that own a row does the writing
    NK = n + mg;
    isg = row < mg;
    isx = row >= mg && row < NK;
    const int xi = row - mg;
    hasbox = isx && xi >= n - nbx;
    hasc = hasbox || isg;
    K0 = smem;
    uv = K0 + TR * TC * KS;        // 2 x PS (sweep) overlaid by 2 x US (iterations)
    cv = uv + 2 * PS;              // 2 x NP
    red = cv + 2 * NPV;            // 3 x NW x 16 (two alternating buffers + the residual check's own)
    SC = red + 3 * NW * 16;        // 13 x NP
    redsel = 0;
    const int m = mg + nbx;
    // ---- load: coalesced global reads, scattered into the thread-major staging area ----------------------------------
    for (int k = tid; k < TR * TC * KS; k += NT) K0[k] = 0.0;
    for (int k = tid; k < 2 * PS + 2 * NPV; k += NT) uv[k] = 0.0;
    __syncthreads();
    auto k0_index = [&](int i, int j) {  // element (row position i, column position j)
      const int qq = j / TC;
      return ((i & 3) * TC + (j - qq * TC)) * KS + (i >> 2) * NB + qq;
    };
    if constexpr (!ELIM) {
      nel = 0;
      nfull = n;
    }
    for (int k = tid; k < n * n; k += NT) {
      const int i = k / n, j = k - i * n;
      K0[k0_index(mg + i, mg + j)] = pb_.P[(size_t)(nel + i) * nfull + nel + j];
    }
    for (int k = tid; k < mg * n; k += NT) {
      const int r = k / n, j = k - r * n;
      const double gv = pb_.G[(size_t)r * nfull + nel + j];
      K0[k0_index(r, mg + j)] = gv;
      K0[k0_index(mg + j, r)] = gv;
    }
    if constexpr (ELIM) {
      GF = SC + 13 * NP + 16;
      PF = GF + mg * nel;
      QF = PF + nel;
      DF = QF + nel;
      UF = DF + nel;
      WV = UF + nel;
      for (int k = tid; k < NPV; k += NT) WV[k] = 0.0;
      wv = 0.0;
      for (int k = tid; k < mg * nel; k += NT) {
        const int r = k / nel, f = k - r * nel;
        GF[k] = pb_.G[(size_t)r * nfull + f];
      }
      for (int f = tid; f < nel; f += NT) {
        PF[f] = pb_.P[(size_t)f * nfull + f];
        QF[f] = pb_.qv[f];
        DF[f] = 1.0;
        UF[f] = 1.0 / PF[f];
      }
    }
    x = 0.0;
    z = 0.0;
    yr = 0.0;
    double qs = 0.0, cb = 0.0, l = 0.0, u = 0.0;
    double D = 1.0, E = 1.0;  // accumulated Ruiz scalings of the owned row (E: its constraint row)
    if (isx) qs = pb_.qv[nel + xi];
    if (isg) {
      l = fmax(pb_.lg[row], -QPC_INFTY);
      u = fmin(pb_.ug[row], QPC_INFTY);
    } else if (hasbox) {
      const int b = xi - (n - nbx);
      l = fmax(pb_.lb[b], -QPC_INFTY);
      u = fmin(pb_.ub[b], QPC_INFTY);
      cb = 1.0;
    }
    __syncthreads();
    load_tile();
    if constexpr (ELIM) {  // -M (unscaled) into the (g, g) block; the offset G_f P_ff^-1 q_f of the general rows into l, u
      __syncthreads();  // K0 (the staging copy) has been read by everyone: its head is scratch for M until store_tile()
      for (int k = tid; k < mg * mg; k += NT) {
        const int i = k / mg, j = k - i * mg;
        double acc = 0.0;
#pragma unroll 1
        for (int f = 0; f < nel; f++) acc = fma(GF[i * nel + f] * UF[f], GF[j * nel + f], acc);
        K0[k] = acc;
      }
      __syncthreads();
#pragma unroll
      for (int r = 0; r < TR; r++) {
        const int i = 4 * g + r;
#pragma unroll
        for (int c = 0; c < TC; c++) {
          const int j = c0 + c;
          if (i < mg && j < mg) a[r][c] = -K0[i * mg + j];
        }
      }
      if (isg) {
        double cg = 0.0;
#pragma unroll 1
        for (int f = 0; f < nel; f++) cg = fma(GF[row * nel + f] * UF[f], QF[f], cg);
        l += cg;
        u += cg;
      }
    }
    // ---- Ruiz equilibration of [P A'; A 0]; A = [G; E_box] -----------------------------------------
    // Lazy form: the tile stays unscaled during the `scaling` passes; the accumulated scaling S of every position lives
    // in shared memory and the scaled magnitudes |a_ij| S_j are formed on the fly (one multiply per element and pass;
    // S_i is a common factor of the row).  Norms are maxima of the high words of those magnitudes (integer max,
    // relative truncation < 2^-20: any positive scaling is admissible, the one actually used is tracked exactly).
    // The cost scaling c applies to the P block only, so every row keeps two maxima: over the constraint columns and
    // over the x columns.  The x-column maximum of an x row is also the P-block column norm the cost scaling needs.
    double cscale = 1.0;
    double* sv = cv;  // accumulated scaling of every position, NP entries, written by the row owners
    double S = 1.0;   // accumulated scaling of the owned row (= D for x rows, E for general rows)
    if (h == 0) sv[vrow] = 1.0;
    __syncthreads();
    unsigned mgm = 0, mxm = 0;  // high words of max_j |a_ij| S_j over constraint columns / x columns, owned row
    double gfm = 0.0, cfm = 0.0;  // ELIM: max_f |G_f[row][f]| D_f (general rows) and max_r |G_f[r][fl]| S_r (fl < nel)
    const int fl = tid >> 2, fp = tid & 3;  // ELIM: eliminated column served by this lane, and its share of the rows
    auto scaled_maxima = [&]() {
      double dcol[TC];
      load_vec<TC>(sv + cvo, dcol);
      unsigned g4[TR], x4[TR];
#pragma unroll
      for (int r = 0; r < TR; r++) g4[r] = x4[r] = 0u;
#pragma unroll
      for (int c = 0; c < TC; c++) {
        const bool xc = c0 + c >= mg;
#pragma unroll
        for (int r = 0; r < TR; r++) {
          const unsigned hw = (unsigned)__double2hiint(a[r][c] * dcol[c]) & 0x7fffffffu;
          g4[r] = max(g4[r], xc ? 0u : hw);
          x4[r] = max(x4[r], xc ? hw : 0u);
        }
      }
      mgm = group_reduce_umax<NB>(g4, q);
      mxm = group_reduce_umax<NB>(x4, q);
      if constexpr (ELIM) {  // the eliminated columns' share: row maxima of G_f D_f (general rows), column maxima of S G_f
        gfm = 0.0;
        if (isg) {  // the LPR lanes that own the row share the columns
#pragma unroll 1
          for (int f = h; f < nel; f += LPR) gfm = fmax(gfm, fabs(GF[row * nel + f]) * DF[f]);
        }
#pragma unroll
        for (int d = 1; d < LPR; d <<= 1) gfm = fmax(gfm, __shfl_xor_sync(0xffffffffu, gfm, d));
        cfm = 0.0;
        if (fl < nel) {  // four lanes per eliminated column share the rows
#pragma unroll 1
          for (int r = fp; r < mg; r += 4) cfm = fmax(cfm, fabs(GF[r * nel + fl]) * sv[VP(r)]);
        }
        cfm = fmax(cfm, __shfl_xor_sync(0xffffffffu, cfm, 1));
        cfm = fmax(cfm, __shfl_xor_sync(0xffffffffu, cfm, 2));
      }
    };
    if (st.scaling > 0) scaled_maxima();
    __syncthreads();  // every warp has read sv before the first pass overwrites it
    for (int it = 0; it < st.scaling; it++) {
      const double ng = S * __hiloint2double((int)mgm, 0), nx = S * __hiloint2double((int)mxm, 0);
      double nr = isx ? fmax(ng, cscale * nx) : fmax(ng, nx);
      double dfn = 1.0;  // ELIM: this pass's scaling of eliminated column tid
      if constexpr (ELIM) {
        if (isg) nr = fmax(nx, S * gfm);  // the (g, g) block holds -M, which is not part of OSQP's KKT matrix
        if (fl < nel) {
          const double d0 = DF[fl];
          dfn = 1.0 / sqrt(limit_scaling(d0 * fmax(cfm, cscale * d0 * fabs(PF[fl]))));
        }
      }
      if (hasbox) nr = fmax(nr, fabs(cb));
      const double sr = row < NK ? 1.0 / sqrt(limit_scaling(nr)) : 1.0;
      const double eb = hasbox ? 1.0 / sqrt(limit_scaling(fabs(cb))) : 1.0;
      S *= sr;
      if (h == 0) sv[vrow] = S;
      if constexpr (ELIM)
        if (fl < nel && fp == 0) DF[fl] *= dfn;
      if (isx) {
        qs *= sr;
        D *= sr;
        if (hasbox) {
          cb *= eb * sr;
          E *= eb;
        }
      } else if (isg) {
        E *= sr;
      }
      __syncthreads();
      scaled_maxima();
      // cost scaling: mean column norm of P_bar (= row norms of the x block, by symmetry) vs |q_bar|_inf
      double v2[2];
      v2[0] = isx ? fabs(qs) : 0.0;
      v2[1] = (isx && h == 0) ? cscale * S * __hiloint2double((int)mxm, 0) : 0.0;
      if constexpr (ELIM)
        if (fl < nel && fp == 0) {  // eliminated columns: |q_bar_f| and the (diagonal) column norm of P_bar_ff
          const double d1 = DF[fl];
          v2[0] = fmax(v2[0], cscale * d1 * fabs(QF[fl]));
          v2[1] += cscale * d1 * d1 * fabs(PF[fl]);
        }
      reg_block_reduce<1, 1>(v2, red, redsel, NW * 16);
      double ct = limit_scaling(n + nel > 0 ? v2[1] / (n + nel) : 1.0);
      const double qn = limit_scaling(v2[0]);
      ct = 1.0 / fmax(ct, qn);
      if (isx) qs *= ct;
      cscale *= ct;
    }
    if (st.scaling > 0) {  // apply: a_ij <- S_i S_j a_ij, times c on the P block
      double dcol[TC], srow[TR];
      load_vec<TC>(sv + cvo, dcol);
#pragma unroll
      for (int r = 0; r < TR; r++) srow[r] = sv[vg4 + r];
#pragma unroll
      for (int c = 0; c < TC; c++) {
        const bool xc = c0 + c >= mg;
#pragma unroll
        for (int r = 0; r < TR; r++) {
          double f = (xc && 4 * g + r >= mg) ? cscale * srow[r] : srow[r];
          if constexpr (ELIM)
            if (!xc && 4 * g + r < mg) f = srow[r] / cscale;  // M_bar = E M E / c
          a[r][c] *= f * dcol[c];
        }
      }
    }
    l *= E;
    u *= E;
    store_tile();
    // Control block in shared memory (written by thread 0 at the end of every residual check, read by everyone after
    // the barrier): keeping these scalars out of the register file leaves the plain-iteration loop with nothing live
    // but the tile, the row state and a few pointers.
    double* CD = &sc(11, 0);                    // 0 cscale | 1 1/cscale | 2 rho | 3 pri_res | 4 dua_res
    int* CI = reinterpret_cast<int*>(CD + 8);   // 0 status | 1 iter | 2 nfac | 3 next check | 4 next adapt | 5 refactor | 6 done
    const int chk_iv = st.check_termination, ada_iv = (st.adaptive_rho && st.adaptive_rho_interval) ? st.adaptive_rho_interval : 0;
    // Warm start = OSQP's implicit warm start between the solves of one workspace: x, y and rho of
    // the previous tick of this batch slot; a slot whose last solve was not accepted (rho_io <= 0) starts cold.
    double rho_start = st.rho;
    bool warm = false;
    if (pb_.rho_io && pb_.x0 && pb_.y0) {
      const double r = *pb_.rho_io;
      warm = r > 0.0;
      if (warm) rho_start = r;
    }
    if (h == 0) {
      sc(0, row) = qs;
      sc(1, row) = l;
      sc(2, row) = u;
      sc(3, row) = cb;
      sc(6, row) = isx ? D : E;
      sc(7, row) = E;
      set_rho(rho_start);
    }
    if (tid == 0) {
      CD[0] = cscale;
      CD[1] = 1.0 / cscale;
      CD[2] = rho_start;
      CD[3] = CD[4] = 0.0;
      CI[0] = -10;
      CI[1] = 0;
      CI[2] = 1;
      CI[3] = chk_iv ? chk_iv : 0x7fffffff;
      CI[4] = ada_iv ? ada_iv : 0x7fffffff;
      CI[5] = 1;
      CI[6] = 0;
    }
    __syncthreads();
    if (warm) {  // x_bar = D^-1 x, y_bar = c E^-1 y, z = A x_bar (unprojected, as osqp_warm_start does)
      if (isx) x = pb_.x0[nel + xi] / D;
      double yw = 0.0;
      if (isg) yw = pb_.y0[row] * cscale / E;
      else if (hasbox) yw = pb_.y0[mg + xi - (n - nbx)] * cscale / E;
      yr = yw * sc(5, row);
      if constexpr (ELIM) wv = isg ? yw : 0.0;
      if (h == 0) cv[vrow] = isx ? x : 0.0;
      __syncthreads();
      double gx = k0_product(cv);
      if constexpr (ELIM)
        if (isg) {  // + E G_f x_f with the previous tick's x_f (plus the offset that l, u carry)
          double acc = 0.0;
#pragma unroll 1
          for (int f = 0; f < nel; f++) acc = fma(GF[row * nel + f], pb_.x0[f] + QF[f] * UF[f], acc);
          gx = fma(E, acc, gx);
        }
      z = isg ? gx : (hasbox ? cb * x : 0.0);
      __syncthreads();
    }
    // ---- iterations ------------------------------------------------------------------------------------------------------
    const double alpha = st.alpha, sigma = st.sigma, oma = 1.0 - st.alpha;
    const int zf = st.max_iter >> 31;  // 0 (max_iter > 0), opaque to the compiler: see iterate()
    constexpr int PARK = QPC_REG_PARK < TR * TC ? QPC_REG_PARK : TR * TC;
    volatile double park[PARK > 0 ? PARK : 1];
    for (;;) {
      if (CI[6]) break;
      int iter = CI[1];
      if (CI[5]) {
        factor(sigma);
        if (h == 0) uv[((iter + 1) & 1) * US + vrow] = rhs_entry(sigma);  // the buffer iteration iter+1 reads
        __syncthreads();
      }
      if (iter >= st.max_iter) break;
      // plain iterations up to the next one that needs residuals: a tight loop with nothing but the solve and the update
      const int next_chk = CI[3], next_ada = CI[4];
      const int next_special = min(min(next_chk, next_ada), st.max_iter);
      {
        const double* ub_ = uv + ((iter + 1) & 1) * US;
        double* un = uv + (iter & 1) * US;
#pragma unroll 1
        for (int k = next_special - iter - 1; k > 0; k--) {
          iterate(ub_, un, alpha, oma, sigma, zf);
          __syncthreads();
          const double* tmp = ub_;
          ub_ = un;
          un = const_cast<double*>(tmp);
        }
      }
      // the special iteration: same update, then residuals
      iter = next_special;
      const bool check = iter == next_chk, adapt = iter == next_ada;
      if (h == 0) {
        sc(9, row) = x;
        sc(10, row) = sc(4, row) * yr;
      }
      iterate(uv + (iter & 1) * US, uv + ((iter + 1) & 1) * US, alpha, oma, sigma, zf);
      // Manual live-range split: the residual check below never touches the tile but needs ~45 registers of its own;
      // left alone, ptxas spills tile entries for the whole solve and reloads them inside the plain-iteration loop
      // (8 LDL.64 per iteration).  Parking PARK entries in local memory across the check keeps the loop spill-free.
#pragma unroll
      for (int i = 0; i < PARK; i++) park[i] = a[i / TC][i % TC];
      const double y = sc(4, row) * yr;
      if (h == 0) {
        cv[vrow] = isx ? x : 0.0;
        cv[NPV + vrow] = isg ? y : 0.0;
        if constexpr (ELIM) WV[vrow] = isg ? wv : 0.0;
      }
      __syncthreads();
      // ---- residuals ------------------------------------------------------------------------------
      const double dx = x - sc(9, row), dy = y - sc(10, row);
      const double D_ = sc(6, row), E_ = hasbox ? sc(7, row) : D_;  // D: x rows; E: the owned constraint row
      const double qs_ = sc(0, row), cbv = sc(3, row), lo = sc(1, row), up = sc(2, row);
      const double cscale_ = CD[0], cinv = CD[1];
      double rho0 = CD[2];
      double px, py;
      if constexpr (ELIM) {
        px = k0_product(cv);
        py = k0_product_yw(cv + NPV, WV);
      } else {
        k0_products(cv, cv + NPV, px, py);
      }
      double* rbuf = red + 2 * (NW * 16);
      double pdy = 0.0;  // delta_y projected on the polar of the recession cone (primal infeasibility certificate)
      {
        double ax = hasc ? (isg ? px : cbv * x) : 0.0;
        if constexpr (ELIM)
          if (isg) ax = px + py;  // G_k x_k - M w = G x with x_f = -P_ff^-1 (q_f + G_f' w), offset folded into l, u
        const double einv = hasc ? 1.0 / E_ : 0.0;
        const double zz = hasc ? z : 0.0;
        const double r = ax - zz;
        red_put<true>(rbuf, 0, fabs(einv * r));
        red_put<true>(rbuf, 1, fabs(r));
        red_put<true>(rbuf, 2, fabs(einv * zz));
        red_put<true>(rbuf, 3, fabs(einv * ax));
        red_put<true>(rbuf, 4, fabs(zz));
        red_put<true>(rbuf, 5, fabs(ax));
        if (hasc) {
          pdy = dy;
          if (up > QPC_INFTY * QPC_MIN_SCALING) pdy = (lo < -QPC_INFTY * QPC_MIN_SCALING) ? 0.0 : fmin(pdy, 0.0);
          else if (lo < -QPC_INFTY * QPC_MIN_SCALING) pdy = fmax(pdy, 0.0);
        }
        red_put<true>(rbuf, 11, fabs(E_ * pdy));
        red_put<false>(rbuf, 13, (hasc && h == 0) ? up * fmax(pdy, 0.0) + lo * fmin(pdy, 0.0) : 0.0);
      }
      {
        const double aty = isx ? py + (hasbox ? cbv * y : 0.0) : 0.0;
        const double dinv = isx ? 1.0 / D_ : 0.0;
        const double pxx = isx ? px : 0.0, qq = isx ? qs_ : 0.0;
        const double r = pxx + qq + aty;
        double big = fmax(fabs(qq), fmax(fabs(aty), fabs(pxx))), dbig = dinv * big;
        double rf = 0.0, drf = 0.0;
        if constexpr (ELIM) {  // eliminated rows: P x_f + q_f = -G_f' w by construction, so the dual residual is G_f' (y - w)
          double uy, uw;
          gf_transpose_times2(cv + NPV, WV, uy, uw);
          if ((tid >> 2) < nel) {
            const double df = DF[tid >> 2], qf = cscale_ * df * QF[tid >> 2];
            const double bf = fmax(fabs(qf), fmax(fabs(uy), fabs(qf + uw)));
            big = fmax(big, bf);
            dbig = fmax(dbig, bf / df);
            rf = fabs(uy - uw);
            drf = rf / df;
          }
        }
        red_put<true>(rbuf, 6, fmax(fabs(dinv * r), drf));
        red_put<true>(rbuf, 7, fmax(fabs(r), rf));
        red_put<true>(rbuf, 8, dbig);
        red_put<true>(rbuf, 9, big);
        red_put<true>(rbuf, 10, (isx && !finite_val(x)) ? 1.0 : 0.0);
        red_put<true>(rbuf, 12, isx ? fabs(D_ * dx) : 0.0);
        red_put<false>(rbuf, 14, (isx && h == 0) ? qs_ * dx : 0.0);
      }
      __syncthreads();
      auto MX = [&](int k) { return red_get<true>(rbuf, k); };
      auto SM = [&](int k) { return red_get<false>(rbuf, k); };
      const double pri_res = MX(0), dua_res = cinv * MX(6);
      int status = -10;
      bool done = false, refactor = false;
      if (MX(10) != 0.0 || !finite_val(pri_res) || !finite_val(dua_res)) {
        status = -8;
        done = true;
      }
      if (!done && (check || iter == st.max_iter)) {
        for (int pass = 0; pass < 2 && !done; pass++) {
          if (pass == 1 && iter != st.max_iter) break;  // the 10x relaxed test only applies at the iteration limit
          const double f = pass ? 10.0 : 1.0;
          const double eps_abs = f * st.eps_abs, eps_rel = f * st.eps_rel;
          const double epi = f * st.eps_prim_inf, edi = f * st.eps_dual_inf;
          const bool prim_ok = m == 0 || pri_res < eps_abs + eps_rel * fmax(MX(2), MX(3));
          const bool dual_ok = dua_res < eps_abs + eps_rel * cinv * MX(8);
          if (prim_ok && dual_ok) {
            status = pass ? 2 : 1;
            done = true;
            break;
          }
          const double ndy = MX(11), ndx = MX(12);
          if (!prim_ok && ndy > epi && SM(13) < -epi * ndy) {
            // primal infeasibility: |D^-1 A' dy|_inf < eps |E dy|_inf
            if (h == 0) {
              cv[vrow] = 0.0;
              cv[NPV + vrow] = isg ? pdy : 0.0;
            }
            __syncthreads();
            double qx, qy;
            k0_products(cv, cv + NPV, qx, qy);
            double na[1];
            na[0] = isx ? fabs((qy + (hasbox ? cbv * pdy : 0.0)) / D_) : 0.0;
            if constexpr (ELIM) {
              double uy, uw;
              gf_transpose_times2(cv + NPV, cv + NPV, uy, uw);
              if ((tid >> 2) < nel) na[0] = fmax(na[0], fabs(uy) / DF[tid >> 2]);
            }
            reg_block_reduce<1, 0>(na, red, redsel, NW * 16);
            if (na[0] < epi * ndy) {
              status = pass ? 3 : -3;
              done = true;
              break;
            }
          }
          if (!dual_ok && ndx > edi && SM(14) < -cscale_ * edi * ndx) {
            // dual infeasibility: |D^-1 P dx|_inf small and A dx inside the recession cone of [l, u]
            if (h == 0) {
              cv[vrow] = isx ? dx : 0.0;
              cv[NPV + vrow] = 0.0;
            }
            __syncthreads();
            double qx, qy;
            k0_products(cv, cv + NPV, qx, qy);
            double nb2[2];
            nb2[0] = isx ? fabs(qx / D_) : 0.0;
            nb2[1] = 0.0;
            if (hasc) {
              const double adx = (isg ? qx : cbv * dx) / E_;
              if ((up < QPC_INFTY * QPC_MIN_SCALING && adx > edi * ndx) ||
                  (lo > -QPC_INFTY * QPC_MIN_SCALING && adx < -edi * ndx))
                nb2[1] = 1.0;
            }
            reg_block_reduce<2, 0>(nb2, red, redsel, NW * 16);
            if (nb2[0] < cscale_ * edi * ndx && nb2[1] == 0.0) {
              status = pass ? 4 : -4;
              done = true;
              break;
            }
          }
        }
        if (!done && iter == st.max_iter) {
          status = -2;
          done = true;
        }
      }
      if (!done && adapt) {
        // rho <- rho sqrt( (r_p / max(|Ax|,|z|)) / (r_d / max(|Px|,|A'y|,|q|)) ) on scaled quantities (B.3 step 6)
        const double pr = MX(1) / (fmax(MX(4), MX(5)) + 1e-10);
        const double dr = MX(7) / (MX(9) + 1e-10);
        double rho_new = rho0 * sqrt(pr / (dr + 1e-10));
        rho_new = fmin(fmax(rho_new, QPC_RHO_MIN), QPC_RHO_MAX);
        if (rho_new > rho0 * st.adaptive_rho_tolerance || rho_new < rho0 / st.adaptive_rho_tolerance) {
          rho0 = rho_new;
          __syncwarp();
          if (h == 0) set_rho(rho0);
          __syncwarp();
          yr = y * sc(5, row);  // y is unchanged by a rho update; yr = y / rho follows the new rho
          refactor = true;
        }
      }
      if (tid == 0) {
        CD[2] = rho0;
        CD[3] = pri_res;
        CD[4] = dua_res;
        CI[0] = status;
        CI[1] = iter;
        if (refactor) CI[2] += 1;
        if (check) CI[3] = next_chk + chk_iv;
        if (adapt) CI[4] = next_ada + ada_iv;
        CI[5] = refactor ? 1 : 0;
        CI[6] = done ? 1 : 0;
      }
#pragma unroll
      for (int i = 0; i < PARK; i++) a[i / TC][i % TC] = park[i];
      __syncthreads();
    }
    // ---- unscale and store -----------------------------------------------------------------------------------------------
    if constexpr (ELIM) {  // x_f = -P_ff^-1 (q_f + G_f' y) on unscaled quantities (y = E y_bar / c)
      if (h == 0) WV[vrow] = isg ? wv : 0.0;
      __syncthreads();
      double uy, uw;
      gf_transpose_times2(WV, WV, uy, uw);
      if ((tid >> 2) < nel && (tid & 3) == 0) {
        const int f = tid >> 2;
        pb_.x[f] = -(QF[f] + uw * CD[1] / DF[f]) / PF[f];
      }
    }
    if (h == 0) {
      const double cinv = CD[1];
      if (isx) pb_.x[nel + xi] = sc(6, row) * x;
      if (pb_.y) {
        const double y = sc(4, row) * yr;
        if (isg) pb_.y[row] = cinv * sc(6, row) * y;
        if (hasbox) pb_.y[mg + xi - (n - nbx)] = cinv * sc(7, row) * y;
      }
    }
    if (tid == 0) {
      *pb_.status = CI[0];
      if (pb_.iters) *pb_.iters = CI[1];
      if (pb_.nfac) *pb_.nfac = CI[2];
      if (pb_.rho_io) *pb_.rho_io = (CI[0] == 1 || CI[0] == 2) ? CD[2] : -1.0;
      if (pb_.res) {
        pb_.res[0] = CD[3];
        pb_.res[1] = CD[4];
      }
    }
    __syncthreads();
  }
};

#endif  // __CUDACC__

}  // namespace qpc
