// admm_warp.cuh -- OSQP-style ADMM, ONE WARP PER QP, the 32 x 32 iteration operator in registers (sm_100a).
//
//   min 1/2 x'Px + q'x   s.t.  G x = b (MG equality rows),  lb <= x_b <= ub,   x = (x_a [NA], x_b [nbx <= 32])
//
// Device-side replacement for `MOI.optimize!(::OSQP.Optimizer)` reached by `solve!(qpmodel)` (reference
// src/lowlevel/momentum.jl:58) for controller programs whose general rows are all equalities and whose unboxed variables
// x_a (free accelerations, task-error slacks: momentum.jl:28,119-126) are determined by those rows (NA <= MG, G_a of
// full column rank) -- the StandingController's program (standing.jl:31-50) is the model case: NA = 21, MG = 24
// (api.cu also instantiates NA = 27, MG = 30: that program plus one weighted 6-row task).
//
// Instead of iterating on the (n + m)-dimensional KKT system, the warp first REDUCES the problem (once per solve):
//   1. Householder QR of G_a applied to [G_a | G_b | b], one matrix column per lane, reflectors broadcast through
//      shared memory: x_a = xa0 - W x_b and the ME = MG - NA remaining rows A3 x_b = b3 (orthonormalised).
//   2. H = P_bb + W'P_aa W, h = q_b - W'(P_aa xa0 + q_a): a QP in the nbx friction-cone multipliers alone.
// and then runs ADMM with the box as the only split constraint; the rows A3 are handled exactly inside the x-update
// and sigma = 0 (K = H + diag(rho) is positive definite):
//      x~ = T (rho .* z - y) + t0,    T = K^-1 - K^-1 A3'(A3 K^-1 A3')^-1 A3 K^-1
//   3. "Factorisation" = in-place Gauss-Jordan inversion of K, row j in the registers of lane j (32 pivots, the pivot
//      row broadcast through shared memory, one __syncwarp per pivot), then the rank-ME correction.
//   4. An iteration = lane j's row of T (32 registers) times the vector u = z - y/rho read from shared memory as 16
//      broadcast 16-byte loads: 32 DFMA per lane, NO cross-lane reduction, one __syncwarp, then relaxation / projection /
//      dual update in the lane's registers.  The residuals of OSQP's termination test are lane-local quantities
//      (primal: x~ - z; dual: the stationarity defect of the x-update, y+ - y - rho (x~ - z)) + warp maxima by redux.sync,
//      tested every 25 iterations (leaving the tight loop more often costs more than the iterations it saves).
// Per-row rho (OSQP's rho_vec: equality rows get 1e3 rho) is extended to the active set: rows whose z sits on a bound
// get kappa rho, interior rows rho / kappa, re-evaluated together with OSQP's residual-balancing rule on a geometric
// schedule (iterations first, first growth, ...), so the number of refactorisations is bounded and the iteration is a
// fixed-rho ADMM from then on.  Measured on the 16,384-state Atlas workload (tools/warp_proto.py): 60-80 iterations
// instead of 218, none above 300 instead of a tail to 5,000.
//
// Equality rows hold to rounding for every iterate (they are eliminated), and the eliminated variables' stationarity
// rows hold exactly by construction, so OSQP's residuals on the full problem are the box rows' (primal) and the x_b
// rows' (dual); the eps_rel normalisers max(|Ax|, |z|) and max(|Px|, |A'y|, |q|) are evaluated on the full problem.
// Anything the reduction cannot handle (rank-deficient G_a or A3, a non-positive pivot, infinite bounds, non-finite
// numbers) is flagged with status QPC_WARP_FALLBACK and appended to a list the register-tile kernel then solves.
#pragma once
#include "admm.cuh"

namespace qpc {

constexpr int QPC_WARP_FALLBACK = -99;

struct WarpParams {
  double kappa;   // rho of rows on a bound = kappa rho, interior rows rho / kappa (1 = OSQP's uniform rho)
  double growth;  // adaptation at first, then every max(first, (growth - 1) x current) iterations
  int first;      // first adaptation iteration
  int check;      // residual check interval
  int aitken;     // extrapolation of slowly converging solves every this many iterations (0 = off)
  double gather_eps;  // the inversion gathers its pivot rows by symmetry when eps_abs >= this or eps_rel >= this / 100
                      // (measured fine down to eps_abs = eps_rel = 1e-8; not at eps_abs 1e-8 with eps_rel 1e-16: WarpSolver::factor)
};

#if defined(__CUDACC__)
#define QPC_NOINLINE __noinline__
#else
#define QPC_NOINLINE
#endif
#ifndef QPC_WARP_CPASYNC
#define QPC_WARP_CPASYNC 1  // prologue: cp.async copies of G into shared memory (0: register-staged loop)
#endif
#ifndef QPC_WARP_PBATCH
#define QPC_WARP_PBATCH 4   // rows of P_bb in flight before the first use
#endif

#if defined(__CUDACC__) || defined(QPC_WARP_EMU)  // QPC_WARP_EMU: tests/emu/warp_emu.cpp runs the body on CPU fibres

// warp-wide maximum of NON-NEGATIVE doubles (NaN sorts above +inf and is therefore propagated): their IEEE bit patterns
// order like unsigned integers, so two 32-bit redux.sync operations do it
__device__ __forceinline__ double warp_max_nonneg_w(double v) {
  const unsigned hi = (unsigned)__double2hiint(v), lo = (unsigned)__double2loint(v);
  const unsigned mh = __reduce_max_sync(0xffffffffu, hi);
  const unsigned ml = __reduce_max_sync(0xffffffffu, hi == mh ? lo : 0u);
  return __hiloint2double((int)mh, (int)ml);
}

// D = A B + C on the fp64 tensor cores: mma.sync m8n8k4 (A 8x4 row-major: lane holds A[lane/4][lane%4]; B 4x8 column-major:
// lane holds B[lane%4][lane/4]; C/D 8x8: lane holds [lane/4][2 (lane%4) + {0,1}]).  SASS: DMMA.8x8x4.
#if defined(__CUDA_ARCH__)
__device__ __forceinline__ void dmma884(double& d0, double& d1, double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(d0), "+d"(d1) : "d"(a), "d"(b));
}
#endif
#ifndef QPC_WARP_DMMA
#define QPC_WARP_DMMA 1  // reduced Hessian W'(P_aa W) by DMMA (0: DFMA on rotating row registers)
#endif

template <int MG, int NA>
struct WarpSolver {
  static_assert(NA < 32 && NA <= MG && MG <= 32, "x_a columns plus the right-hand side must fit one column per lane");
  static constexpr int ME = MG - NA;
  static constexpr int MEP = ME > 0 ? ME : 1;
  static constexpr int QR_NAP = NA + 1 + ((NA + 1) & 1);  // row stride of the QR scratch (even, >= NA + 1)
  // the H area also holds the QR scratch before H exists: two [NA][QR_NAP] blocks (NA > 22: larger than H itself)
  static constexpr int HAREA = 2 * NA * QR_NAP > 1024 ? 2 * NA * QR_NAP : 1024;
  static constexpr int OFF_H = 0;                     // H, element (i, lane) at i * 32 + lane; setup scratch before
  static constexpr int OFF_W = OFF_H + HAREA;         // W skewed: (k, j) at k * 32 + ((j + k) & 31)
  static constexpr int OFF_A3 = OFF_W + NA * 32;      // A3 (a, j) at a * 32 + j
  static constexpr int OFF_U = OFF_A3 + MEP * 32;     // U = A3 K^-1, same layout
  static constexpr int OFF_V = OFF_U + MEP * 32;      // 2 x 34: iteration vectors / broadcast rows (+ one scalar each)
  static constexpr int OFF_C = OFF_V + 68;            // xa0 [NA] | g = P_aa xa0 + q_a [NA] | diag P_aa [NA] | q_a [NA]
  static constexpr int OFF_B3 = OFF_C + 4 * NA;       // b3 [MEP], padded to 4 (4 NA is even: 16-byte alignment holds)
  static constexpr int OFF_HV = OFF_B3 + ((MEP + 3) & ~3);  // h [32]
  static constexpr int OFF_RHO = OFF_HV + 32;         // rho_i [32]
  static constexpr int OFF_AK = OFF_RHO + 32;         // extrapolation history: z, y at the previous point, their last increments [4][32]
  static constexpr int SMEM_DOUBLES = OFF_AK + 128;
  static constexpr unsigned FULL = 0xffffffffu;

  // warp sum; NOT inlined: ~55 call sites x 25 instructions would otherwise be a fifth of the kernel's code, and the
  // kernel is instruction-fetch bound before it is anything else (profiles/r2_warp_*)
  static __device__ QPC_NOINLINE double wsum(double a) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) a += __shfl_xor_sync(FULL, a, o);
    return a;
  }
  static __device__ __forceinline__ double wmax(double a) { return warp_max_nonneg_w(fabs(a)); }
  // 1 / x for a positive, normal x (pivots): hardware seed + two Newton steps, without the special-case code of the
  // IEEE division (25 instructions per pivot in a 100-instruction loop body)
  static __device__ __forceinline__ double rcp_pos(double x) {
#if defined(__CUDA_ARCH__)
    double r;
    asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(x));
    r = fma(r, fma(-x, r, 1.0), r);
    r = fma(r, fma(-x, r, 1.0), r);
    return r;
#else
    return 1.0 / x;
#endif
  }

  // 1 / x for a normal x of either sign (the IEEE division is ~25 instructions, most of them for cases that cannot occur here)
  static __device__ __forceinline__ double rcp_any(double x) { return copysign(rcp_pos(fabs(x)), x); }

  // x~ = t0 + T~ u for the vector u at `vec` (32 doubles, 16-byte aligned): four accumulation chains
  static __device__ __forceinline__ double row_times(const double (&t)[32], const double* vec, double t0) {
    double a0 = t0, a1 = 0.0, a2 = 0.0, a3 = 0.0;
    const double2* v2 = reinterpret_cast<const double2*>(vec);
#pragma unroll
    for (int c = 0; c < 16; c += 2) {
      const double2 p = v2[c], q = v2[c + 1];
      a0 = fma(t[2 * c], p.x, a0);
      a1 = fma(t[2 * c + 1], p.y, a1);
      a2 = fma(t[2 * c + 2], q.x, a2);
      a3 = fma(t[2 * c + 3], q.y, a3);
    }
    return (a0 + a1) + (a2 + a3);
  }

  // K = H + diag(rho) -> T~ = T diag(rho) in t[], t0; returns false on a non-positive pivot.  `rho_i` = this lane's rho.
  static __device__ __forceinline__ bool factor(double* sm, double (&t)[32], double& t0, double rho_i,
                                                const double (&a3)[MEP], const double (&b3)[MEP], int lane, bool gather) {
    double* Hs = sm + OFF_H;
    double* vb = sm + OFF_V;
    // K = H + diag(rho): the lane's diagonal entry is bumped in shared memory around the load (only this lane reads it)
    const double hdiag = Hs[lane * 33];
    Hs[lane * 33] = hdiag + rho_i;
#pragma unroll
    for (int i = 0; i < 32; i++) t[i] = Hs[i * 32 + lane];
    Hs[lane * 33] = hdiag;
    sm[OFF_RHO + lane] = rho_i;
    // Gauss-Jordan, pivot k at step k, as a ROLLED loop (instruction-cache footprint: the unrolled form was 10k
    // instructions and the kernel stalled 70 % of the time on instruction fetch).  The row registers rotate by one per
    // step -- the FMA that updates column c writes register c - 1 -- so the pivot column is always register 0 and the
    // new inverse column enters at register 31; after 32 steps the rotation is the identity.  The pivot row is not
    // scaled in place: lane k keeps its stored row and the pending factor s = 1 / pivot (applied once at the end); for
    // every other row the update in stored units is the same formula whether or not the row has been a pivot row
    // already, so the step is branch-free.
    // The reciprocal of the NEXT pivot is taken right after the first FMA of a step (register 0 of lane k + 1 is then
    // its diagonal entry) and travels with the pivot row: the ~60-cycle rcp chain overlaps the step's other 31 FMAs
    // instead of sitting between the broadcast and the multiplier of every step.
    // `gather`: the pivot row is GATHERED instead of broadcast.  K is symmetric and Gauss-Jordan keeps the working matrix
    // symmetric up to sign (rows that have been pivot rows carry -(their column)), so row k is the column-k entry of every
    // lane: one 8-byte store per lane -- f t[0], f = 1 before the lane's own pivot step, -(pending scale) after -- at the
    // rotated position, instead of sixteen 16-byte stores of lane k's registers that 31 lanes issue predicated off: the ADMM
    // stage of 16,384 solves 1.17 -> 1.01 ms.  Lane k keeps its own row, which equals the gathered one only up to the
    // asymmetry the rounding errors have built up: fine at eps 1e-5, a floor on the residuals at 1e-8 (solves stalled at the
    // iteration limit with eps_rel = 1e-16, i.e. 1e-11 relative), so such tolerances take the broadcast (WarpParams::gather_eps).
    double s = 1.0, f = 1.0;
    bool ok = true;
    double dk_mine = rcp_pos(t[0]);
#pragma unroll 1
    for (int k = 0; k < 32; k++) {
      double* rb = vb + (k & 1) * 34;
      if (gather) {
        rb[(lane - k) & 31] = f * t[0];
        if (lane == k) rb[32] = dk_mine;
      } else if (lane == k) {
        double2* r2 = reinterpret_cast<double2*>(rb);
#pragma unroll
        for (int c = 0; c < 16; c++) r2[c] = make_double2(t[2 * c], t[2 * c + 1]);
        rb[32] = dk_mine;
      }
      __syncwarp();
      const double2* r2 = reinterpret_cast<const double2*>(rb);
      const double2 p0 = r2[0];
      const double piv = p0.x;
      ok = ok && (piv > 0.0);
      const double dk = rb[32];
      const double m = (lane == k) ? 0.0 : t[0] * dk;
      t[0] = fma(-m, p0.y, t[1]);
      dk_mine = rcp_pos(t[0]);
#pragma unroll
      for (int c = 1; c < 16; c++) {
        const double2 p = r2[c];
        t[2 * c - 1] = fma(-m, p.x, t[2 * c]);
        t[2 * c] = fma(-m, p.y, t[2 * c + 1]);
      }
      t[31] = (lane == k) ? 1.0 : -m;
      if (lane == k) {
        s = dk;
        f = -dk;
      }
    }
#pragma unroll
    for (int i = 0; i < 32; i++) t[i] *= s;
    __syncwarp();
    // rank-ME correction for the rows A3 x = b3:  T = Ki - U' S^-1 U,  U = A3 Ki,  S = U A3'
    double w[MEP];
    if constexpr (ME > 0) {
      double u[ME];
#pragma unroll
      for (int a = 0; a < ME; a++) {
        const double2* r2 = reinterpret_cast<const double2*>(sm + OFF_A3 + a * 32);
        double acc0 = 0.0, acc1 = 0.0;
#pragma unroll
        for (int c = 0; c < 16; c++) {
          const double2 p = r2[c];
          acc0 = fma(t[2 * c], p.x, acc0);
          acc1 = fma(t[2 * c + 1], p.y, acc1);
        }
        u[a] = acc0 + acc1;
        sm[OFF_U + a * 32 + lane] = u[a];
      }
      double S[ME][ME];
#pragma unroll
      for (int a = 0; a < ME; a++)
#pragma unroll
        for (int b = a; b < ME; b++) S[a][b] = S[b][a] = wsum(a3[a] * u[b]);
      // S^-1 by Gauss-Jordan (ME x ME, every lane the same arithmetic)
      double Si[ME][ME];
#pragma unroll
      for (int a = 0; a < ME; a++)
#pragma unroll
        for (int b = 0; b < ME; b++) Si[a][b] = a == b ? 1.0 : 0.0;
#pragma unroll
      for (int k = 0; k < ME; k++) {
        ok = ok && (S[k][k] > 0.0);
        const double d = rcp_pos(S[k][k]);  // S is positive definite (checked just above)
#pragma unroll
        for (int b = 0; b < ME; b++) {
          S[k][b] *= d;
          Si[k][b] *= d;
        }
#pragma unroll
        for (int a = 0; a < ME; a++) {
          if (a == k) continue;
          const double f = S[a][k];
#pragma unroll
          for (int b = 0; b < ME; b++) {
            S[a][b] = fma(-f, S[k][b], S[a][b]);
            Si[a][b] = fma(-f, Si[k][b], Si[a][b]);
          }
        }
      }
#pragma unroll
      for (int a = 0; a < ME; a++) {
        double acc = 0.0;
#pragma unroll
        for (int b = 0; b < ME; b++) acc = fma(Si[a][b], u[b], acc);
        w[a] = acc;
      }
      __syncwarp();
#pragma unroll
      for (int a = 0; a < ME; a++) {
        const double2* r2 = reinterpret_cast<const double2*>(sm + OFF_U + a * 32);
#pragma unroll
        for (int c = 0; c < 16; c++) {
          const double2 p = r2[c];
          t[2 * c] = fma(-w[a], p.x, t[2 * c]);
          t[2 * c + 1] = fma(-w[a], p.y, t[2 * c + 1]);
        }
      }
    }
    // t0 = -T h + U' S^-1 b3, then fold rho into the columns
    {
      double c0 = 0.0;
      if constexpr (ME > 0) {
#pragma unroll
        for (int a = 0; a < ME; a++) c0 = fma(w[a], b3[a], c0);
      }
      t0 = -row_times(t, sm + OFF_HV, -c0);
      const double2* r2 = reinterpret_cast<const double2*>(sm + OFF_RHO);
#pragma unroll
      for (int c = 0; c < 16; c++) {
        const double2 p = r2[c];
        t[2 * c] *= p.x;
        t[2 * c + 1] *= p.y;
      }
    }
    return ok;
  }

  // dbg (optional, instance `base` only): H [1024] h [32] A3 [ME*32] b3 [ME] xa0 [NA] W [NA*32] cs
  // pflags: bit 0 = P_aa is diagonal; bits 8.. = block size of a block-diagonal P_bb (the controller's contact cost
  // 2 w_c B'B couples only the N multipliers of one contact point; 0 = dense)
  static __device__ void solve(const Settings& st, const WarpParams& wp, const AdmmProblem& pb, int n, int nbx,
                               int pflags, double* sm, int& fallback, double* dbg) {
    const int paa_diag = pflags & 1, pbb_block = pflags >> 8;
    const int lane = threadIdx.x & 31;
    const int na = NA;
    fallback = 0;  // reason code when the instance is handed back: 1 bounds, 2 rank of G_a, 3 rank of A3, 4 cost scale,
                   // 5 pivot, 6 non-finite residual
    double* Hs = sm + OFF_H;
    double* Ws = sm + OFF_W;
    double* vb = sm + OFF_V;
    double* Cs = sm + OFF_C;
    // ---- load: one matrix column per lane -------------------------------------------------------------------------------
    // G and b are staged through shared memory by a rolled, coalesced copy loop (the unrolled per-row global loads were
    // 740 instructions of straight-line code)
    double ca[MG], cb[MG];  // column `lane` of [G_a | b] (b in lane NA) and of G_b
    const bool hasb = lane < nbx;
    {
      double* Gs = sm + OFF_H;  // [MG][n] then b [MG]; MG (n + 1) <= HAREA + 32 NA doubles (checked below)
      const int tot = MG * n;
#if defined(__CUDA_ARCH__) && QPC_WARP_CPASYNC
      // asynchronous 8-byte copies global -> shared: all ~40 per lane in flight at once (the register-staged loop paid one
      // DRAM round trip per four elements: 7 % of the kernel's stall samples sat in this prologue)
#pragma unroll 4
      for (int idx = lane; idx < tot; idx += 32)
        asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"((unsigned)__cvta_generic_to_shared(Gs + idx)),
                     "l"(pb.G + idx));
      if (lane < MG)
        asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"((unsigned)__cvta_generic_to_shared(Gs + tot + lane)),
                     "l"(pb.lg + lane));
      asm volatile("cp.async.commit_group;");
      asm volatile("cp.async.wait_group 0;" ::: "memory");
#else
      for (int idx = lane; idx < tot; idx += 32) Gs[idx] = pb.G[idx];
      if (lane < MG) Gs[tot + lane] = pb.lg[lane];
#endif
      __syncwarp();
      const double* pa = lane < NA ? Gs + lane : Gs + tot;
      const int sa_ = lane < NA ? n : 1;
      const double* pbc = Gs + na + (hasb ? lane : 0);
#pragma unroll
      for (int r = 0; r < MG; r++) {
        const double va = pa[r * sa_], vb_ = pbc[r * n];
        ca[r] = lane <= NA ? va : 0.0;
        cb[r] = hasb ? vb_ : 0.0;
      }
      __syncwarp();
    }
    double lo = hasb ? fmax(pb.lb[lane], -QPC_INFTY) : 0.0;
    double up = hasb ? fmin(pb.ub[lane], QPC_INFTY) : 0.0;
    const double qb = hasb ? pb.qv[na + lane] : 0.0;
    const double qa = lane < NA ? pb.qv[lane] : 0.0;
    const double paa = lane < NA ? pb.P[(size_t)lane * n + lane] : 0.0;
    const double bn = wmax(lane < MG ? pb.lg[lane] : 0.0);  // |b|_inf
    const double qn = wmax(fmax(fabs(qa), fabs(qb)));
    int bad = (!(lo > -1e19 && up < 1e19) || !(lo <= up)) ? 1 : 0;
    // ---- Householder QR of G_a, reflectors applied to every column ---------------------------------------------------------
    // A rolled loop on ROTATING registers (instruction-cache footprint, and no masks): at step k the pivot row is
    // register 0; the update of row p writes register p - 1, the finished row k goes to shared memory (RA: columns of
    // [G_a | b], RB: columns of G_b) and a zero enters at register MG - 1.  Retired positions therefore hold zeros in
    // every column, including the owner's, so the reflector needs no length bookkeeping: v = column - alpha e_0 over all
    // MG registers.  After NA steps registers 0 .. ME-1 hold the rows that became A3 / b3.
    constexpr int NAP = QR_NAP;                   // row stride of RA and RS (even, >= NA + 1)
    double* RA = sm + OFF_H;                       // [NA][NAP]: R (upper triangle) and Q1'b in column NA
    double* RB = RA + NA * NAP;                    // [NA][32]:  Q1'G_b
    double* RS = RB + NA * 32;                     // [NA][NAP]: R in the order the back-substitution consumes it
    static_assert(NA * NAP * 2 + NA * 32 <= HAREA + NA * 32, "QR scratch must fit the H and W areas");
    static_assert(MG <= 32 && NA < 32 && NA <= MG, "one lane per row / column");
    static_assert(MG * (NA + 33) <= HAREA + NA * 32, "the staged copy of [G | b] must fit the H and W areas");
#pragma unroll 1
    for (int k = 0; k < NA; k++) {
      double* rb = vb + (k & 1) * 34;
      if (lane == k) {
        double n0 = 0.0, n1 = 0.0, n2 = 0.0, n3 = 0.0;  // four chains: this lane's serial section is every lane's wait
#pragma unroll
        for (int r = 0; r + 3 < MG; r += 4) {
          n0 = fma(ca[r], ca[r], n0);
          n1 = fma(ca[r + 1], ca[r + 1], n1);
          n2 = fma(ca[r + 2], ca[r + 2], n2);
          n3 = fma(ca[r + 3], ca[r + 3], n3);
        }
#pragma unroll
        for (int r = MG & ~3; r < MG; r++) n0 = fma(ca[r], ca[r], n0);
        const double nrm2 = (n0 + n1) + (n2 + n3);
        const double nrm = sqrt(nrm2);
        const double alpha = ca[0] > 0.0 ? -nrm : nrm;
        const double den = nrm2 - alpha * ca[0];  // = v'v / 2 with v = x - alpha e_0
        rb[0] = ca[0] - alpha;
#pragma unroll
        for (int r = 1; r < MG; r++) rb[r] = ca[r];
        rb[MG] = den > 0.0 ? rcp_pos(den) : 0.0;
      }
      __syncwarp();
      const double beta = rb[MG];
      double sa = 0.0, sb = 0.0, sa1 = 0.0, sb1 = 0.0;  // two chains per dot product
#pragma unroll
      for (int r = 0; r + 1 < MG; r += 2) {
        const double v0 = rb[r], v1 = rb[r + 1];
        sa = fma(v0, ca[r], sa);
        sb = fma(v0, cb[r], sb);
        sa1 = fma(v1, ca[r + 1], sa1);
        sb1 = fma(v1, cb[r + 1], sb1);
      }
      if (MG & 1) {
        sa = fma(rb[MG - 1], ca[MG - 1], sa);
        sb = fma(rb[MG - 1], cb[MG - 1], sb);
      }
      sa = (sa + sa1) * beta;
      sb = (sb + sb1) * beta;
      {
        const double v0 = rb[0];
        if (lane <= NA) RA[k * NAP + lane] = fma(-sa, v0, ca[0]);
        RB[k * 32 + lane] = fma(-sb, v0, cb[0]);
      }
#pragma unroll
      for (int r = 1; r < MG; r++) {  // the reflector is read again rather than kept: 2 MG registers less
        const double vr = rb[r];
        ca[r - 1] = fma(-sa, vr, ca[r]);
        cb[r - 1] = fma(-sb, vr, cb[r]);
      }
      ca[MG - 1] = 0.0;
      cb[MG - 1] = 0.0;
    }
    __syncwarp();
    // the rows that remain: A3 column `lane` and (lane NA) b3
    double a3[MEP], b3[MEP];
    if constexpr (ME > 0) {
#pragma unroll
      for (int a = 0; a < ME; a++) {
        a3[a] = cb[a];
        b3[a] = __shfl_sync(FULL, ca[a], NA);
      }
    } else {
      a3[0] = b3[0] = 0.0;
    }
    // rank test on the diagonal of R; RS[s] holds, for step s of the back-substitution (row r = NA-1-s), R[i-s][r] at
    // s <= i < NA-1, zero at i < s (solved entries only move), 1 / R[r][r] at NAP-1
    {
      const double ad = lane < NA ? fabs(RA[lane * NAP + lane]) : 0.0;
      const double dmax = wmax(ad);
      double dmin = lane < NA ? ad : 1e300;
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) dmin = fmin(dmin, __shfl_xor_sync(FULL, dmin, o));
      if (!(dmin > 1e-11 * dmax) || !finite_val(dmax)) bad = bad ? bad : 2;
      for (int idx = lane; idx < NA * NAP; idx += 32) {
        const int s_ = idx / NAP, i = idx - s_ * NAP, r = NA - 1 - s_;
        if (i == NAP - 1) continue;  // the reciprocals of the diagonal: below, one per lane (not a division inside this loop)
        double val = 0.0;
        if (i >= s_ && i - s_ < r) val = RA[(i - s_) * NAP + r];
        RS[idx] = val;
      }
      if (lane < NA) RS[(NA - 1 - lane) * NAP + NAP - 1] = rcp_any(RA[lane * NAP + lane]);
    }
    __syncwarp();
    // Back-substitution, column-oriented on ROTATING registers: at step s the entry to solve sits in register NA-1,
    // every register moves up by one while the unsolved ones take their update (c[i+1] = c[i] - R[i-s][r] w) and the
    // result enters at register 0; after NA steps c[r] = w_r.  Two right-hand sides per lane: cb (-> column of W) and
    // ca (lane NA: -> xa0).
#pragma unroll
    for (int r = 0; r < NA; r++) {
      cb[r] = RB[r * 32 + lane];
      ca[r] = RA[r * NAP + NA];
    }
#pragma unroll 1
    for (int st_ = 0; st_ < NA; st_++) {
      const double* rs = RS + st_ * NAP;
      const double ri = rs[NAP - 1];
      const double wB = cb[NA - 1] * ri, wA = ca[NA - 1] * ri;
#pragma unroll
      for (int i = NA - 2; i >= 0; i--) {
        const double rc = rs[i];
        cb[i + 1] = fma(-rc, wB, cb[i]);
        ca[i + 1] = fma(-rc, wA, ca[i]);
      }
      cb[0] = wB;
      ca[0] = wA;
    }
    __syncwarp();
    // publish W (unpadded copy in the scratch area for the broadcast reads of the H build, skewed copy for later),
    // xa0, diag P_aa, q_a
    double* Wu = Hs;
#pragma unroll
    for (int k = 0; k < NA; k++) {
      Wu[k * 32 + lane] = cb[k];
      Ws[k * 32 + ((lane + k) & 31)] = cb[k];
      if (lane == 0) Cs[k] = ca[k];  // every lane solved the right-hand side column: xa0
    }
    if (lane < NA) {
      Cs[2 * NA + lane] = paa;
      Cs[3 * NA + lane] = qa;
    }
    __syncwarp();
    // g = P_aa xa0 + q_a  (lane k < NA)
    {
      double g = 0.0;
      if (lane < NA) {
        if (paa_diag) {
          g = fma(paa, Cs[lane], qa);
        } else {
          g = qa;
          for (int l = 0; l < NA; l++) g = fma(pb.P[(size_t)lane * n + l], Cs[l], g);
        }
        Cs[NA + lane] = g;
      }
    }
    __syncwarp();
    // A3 and b3: orthonormalised rows (modified Gram-Schmidt, two passes)
    if constexpr (ME > 0) {
#pragma unroll
      for (int a = 0; a < ME; a++) {
        const double n0 = wsum(a3[a] * a3[a]);
#pragma unroll
        for (int pass = 0; pass < 2; pass++) {
#pragma unroll
          for (int c = 0; c < a; c++) {
            const double d = wsum(a3[c] * a3[a]);
            a3[a] = fma(-d, a3[c], a3[a]);
            b3[a] = fma(-d, b3[c], b3[a]);
          }
        }
        const double n1 = wsum(a3[a] * a3[a]);
        if (!(n1 > 1e-20 * n0) || !(n0 > 0.0)) bad = bad ? bad : 3;
        const double inv = 1.0 / sqrt(n1);
        a3[a] *= inv;
        b3[a] *= inv;
      }
#pragma unroll
      for (int a = 0; a < ME; a++) {
        sm[OFF_A3 + a * 32 + lane] = a3[a];
        if (lane == 0) sm[OFF_B3 + a] = b3[a];
      }
    }
    // ---- reduced Hessian column `lane`: P_bb + W'(P_aa W), and h --------------------------------------------------------------
    double t[32];
    double hj = qb;
    if (QPC_WARP_DMMA && paa_diag) {
      // W'(P_aa W) as 4 x 4 tiles of 8 x 8 on the fp64 tensor cores: 96 DMMA.8x8x4 instead of 672 DFMA per lane (the
      // kernel is bound by issue slots and latency, not by the fp64 pipe: profiles/r2_warp_dmma_*).  Operands come from the
      // skewed W in shared memory; the accumulator fragments go straight to H in shared memory.
      const int gid = lane >> 2, tig = lane & 3;
      constexpr int KS = (NA + 3) / 4;
      double aw[4][KS], pk[KS];
#pragma unroll
      for (int ks = 0; ks < KS; ks++) {
        const int k = 4 * ks + tig;
        const bool valid = k < NA;
        const int kk = valid ? k : 0;
        pk[ks] = valid ? Cs[2 * NA + kk] : 0.0;
#pragma unroll
        for (int tq = 0; tq < 4; tq++) aw[tq][ks] = valid ? Ws[kk * 32 + ((8 * tq + gid + kk) & 31)] : 0.0;
      }
#pragma unroll 1
      for (int k = 0; k < NA; k++) hj = fma(-Ws[k * 32 + ((lane + k) & 31)], Cs[NA + k], hj);
      __syncwarp();  // the scratch copy of W in the H area is dead (the A3 orthonormalisation was its last barrier)
#pragma unroll
      for (int ti = 0; ti < 4; ti++) {
#pragma unroll
        for (int tj = 0; tj < 4; tj++) {
          double c0 = 0.0, c1 = 0.0;
#pragma unroll
          for (int ks = 0; ks < KS; ks++) dmma884(c0, c1, aw[ti][ks], pk[ks] * aw[tj][ks]);
          *reinterpret_cast<double2*>(Hs + (8 * ti + gid) * 32 + 8 * tj + 2 * tig) = make_double2(c0, c1);
        }
      }
      __syncwarp();
    } else {
#pragma unroll
      for (int i = 0; i < 32; i++) t[i] = 0.0;
#pragma unroll 1
      for (int k = 0; k < NA; k++) {
        const double wk = Ws[k * 32 + ((lane + k) & 31)];  // W[k][lane]
        double u;                                          // (P_aa W)[k][lane]
        if (paa_diag) {
          u = Cs[2 * NA + k] * wk;
        } else {
          u = 0.0;
          for (int l = 0; l < NA; l++) u = fma(pb.P[(size_t)k * n + l], Wu[l * 32 + lane], u);
        }
        hj = fma(-wk, Cs[NA + k], hj);
        const double2* w2 = reinterpret_cast<const double2*>(Wu + k * 32);
#pragma unroll
        for (int c = 0; c < 16; c++) {
          const double2 p = w2[c];
          t[2 * c] = fma(p.x, u, t[2 * c]);
          t[2 * c + 1] = fma(p.y, u, t[2 * c + 1]);
        }
      }
      __syncwarp();  // everyone is done reading the scratch copy of W
#pragma unroll
      for (int i = 0; i < 32; i++) Hs[i * 32 + lane] = t[i];
    }
    // + P_bb (rows beyond nbx: identity padding), by a rolled loop: row i of P_bb is read coalesced, which by symmetry is
    // column i of every lane's row
    {
      const double* pp = pb.P + (size_t)na * n + na + lane;
      if (pbb_block > 0) {
        // block-diagonal P_bb: the lane's column has entries only in the rows of its own contact point -- N loads instead
        // of 32 (1 KB instead of 8 KB of P read per solve)
        const int b0 = (lane / pbb_block) * pbb_block;
        if (hasb) {
#pragma unroll 1
          for (int i = b0; i < b0 + pbb_block && i < nbx; i++) Hs[i * 32 + lane] += pp[(size_t)i * n];
        } else {
          Hs[lane * 32 + lane] += 1.0;
        }
      } else
#pragma unroll 1
      for (int i0 = 0; i0 < 32; i0 += QPC_WARP_PBATCH) {  // QPC_WARP_PBATCH loads in flight before the first use
        double pv[QPC_WARP_PBATCH];
#pragma unroll
        for (int i = 0; i < QPC_WARP_PBATCH; i++)
          pv[i] = (hasb && i0 + i < nbx) ? pp[(size_t)(i0 + i) * n] : ((i0 + i == lane && !hasb) ? 1.0 : 0.0);
#pragma unroll
        for (int i = 0; i < QPC_WARP_PBATCH; i++) Hs[(i0 + i) * 32 + lane] += pv[i];
      }
    }
    const double tr = hasb ? Hs[lane * 32 + lane] : 0.0;
    sm[OFF_HV + lane] = hj;
    const double cs = wsum(tr) / (nbx > 0 ? nbx : 1);  // cost scale: rho is quoted relative to the mean curvature
    if (!(cs > 0.0) || !finite_val(cs)) bad = bad ? bad : 4;
    __syncwarp();
#if defined(QPC_WARP_DEBUG_DUMP) || defined(QPC_WARP_EMU)
    if (dbg) {
      for (int i = 0; i < 32; i++) dbg[i * 32 + lane] = Hs[i * 32 + lane];
      dbg[1024 + lane] = hj;
#pragma unroll
      for (int a = 0; a < ME; a++) dbg[1056 + a * 32 + lane] = a3[a];
      if (lane == 0) {
#pragma unroll
        for (int a = 0; a < ME; a++) dbg[1056 + ME * 32 + a] = b3[a];
        for (int k = 0; k < NA; k++) dbg[1056 + ME * 33 + k] = Cs[k];
        dbg[1056 + ME * 33 + NA + NA * 32] = cs;
      }
      for (int k = 0; k < NA; k++) dbg[1056 + ME * 33 + NA + k * 32 + lane] = Ws[k * 32 + ((lane + k) & 31)];
    }
#else
    (void)dbg;
#endif
    bad = __reduce_max_sync(FULL, bad);
    if (bad) {
      fallback = bad;
      return;
    }
    // ---- ADMM -----------------------------------------------------------------------------------------------------------
    const bool iseq = (up - lo) < QPC_RHO_TOL;
    const bool adaptive = st.adaptive_rho != 0;
    const double kap = adaptive ? wp.kappa : 1.0, kinv = 1.0 / kap;
    double rho = st.rho * cs;
    double z = 0.0, yr = 0.0;
    bool act = false;
    bool warm = false;
    if (pb.rho_io && pb.x0 && pb.y0) {
      const double r = *pb.rho_io;
      warm = r > 0.0 && r < 1e30;
      if (warm) {
        rho = r * cs;
        z = hasb ? pb.x0[na + lane] : 0.0;
        act = (z <= lo) || (z >= up);
      }
    }
    auto rho_of = [&](bool active) { return iseq ? QPC_RHO_EQ_FACTOR * rho : (active ? kap * rho : kinv * rho); };
    double rho_i = rho_of(act);
    if (warm) yr = (hasb ? pb.y0[MG + lane] : 0.0) / rho_i;
    double t0 = 0.0;
    int nfac = 0;
    int status = -10, iter = 0;
    int next_adapt = (adaptive && wp.first > 0) ? wp.first : 0x7fffffff;
    const int chk = wp.check > 0 ? wp.check : 0x7fffffff;
    // warm-started solves (closed loop near a fixed point, restarts from a solution) often need a handful of iterations:
    // they get dense checks (every 5) up to the first full check at 25; cold solves only pay for checks that can succeed
    const int chk_early = min(chk, 5);
    int next_chk = warm ? chk_early : chk;
    // iterations that take the full check whatever `chk` is: multiples of 25 (infeasibility test) and of the extrapolation
    // interval; `next_full` = the next of them
    const int ait = wp.aitken > 0 ? wp.aitken : 0x7fffffff;
    auto full_after = [&](int it) { return min((it / 25 + 1) * 25, ait < 0x7fffffff ? (it / ait + 1) * ait : 0x7fffffff); };
    int next_full = full_after(0);
    const double alpha = st.alpha, oma = 1.0 - st.alpha;
    double pri_res = 0.0, dua_res = 0.0, xt = 0.0;
    double ds_last = 1e300;  // last evaluated dual normaliser max(|Px|, |A'y|, |q|)
    double ps_last = 1e300;  // primal normaliser max(|Ax|, |z|) of the last full check (1e300: none yet)
    double rp_last = -1.0;   // primal residual at the previous adaptation point
    bool refactor = true;
    int hist = 0;            // extrapolation history: 0 none, 1 a previous point, 2 also a previous increment
    for (;;) {
      if (refactor) {
        hist = 0;  // the only call site: the inversion is ~1,400 instructions per inlined copy
        const bool okf = factor(sm, t, t0, rho_i, a3, b3, lane, st.eps_abs >= wp.gather_eps || st.eps_rel >= 1e-2 * wp.gather_eps);
        nfac++;
        refactor = false;
        if (!__all_sync(FULL, okf)) {
          fallback = 5;
          return;
        }
      }
      // plain iterations up to the next special one
      const int stop = min(min(min(next_chk, next_full), next_adapt), st.max_iter);
#pragma unroll 1
      for (; iter < stop - 1; iter++) {
        double* ub_ = vb + (iter & 1) * 34;
        ub_[lane] = z - yr;
        __syncwarp();
        xt = row_times(t, ub_, t0);
        const double w = fma(alpha, xt, fma(oma, z, yr));
        double zn = w < lo ? lo : w;
        zn = zn > up ? up : zn;
        yr = w - zn;
        z = zn;
      }
      if (iter >= st.max_iter) {  // max_iter < 1
        status = -2;
        break;
      }
      // the special iteration: same update, old values kept for the residuals
      double* ub_ = vb + (iter & 1) * 34;
      ub_[lane] = z - yr;
      __syncwarp();
      xt = row_times(t, ub_, t0);
      const double zo = z, yro = yr;
      {
        const double w = fma(alpha, xt, fma(oma, z, yr));
        double zn = w < lo ? lo : w;
        zn = zn > up ? up : zn;
        yr = w - zn;
        z = zn;
      }
      iter++;
      const bool adapt = iter == next_adapt, last = iter >= st.max_iter;
      const double dy = rho_i * (yr - yro);
      const double rdv = dy - rho_i * (xt - zo);  // = H x~ + h + A3'nu + y+ : stationarity defect of the x_b rows
      // Quick reject at the dense check points (one vote instead of four warp maxima and the bookkeeping of a full check):
      // against the normalisers of the last full check, with slack for their drift, a lane whose own residual is above the
      // tolerance proves "not converged".  It can only delay a termination to the next full check (every 25th iteration,
      // the extrapolation and adaptation points), never cause one.
      if (!adapt && !last && iter != next_full &&
          __any_sync(FULL, fabs(xt - z) >= st.eps_abs + st.eps_rel * 2.0 * ps_last ||
                               fabs(rdv) >= st.eps_abs + st.eps_rel * 8.0 * ds_last)) {
        next_chk = iter + ((warm && iter < 25) ? chk_early : chk);
        continue;
      }
      pri_res = wmax(xt - z);
      dua_res = wmax(rdv);
      if (!finite_val(pri_res) || !finite_val(dua_res)) {
        fallback = 6;
        return;
      }
      const double xzmax = fmax(wmax(xt), wmax(z));
      const double ps = fmax(bn, xzmax);
      ps_last = ps;
      const bool pok = pri_res < st.eps_abs + st.eps_rel * ps;
      bool done = false;
      if (pok && dua_res < st.eps_abs) {
        status = 1;
        done = true;
      }
      double ds = ds_last;
      const bool want_ds = adapt || last || (pok && st.eps_rel > 0.0 && dua_res < st.eps_abs + st.eps_rel * 8.0 * ds_last);
      if (!done && want_ds) {
        // full-problem normaliser: x_a = xa0 - W x~;  (Px)_a = P_aa x_a;  (Px)_b = H x~ - W' P_aa W x~;
        // A'y = -(Px + q) + (0, rdv)  (the x_a rows of the stationarity condition hold exactly)
        double* xb_ = vb + (iter & 1) * 34;  // the buffer the special iteration did not read
        xb_[lane] = xt;
        __syncwarp();
        const int wrow = lane < NA ? lane : 0;
        double wx, hx;  // (W x~)_lane for lane < NA, (H x~)_lane
        {
          const double2* x2 = reinterpret_cast<const double2*>(xb_);
          double a0 = 0.0, a1 = 0.0, h0 = 0.0, h1 = 0.0;
#pragma unroll
          for (int c = 0; c < 16; c++) {
            const double2 p = x2[c];
            a0 = fma(Ws[wrow * 32 + ((2 * c + wrow) & 31)], p.x, a0);
            a1 = fma(Ws[wrow * 32 + ((2 * c + 1 + wrow) & 31)], p.y, a1);
            h0 = fma(Hs[(2 * c) * 32 + lane], p.x, h0);
            h1 = fma(Hs[(2 * c + 1) * 32 + lane], p.y, h1);
          }
          wx = lane < NA ? a0 + a1 : 0.0;
          hx = h0 + h1;
        }
        const double xa = lane < NA ? Cs[lane] - wx : 0.0;
        double pxa = 0.0, pwx = 0.0;  // (P_aa x_a)_lane, (P_aa W x~)_lane
        if (paa_diag) {
          pxa = paa * xa;
          pwx = paa * wx;
        } else {
          __syncwarp();
          xb_[lane] = xa;  // dense P_aa: publish x_a (W x~ = xa0 - x_a)
          __syncwarp();
          if (lane < NA)
            for (int l = 0; l < NA; l++) {
              const double pv = pb.P[(size_t)lane * n + l];
              pxa = fma(pv, xb_[l], pxa);
              pwx = fma(pv, Cs[l] - xb_[l], pwx);
            }
        }
        __syncwarp();
        xb_[lane] = pwx;
        __syncwarp();
        double wtp = 0.0;  // (W' P_aa W x~)_lane
#pragma unroll
        for (int k = 0; k < NA; k++) wtp = fma(Ws[k * 32 + ((lane + k) & 31)], xb_[k], wtp);
        const double pxb = hasb ? hx - wtp : 0.0;
        const double atya = -(pxa + qa), atyb = hasb ? -(pxb + qb) + rdv : 0.0;
        ds = fmax(qn, fmax(wmax(fmax(fabs(pxa), fabs(pxb))), wmax(fmax(fabs(atya), fabs(atyb)))));
        ds_last = ds;
        __syncwarp();
      }
      if (!done && want_ds && pok && dua_res < st.eps_abs + st.eps_rel * ds) {
        status = 1;
        done = true;
      }
      if (!done && !pok && (iter % 25 == 0 || last)) {
        // primal infeasibility of {A3 x = b3, lb <= x <= ub}: dy + A3'mu = 0 and u'dy+ + l'dy- + b3'mu < 0
        const double ndy = wmax(dy);
        if (ndy > st.eps_prim_inf) {
          double sup = hasb ? up * fmax(dy, 0.0) + lo * fmin(dy, 0.0) : 0.0;
          double res = dy;
          if constexpr (ME > 0) {
#pragma unroll
            for (int a = 0; a < ME; a++) {
              const double mu = -wsum(a3[a] * dy);
              sup = fma(lane == 0 ? b3[a] : 0.0, mu, sup);
              res = fma(a3[a], mu, res);
            }
          }
          sup = wsum(sup);
          if (sup < -st.eps_prim_inf * ndy && wmax(res) < st.eps_prim_inf * ndy) {
            status = -3;
            done = true;
          }
        }
      }
      if (!done && last) {
        const bool p10 = pri_res < 10.0 * (st.eps_abs + st.eps_rel * ps);
        const bool d10 = dua_res < 10.0 * (st.eps_abs + st.eps_rel * (ds < 1e299 ? ds : 0.0));
        status = (p10 && d10) ? 2 : -2;
        done = true;
      }
      if (done) break;
      if (iter == next_chk) next_chk += (warm && iter < 25) ? chk_early : chk;
      if (iter == next_full) next_full = full_after(iter);
      if (adapt) {
        // schedule: every `first` iterations early on, then geometric (x growth): solves that need several re-weightings of
        // their rows get them quickly (they used to sit out the gaps of a x2 schedule: 410 iterations instead of 150),
        // while the number of refactorisations of a long solve stays logarithmic in its length
        next_adapt = (int)fmin(fmax(ceil(next_adapt * wp.growth), (double)next_adapt + wp.first), 2.0e9);
        // OSQP's rule without its 1e-10 guards: at tight tolerances (residual / norm ~ 1e-11) they saturate the ratio
        // and stop rho from growing when the primal residual sits on its rounding floor (~ cond(K) eps |x|)
        const double prn = pri_res / (ps + 1e-300), drn = dua_res / (ds + 1e-300);
        double rn = rho * sqrt(prn / (drn + 1e-300));
        // rho floor: the explicit inverse of K = H + diag(rho) puts a rounding floor ~ eps_mach |x| lambda_max(H) / rho_row
        // on the primal residual (lambda_max <= trace(H) = nbx cs); keep it below eps_abs
        const double rho_floor = kap * 2.2e-16 * fmax(xzmax, 1.0) * (double)nbx * cs / st.eps_abs;
        rn = fmin(fmax(rn, fmax(QPC_RHO_MIN * cs, rho_floor)), QPC_RHO_MAX * cs);
        // a primal residual that has stopped falling since the last adaptation takes any increase of rho, not only 5x ones
        const bool stalled = !pok && rp_last >= 0.0 && pri_res >= 0.5 * rp_last && rn > 1.5 * rho;
        rp_last = pri_res;
        const bool big = stalled || rn > rho * st.adaptive_rho_tolerance || rn < rho / st.adaptive_rho_tolerance;
        const bool an = (z <= lo) || (z >= up);
        const bool changed = kap != 1.0 && __any_sync(FULL, an != act);
#if defined(QPC_WARP_EMU) && defined(QPC_WARP_ADAPT_STATS)
        {  // development statistics (tests/emu only): iteration, global change, rows that change their rho class
          const int nch = __reduce_max_sync(FULL, 0) + 0;  // keep the fibres in step
          (void)nch;
          int cnt = 0;
          for (int l = 0; l < 32; l++) cnt += (__shfl_sync(FULL, (double)(an != act), l) != 0.0) ? 1 : 0;
          if (lane == 0) qpc_adapt_stats(iter, big ? 1 : 0, cnt);
        }
#endif
        if (big || changed) {
          if (big) rho = rn;
          act = an;
          const double rnew = rho_of(act);
          yr *= rho_i / rnew;  // y is unchanged by a rho update; yr = y / rho follows the new rho
          rho_i = rnew;
          refactor = true;
        }
      } else if (wp.aitken > 0 && iter % wp.aitken == 0) {
        // Extrapolation of slowly converging solves (between adaptations): once the active set has settled the iteration
        // is affine, s+ = M s + c on s = (z, y); when one real mode dominates, successive increments are parallel,
        // d_k = r d_{k-1}, and the limit is s + d r / (1 - r) (Aitken); r -> 1 is the dual drift of a wrongly active row.
        // The step is cut so that no clipped row reaches its release point (y crossing 0) and no interior row leaves
        // the box: the active set, hence the affine regime, is preserved, and ADMM simply continues from the new point
        // (it converges from any (z, y) at fixed rho).  The heavy tail of the iteration count -- 1,755 iterations on
        // the worst of 131,072 Atlas states -- becomes 265; typical solves never get here.
        double* AK = sm + OFF_AK;
        const double yv = rho_i * yr;
        bool jumped = false;
        if (hist >= 1) {
          const double dz = z - AK[lane], dyv = yv - AK[32 + lane];
          if (hist >= 2) {
            const double pz = AK[64 + lane], py = AK[96 + lane];
            const double dd = wsum(fma(dz, dz, dyv * dyv)), pp = wsum(fma(pz, pz, py * py)), dp = wsum(fma(dz, pz, dyv * py));
            if (dd > 0.0 && pp > 0.0) {
              const double cosv = dp / sqrt(dd * pp), rr = dp / pp;
              if (cosv > 1.0 - 1e-4 && rr > 0.0 && rr < 1.0 - 1e-12) {
                double bnd = 1e300;
                const bool clipped = ((z <= lo) || (z >= up)) && !iseq;
                if (clipped) {
                  if (yv * dyv < 0.0) bnd = 0.9 * (-yv / dyv);
                } else if (!iseq) {
                  if (dz > 0.0) bnd = 0.9 * (up - z) / dz;
                  else if (dz < 0.0) bnd = 0.9 * (lo - z) / dz;
                }
#pragma unroll
                for (int o = 16; o > 0; o >>= 1) bnd = fmin(bnd, __shfl_xor_sync(FULL, bnd, o));
                const double gain = fmin(rr / (1.0 - rr), bnd);
                if (gain >= 1.0) {
                  double zn = fma(gain, dz, z);
                  zn = zn < lo ? lo : zn;
                  z = zn > up ? up : zn;
                  yr = fma(gain, dyv, yv) / rho_i;
                  hist = 0;
                  jumped = true;
                }
              }
            }
          }
          if (!jumped) {
            AK[64 + lane] = dz;
            AK[96 + lane] = dyv;
            hist = 2;
          }
        }
        if (!jumped) {
          AK[lane] = z;
          AK[32 + lane] = yv;
          if (hist < 1) hist = 1;
        }
      }
    }
    // ---- store: x_b = x~ (satisfies the equality rows exactly), x_a = xa0 - W x_b, y_b ----------------------------------------
    {
      double* xb_ = vb;
      __syncwarp();
      xb_[lane] = xt;
      __syncwarp();
      if (lane < NA) {
        double a0 = Cs[lane];
#pragma unroll 4
        for (int j = 0; j < 32; j++) a0 = fma(-Ws[lane * 32 + ((j + lane) & 31)], xb_[j], a0);
        pb.x[lane] = a0;
      }
      if (hasb) {
        pb.x[na + lane] = xt;
        if (pb.y) pb.y[MG + lane] = rho_i * yr;
      }
      if (pb.y && lane < MG) pb.y[lane] = 0.0;  // multipliers of the eliminated rows are not carried
      if (lane == 0) {
        *pb.status = status;
        if (pb.iters) *pb.iters = iter;
        if (pb.nfac) *pb.nfac = nfac;
        if (pb.rho_io) *pb.rho_io = (status == 1 || status == 2) ? rho / cs : -1.0;
        if (pb.res) {
          pb.res[0] = pri_res;
          pb.res[1] = dua_res;
        }
      }
    }
  }
};

#endif  // __CUDACC__ || QPC_WARP_EMU

}  // namespace qpc
