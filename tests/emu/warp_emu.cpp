// TEST INFRASTRUCTURE ONLY -- CPU emulation of ONE WARP for the kernel body in qpcontrol.jl_b200/csrc/admm_warp.cuh.
// The 32 lanes are ucontext fibres run round-robin by one OS thread: every warp-synchronous primitive (__syncwarp,
// shuffles, votes, redux) is a yield point, so the lanes advance in lock-step from primitive to primitive exactly as
// a converged warp does.  Lets the reduction / inversion / iteration logic of the one-warp ADMM kernel be debugged and
// tested against the numpy prototype and the oracle without a GPU.  The product library never loads this file.
#include <math.h>
#include <stdint.h>
#include <string.h>
#include <ucontext.h>

#include <algorithm>
#include <vector>

#define QPC_WARP_EMU 1
#define __device__
#define __forceinline__ inline
#define __global__

namespace emu {
constexpr int W = 32;
static ucontext_t g_main, g_ctx[W];
static int g_lane = 0;
static bool g_done[W];
static double g_slot_d[2][W];
static unsigned g_slot_u[2][W];
static int g_phase[W];  // per-lane count of exchange primitives, selects the slot buffer
struct Tid {
  int x;
};
static inline void yield_lane() {  // hand over to the next live lane (round-robin), or back to the scheduler
  int cur = g_lane;
  for (int k = 1; k <= W; k++) {
    int nxt = (cur + k) % W;
    if (!g_done[nxt]) {
      if (nxt == cur) return;
      g_lane = nxt;
      swapcontext(&g_ctx[cur], &g_ctx[nxt]);
      return;
    }
  }
}
}  // namespace emu

#define threadIdx (emu::Tid{emu::g_lane})
static inline void __syncwarp(unsigned = 0xffffffffu) { emu::yield_lane(); }
static inline double __shfl_xor_sync(unsigned, double v, int o) {
  const int me = emu::g_lane, b = emu::g_phase[me]++ & 1;
  emu::g_slot_d[b][me] = v;
  emu::yield_lane();
  return emu::g_slot_d[b][me ^ o];
}
static inline double __shfl_sync(unsigned, double v, int src) {
  const int me = emu::g_lane, b = emu::g_phase[me]++ & 1;
  emu::g_slot_d[b][me] = v;
  emu::yield_lane();
  return emu::g_slot_d[b][src];
}
static inline unsigned reduce_max_u(unsigned v) {
  const int me = emu::g_lane, b = emu::g_phase[me]++ & 1;
  emu::g_slot_u[b][me] = v;
  emu::yield_lane();
  unsigned m = 0;
  for (int i = 0; i < emu::W; i++) m = std::max(m, emu::g_slot_u[b][i]);
  return m;
}
static inline unsigned __reduce_max_sync(unsigned, unsigned v) { return reduce_max_u(v); }
static inline int __reduce_max_sync(unsigned, int v) { return (int)reduce_max_u((unsigned)v); }  // non-negative ints only
static inline bool __any_sync(unsigned, bool p) { return reduce_max_u(p ? 1u : 0u) != 0u; }
static inline bool __all_sync(unsigned, bool p) { return reduce_max_u(p ? 0u : 1u) == 0u; }
static inline int __double2hiint(double v) {
  uint64_t u;
  memcpy(&u, &v, 8);
  return (int)(u >> 32);
}
static inline int __double2loint(double v) {
  uint64_t u;
  memcpy(&u, &v, 8);
  return (int)(u & 0xffffffffu);
}
static inline double __hiloint2double(int hi, int lo) {
  uint64_t u = ((uint64_t)(uint32_t)hi << 32) | (uint32_t)lo;
  double v;
  memcpy(&v, &u, 8);
  return v;
}
struct double2 {
  double x, y;
};
static inline double2 make_double2(double x, double y) { return double2{x, y}; }
using std::max;
using std::min;
// mma.sync.m8n8k4.f64 on the fibres: A[lane/4][lane%4], B[lane%4][lane/4] published, D[lane/4][2 (lane%4) + {0,1}] computed
static inline void dmma884(double& d0, double& d1, double a, double b) {
  const int me = emu::g_lane, buf = emu::g_phase[me]++ & 1;
  static double sa[2][emu::W], sb[2][emu::W];
  sa[buf][me] = a;
  sb[buf][me] = b;
  emu::yield_lane();
  const int row = me >> 2, c0 = 2 * (me & 3);
  for (int k = 0; k < 4; k++) {
    d0 = fma(sa[buf][row * 4 + k], sb[buf][c0 * 4 + k], d0);
    d1 = fma(sa[buf][row * 4 + k], sb[buf][(c0 + 1) * 4 + k], d1);
  }
}

#include "../../qpcontrol.jl_b200/csrc/admm_warp.cuh"

using namespace qpc;

namespace {
struct Job {
  const Settings* st;
  const WarpParams* wp;
  AdmmProblem pb;
  int n, nbx, paa_diag;
  double* smem;
  double* dbg;
  int fallback[emu::W];
};
Job* g_job;
template <int MG, int NA>
void lane_main() {
  Job& j = *g_job;
  int fb = 0;
  WarpSolver<MG, NA>::solve(*j.st, *j.wp, j.pb, j.n, j.nbx, j.paa_diag, j.smem, fb, j.dbg);
  j.fallback[emu::g_lane] = fb;
  emu::g_done[emu::g_lane] = true;
  // hand over to another live lane, or to the scheduler when this was the last one
  for (int k = 1; k < emu::W; k++) {
    int nxt = (emu::g_lane + k) % emu::W;
    if (!emu::g_done[nxt]) {
      int cur = emu::g_lane;
      emu::g_lane = nxt;
      swapcontext(&emu::g_ctx[cur], &emu::g_ctx[nxt]);
    }
  }
  setcontext(&emu::g_main);
}
template <int MG, int NA>
int run_warp(Job& job) {
  static std::vector<char> stacks(emu::W * (1 << 18));
  g_job = &job;
  for (int l = 0; l < emu::W; l++) {
    emu::g_done[l] = false;
    emu::g_phase[l] = 0;
    getcontext(&emu::g_ctx[l]);
    emu::g_ctx[l].uc_stack.ss_sp = stacks.data() + (size_t)l * (1 << 18);
    emu::g_ctx[l].uc_stack.ss_size = 1 << 18;
    emu::g_ctx[l].uc_link = &emu::g_main;
    makecontext(&emu::g_ctx[l], (void (*)())lane_main<MG, NA>, 0);
  }
  emu::g_lane = 0;
  swapcontext(&emu::g_main, &emu::g_ctx[0]);
  return job.fallback[0];
}
}  // namespace

extern "C" {

// Solves B QPs with the one-warp kernel body (MG = 24, NA = 21: the StandingController shape).  settings: the same
// struct the C ABI takes is not needed here -- plain arguments.  Returns 0, or -1 for an unsupported shape.
// fallback[i] receives the hand-back reason code (0 = solved by the warp body).
int emu_warp_solve_qp_batch(int64_t B, int32_t n, int32_t mg, int32_t nbx, const double* P, const double* qv, const double* G,
                            const double* lg, const double* lb, const double* ub, double rho, double alpha,
                            double eps_abs, double eps_rel, double eps_prim_inf, int32_t max_iter, int32_t adaptive_rho,
                            double adaptive_rho_tolerance, double kappa, double growth, int32_t first, int32_t check, int32_t aitken,
                            int32_t paa_diag, int32_t warm, double* x, double* y, double* rho_io, int32_t* status,
                            int32_t* iters, double* res, int32_t* nfac, int32_t* fallback, double* dbg) {
  const bool s2421 = mg == 24 && n - nbx == 21, s3027 = mg == 30 && n - nbx == 27;
  if (!((s2421 || s3027) && nbx >= 1 && nbx <= 32)) return -1;
  Settings st;
  memset(&st, 0, sizeof(st));
  st.rho = rho;
  st.sigma = 1e-6;
  st.alpha = alpha;
  st.eps_abs = eps_abs;
  st.eps_rel = eps_rel;
  st.eps_prim_inf = eps_prim_inf;
  st.eps_dual_inf = 1e-4;
  st.adaptive_rho_tolerance = adaptive_rho_tolerance;
  st.max_iter = max_iter;
  st.scaling = 10;
  st.adaptive_rho = adaptive_rho;
  st.adaptive_rho_interval = 25;
  st.check_termination = 25;
  WarpParams wp;
  wp.kappa = kappa;
  wp.growth = growth;
  wp.first = first;
  wp.check = check;
  wp.aitken = aitken;
  wp.gather_eps = 1e-6;
  std::vector<double> smem((s2421 ? WarpSolver<24, 21>::SMEM_DOUBLES : WarpSolver<30, 27>::SMEM_DOUBLES) + 8);
  for (int64_t i = 0; i < B; i++) {
    Job job;
    job.st = &st;
    job.wp = &wp;
    job.n = n;
    job.nbx = nbx;
    job.paa_diag = paa_diag;
    job.smem = smem.data();
    job.dbg = (dbg && i == 0) ? dbg : nullptr;
    AdmmProblem& pb = job.pb;
    pb.P = P + i * n * n;
    pb.qv = qv + i * n;
    pb.G = G + i * mg * n;
    pb.lg = lg + i * mg;
    pb.ug = lg + i * mg;
    pb.lb = lb + i * nbx;
    pb.ub = ub + i * nbx;
    pb.x = x + i * n;
    pb.y = y + i * (mg + nbx);
    pb.status = status + i;
    pb.iters = iters + i;
    pb.res = res + 2 * i;
    pb.nfac = nfac + i;
    pb.rho_io = rho_io + i;  // always written; read only on a warm start
    if (warm) {
      pb.x0 = pb.x;
      pb.y0 = pb.y;
    }
    fallback[i] = s2421 ? run_warp<24, 21>(job) : run_warp<30, 27>(job);
    if (fallback[i]) status[i] = QPC_WARP_FALLBACK;
  }
  return 0;
}

}  // extern "C"
