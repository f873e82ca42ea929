// TEST INFRASTRUCTURE ONLY -- CPU emulation of the CUDA kernel bodies in qpcontrol.jl_b200/csrc/*.cuh.
// The bodies are compiled by g++ as a single "thread" (QPC_TID = 0, QPC_NT = 1, barriers no-ops, see qpc_common.h)
// so their arithmetic can be checked against the oracle on a machine without a GPU.  The product library never loads
// this file; it exists so that logic errors are found before GPU minutes are spent.
#include <omp.h>

#include <vector>

#include <mutex>
struct Backend {
  bool dirty = false;
  std::mutex mu;
};
#include "../../qpcontrol.jl_b200/csrc/setup_api.h"
#include "../../qpcontrol.jl_b200/csrc/admm.cuh"
#include "../../qpcontrol.jl_b200/csrc/kin.cuh"

using namespace qpc;

extern "C" {

int qpc_device_count(void) { return 0; }

int qpc_finalize(qpc_controller* c, int32_t) {
  if (!c) return qpc_fail(QPC_ERR_ARG, "null controller");
  std::string err = compile_program(c->hc, c->prog);
  if (!err.empty()) return qpc_fail(QPC_ERR_LIMIT, err);
  c->finalized = true;
  return QPC_OK;
}
void qpc_controller_destroy(qpc_controller* c) { delete c; }

static BatchIO make_io(const qpc_batch_in* in) {
  BatchIO io;
  io.q = in->q;
  io.v = in->v;
  io.desired = in->desired;
  io.cweight = in->contact_weight;
  io.cmaxnf = in->contact_maxnormalforce;
  io.desired_stride = in->desired_stride;
  io.contact_stride = in->contact_stride;
  io.tweight = in->task_weight;
  io.tweight_stride = in->task_weight_stride;
  io.cgeom = in->contact_geometry;
  io.cgeom_stride = in->contact_geometry_stride;
  io.twmat = in->task_weight_matrix;
  io.twmat_stride = in->task_weight_matrix_stride;
  io.time = in->time;
  io.time_stride = in->time_stride ? 1 : 0;
  return io;
}

int emu_assemble_batch(qpc_controller* c, int64_t B, const qpc_batch_in* in, double* P, double* qv, double* G,
                       double* lg, double* ug, double* lb, double* ub, double* des_out) {
  const DevProgram& p = c->prog;
  BatchIO io = make_io(in);
#pragma omp parallel
  {
    std::vector<double> smem(kin_smem_doubles(p.nb, p.nq, p.nv, p.ndes, p.ncontacts, p.N));
#pragma omp for
    for (int64_t i = 0; i < B; i++) {
      KinSmem s = kin_layout(smem.data(), p.nb, p.nq, p.nv, p.ndes, p.ncontacts, p.N);
      kin_load(&p, io, i, s);
      kin_forward(&p, s);
      kin_composite(&p, s);
      kin_standing(&p, s);
      kin_se3pd(&p, io, i, s);
      kin_contacts(&p, s);
      kin_assemble(&p, s, P + i * p.n * p.n, qv + i * p.n, G + i * p.mg * p.n, lg + i * p.mg, ug + i * p.mg,
                   lb + i * p.nbx, ub + i * p.nbx);
      if (des_out)
        for (int k = 0; k < p.ndes; k++) des_out[i * p.ndes + k] = s.des[k];
    }
  }
  return QPC_OK;
}

// rho_io != NULL: warm start -- x, y hold the previous solution on entry, rho_io the rho it ended with (<= 0: cold)
int emu_solve_qp_batch_warm(int64_t B, int32_t n, int32_t mg, int32_t nbox, const double* P, const double* qv,
                            const double* G, const double* lg, const double* ug, const double* lb, const double* ub,
                            const qpc_settings* st, double* x, double* y, int32_t* status, int32_t* iters, double* res,
                            double* rho_io);
int emu_solve_qp_batch(int64_t B, int32_t n, int32_t mg, int32_t nbox, const double* P, const double* qv,
                       const double* G, const double* lg, const double* ug, const double* lb, const double* ub,
                       const qpc_settings* st, double* x, double* y, int32_t* status, int32_t* iters, double* res) {
  return emu_solve_qp_batch_warm(B, n, mg, nbox, P, qv, G, lg, ug, lb, ub, st, x, y, status, iters, res, nullptr);
}
int emu_solve_qp_batch_warm(int64_t B, int32_t n, int32_t mg, int32_t nbox, const double* P, const double* qv,
                            const double* G, const double* lg, const double* ug, const double* lb, const double* ub,
                            const qpc_settings* st, double* x, double* y, int32_t* status, int32_t* iters, double* res,
                            double* rho_io) {
  Settings s;
  qpc_copy_settings(st, s);
#pragma omp parallel
  {
    std::vector<double> smem(admm_smem_doubles(n, mg, nbox));
#pragma omp for schedule(dynamic, 1)
    for (int64_t i = 0; i < B; i++) {
      AdmmProblem pb;
      pb.P = P + i * n * n;
      pb.qv = qv + i * n;
      pb.G = G + i * mg * n;
      pb.lg = lg + i * mg;
      pb.ug = ug + i * mg;
      pb.lb = lb + i * nbox;
      pb.ub = ub + i * nbox;
      pb.x = x + i * n;
      pb.y = y ? y + i * (mg + nbox) : nullptr;
      pb.status = status + i;
      pb.iters = iters ? iters + i : nullptr;
      pb.res = res ? res + 2 * i : nullptr;
      if (rho_io && y) {
        pb.x0 = pb.x;
        pb.y0 = pb.y;
        pb.rho_io = rho_io + i;
      }
      if (n == 0) {
        *pb.status = 1;
        if (pb.iters) *pb.iters = 0;
        continue;
      }
      admm_solve(s, pb, n, mg, nbox, smem.data());
    }
  }
  return QPC_OK;
}

int emu_solve_batch(qpc_controller* c, int64_t B, const qpc_batch_in* in, const qpc_batch_out* out) {
  const DevProgram& p = c->prog;
  BatchIO io = make_io(in);
#pragma omp parallel
  {
    std::vector<double> ksm(kin_smem_doubles(p.nb, p.nq, p.nv, p.ndes, p.ncontacts, p.N));
    std::vector<double> asm_(admm_smem_doubles(p.n, p.mg, p.nbx));
    std::vector<double> P(p.n * p.n + 1), qv(p.n + 1), G(p.mg * p.n + 1), lg(p.mg + 1), ug(p.mg + 1), lb(p.nbx + 1),
        ub(p.nbx + 1), x(p.n + 1), y(p.mg + p.nbx + 1), tau(p.nv), vd(p.nv), wr(p.ncontacts * 6 + 1);
#pragma omp for schedule(dynamic, 1)
    for (int64_t i = 0; i < B; i++) {
      KinSmem s = kin_layout(ksm.data(), p.nb, p.nq, p.nv, p.ndes, p.ncontacts, p.N);
      kin_load(&p, io, i, s);
      kin_forward(&p, s);
      kin_composite(&p, s);
      kin_standing(&p, s);
      kin_se3pd(&p, io, i, s);
      kin_contacts(&p, s);
      kin_assemble(&p, s, P.data(), qv.data(), G.data(), lg.data(), ug.data(), lb.data(), ub.data());
      int status = 1, iters = 0;
      double res[2] = {0, 0};
      if (p.n > 0) {
        AdmmProblem pb{P.data(), qv.data(), G.data(), lg.data(), ug.data(), lb.data(), ub.data(),
                       x.data(), y.data(), &status, &iters, res};
        admm_solve(p.settings, pb, p.n, p.mg, p.nbx, asm_.data());
      }
      kin_inverse_dynamics(&p, s, x.data(), vd.data(), wr.data(), tau.data());
      for (int k = 0; k < p.nv; k++) {
        if (out->tau) out->tau[i * p.nv + k] = tau[k];
        if (out->vdot) out->vdot[i * p.nv + k] = vd[k];
      }
      if (out->wrench)
        for (int k = 0; k < p.ncontacts * 6; k++) out->wrench[i * p.ncontacts * 6 + k] = wr[k];
      if (out->status) out->status[i] = status;
      if (out->iters) out->iters[i] = iters;
      if (out->residuals) {
        out->residuals[2 * i] = res[0];
        out->residuals[2 * i + 1] = res[1];
      }
    }
  }
  return QPC_OK;
}

// kinematics + the inverse-dynamics epilogue for given QP solutions x [B][n] (the stage after the ADMM kernel); lets a
// tick be assembled from emu_assemble_batch + any solver emulation (tests/emu/warp_emu.cpp) + this
int emu_id_batch(qpc_controller* c, int64_t B, const qpc_batch_in* in, const double* x, const qpc_batch_out* out) {
  const DevProgram& p = c->prog;
  BatchIO io = make_io(in);
#pragma omp parallel
  {
    std::vector<double> ksm(kin_smem_doubles(p.nb, p.nq, p.nv, p.ndes, p.ncontacts, p.N));
    std::vector<double> tau(p.nv), vd(p.nv), wr(p.ncontacts * 6 + 1);
#pragma omp for schedule(dynamic, 1)
    for (int64_t i = 0; i < B; i++) {
      KinSmem s = kin_layout(ksm.data(), p.nb, p.nq, p.nv, p.ndes, p.ncontacts, p.N);
      kin_load(&p, io, i, s);
      kin_forward(&p, s);
      kin_composite(&p, s);
      kin_standing(&p, s);
      kin_se3pd(&p, io, i, s);
      kin_contacts(&p, s);
      kin_inverse_dynamics(&p, s, x + i * p.n, vd.data(), wr.data(), tau.data());
      for (int k = 0; k < p.nv; k++) {
        if (out->tau) out->tau[i * p.nv + k] = tau[k];
        if (out->vdot) out->vdot[i * p.nv + k] = vd[k];
      }
      if (out->wrench)
        for (int k = 0; k < p.ncontacts * 6; k++) out->wrench[i * p.ncontacts * 6 + k] = wr[k];
    }
  }
  return QPC_OK;
}

// forward dynamics under the soft ground contact (kin.cuh: kin_forward_dynamics) for given applied torques
int emu_forward_dynamics_batch(qpc_controller* c, int64_t B, const double* q, const double* v, const double* tau,
                               double k, double d, double mu, double v_eps, double ground_z, double* vd_out,
                               double* fc_out) {
  const DevProgram& p = c->prog;
  BatchIO io;
  io.q = q;
  io.v = v;
  io.desired = nullptr;
  io.cweight = io.cmaxnf = nullptr;
  io.desired_stride = io.contact_stride = 0;
  ContactModel cm{k, d, mu, v_eps, ground_z};
#pragma omp parallel
  {
    std::vector<double> ksm(kin_smem_doubles(p.nb, p.nq, p.nv, p.ndes, p.ncontacts, p.N) + kin_fd_extra_doubles(p.nv));
#pragma omp for schedule(dynamic, 1)
    for (int64_t i = 0; i < B; i++) {
      KinSmem s = kin_layout(ksm.data(), p.nb, p.nq, p.nv, p.ndes, p.ncontacts, p.N);
      double* M = ksm.data() + kin_smem_doubles(p.nb, p.nq, p.nv, p.ndes, p.ncontacts, p.N);
      kin_load(&p, io, i, s);
      kin_forward(&p, s);
      kin_composite(&p, s);
      kin_forward_dynamics(&p, s, cm, tau + i * p.nv, M, M + p.nv * p.nv, M + p.nv * p.nv + p.nv, vd_out + i * p.nv,
                           fc_out ? fc_out + i * p.ncontacts * 3 : nullptr);
    }
  }
  return QPC_OK;
}

}  // extern "C"
