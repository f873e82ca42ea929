"""ORACLE -- TEST INFRASTRUCTURE ONLY.

ctypes wrapper around oracle/libqpc_oracle.so, the CPU fp64 restatement of the reference's control tick
(reference src/lowlevel/momentum.jl, src/tasks.jl, src/contacts.jl, src/highlevel/standing.jl + the RigidBodyDynamics /
Parametron / OSQP semantics of SURVEY.md appendix B).  Only tests/, __graft_entry__.smoke() and bench.py's
cpu_baseline / `--impl reference` legs may import this module.  PARITY UNPINNED: the reference has no golden vectors
and cannot run offline; see oracle/rbd.hpp.

The wrapper duck-types the product's host-side description objects (a `Mechanism` and a `Program`), so the oracle and
the CUDA path are always built from one and the same description.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess
from typing import Optional

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "libqpc_oracle.so")
_lib = None

_dp = C.POINTER(C.c_double)
_ip = C.POINTER(C.c_int32)


def build(force: bool = False) -> str:
    srcs = [os.path.join(_HERE, f) for f in ("capi.cpp", "controller.hpp", "osqp_port.hpp", "rbd.hpp", "Makefile")]
    stale = (not os.path.exists(_LIB_PATH)) or any(
        os.path.exists(s) and os.path.getmtime(s) > os.path.getmtime(_LIB_PATH) for s in srcs)
    if force or stale:
        subprocess.run(["make", "-C", _HERE, "-B" if force else "-s"], check=True, capture_output=True)
    return _LIB_PATH


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(_LIB_PATH):
            build()
        _lib = C.CDLL(_LIB_PATH)
        for name in ("orc_mechanism_create", "orc_state_create", "orc_controller_create"):
            getattr(_lib, name).restype = C.c_void_p
        _lib.orc_solve_batch.restype = C.c_double
        _lib.orc_solve_dense_qp_batch.restype = C.c_double
    return _lib


def _d(a):
    return None if a is None else a.ctypes.data_as(_dp)


def _i(a):
    return None if a is None else a.ctypes.data_as(_ip)


def _c(a, dtype=np.float64):
    return None if a is None else np.ascontiguousarray(a, dtype=dtype)


class OracleMechanism:
    def __init__(self, mech):
        self.mech = mech
        self._keep = [_c(mech.parent, np.int32), _c(mech.jtype, np.int32), _c(mech.axis), _c(mech.X_R), _c(mech.X_p),
                      _c(mech.mass), _c(mech.com), _c(mech.inertia_origin()), _c(mech.gravity)]
        k = self._keep
        self.h = C.c_void_p(lib().orc_mechanism_create(C.c_int(mech.nb), _i(k[0]), _i(k[1]), _d(k[2]), _d(k[3]),
                                                       _d(k[4]), _d(k[5]), _d(k[6]), _d(k[7]), _d(k[8])))

    def __del__(self):
        if getattr(self, "h", None):
            lib().orc_mechanism_destroy(self.h)
            self.h = None


class OracleState:
    """Single-state queries mirroring the RigidBodyDynamics calls of SURVEY.md 8(a) a15."""

    def __init__(self, omech: OracleMechanism):
        self.om = omech
        self.m = omech.mech
        self.h = C.c_void_p(lib().orc_state_create(omech.h))

    def __del__(self):
        if getattr(self, "h", None):
            lib().orc_state_destroy(self.h)
            self.h = None

    def set(self, q, v):
        q, v = _c(q), _c(v)
        assert q.shape == (self.m.nq,) and v.shape == (self.m.nv,)
        lib().orc_state_set(self.h, _d(q), _d(v))
        return self

    def _out(self, fn, n, *args):
        out = np.zeros(n)
        getattr(lib(), fn)(self.h, *args, _d(out))
        return out

    def center_of_mass(self):
        return self._out("orc_com", 3)

    def momentum(self):
        return self._out("orc_momentum", 6)

    def momentum_rate_bias(self):
        return self._out("orc_momentum_rate_bias", 6)

    def transform_to_root(self, body):
        R, p = np.zeros((3, 3)), np.zeros(3)
        lib().orc_transform_to_root(self.h, C.c_int(body), _d(R), _d(p))
        return R, p

    def twist_wrt_world(self, body):
        return self._out("orc_twist", 6, C.c_int(body))

    def bias_acceleration(self, body):
        return self._out("orc_bias_acceleration", 6, C.c_int(body))

    def geometric_jacobian(self, source, target, frame):
        return self._out("orc_geometric_jacobian", 6 * self.m.nv, C.c_int(source), C.c_int(target),
                         C.c_int(frame)).reshape(6, self.m.nv)

    def bias_in_frame(self, source, target, frame):
        return self._out("orc_bias_in_frame", 6, C.c_int(source), C.c_int(target), C.c_int(frame))

    def momentum_matrix(self, centroidal=False):
        return self._out("orc_momentum_matrix", 6 * self.m.nv, C.c_int(int(centroidal))).reshape(6, self.m.nv)

    def mass_matrix(self):
        return self._out("orc_mass_matrix", self.m.nv ** 2).reshape(self.m.nv, self.m.nv)

    def inverse_dynamics(self, vd, ext=None):
        vd = _c(vd)
        ext = _c(ext)
        tau = np.zeros(self.m.nv)
        lib().orc_inverse_dynamics(self.h, _d(vd), _d(ext), _d(tau))
        return tau


class OracleController:
    """The reference's `MomentumBasedController` (+ optional `StandingController`) in its lifted QP form."""

    def __init__(self, program):
        self.program = program
        self.m = program.mechanism
        self.om = OracleMechanism(self.m)
        L = lib()
        self.h = C.c_void_p(L.orc_controller_create(self.om.h, C.c_int(program.N), C.c_int(program.floating_body)))
        for kind, idx in program.events:
            if kind == "contact":
                c = program.contacts[idx]
                got = L.orc_add_contact(self.h, C.c_int(c.body), _d(_c(c.position)), _d(_c(c.normal)), C.c_double(c.mu))
                assert got == idx
                L.orc_set_contact_defaults(self.h, C.c_int(idx), C.c_double(c.weight), C.c_double(c.maxnormalforce))
            else:
                e = program.tasks[idx]
                t = e.task
                W = _c(e.W)
                got = L.orc_add_task(self.h, C.c_int(t.kind), C.c_int(t.source), C.c_int(t.target), C.c_int(t.frame),
                                     _d(_c(np.array(t.point))), C.c_int(t.joint), C.c_int(e.mode),
                                     C.c_double(e.weight), _d(W))
                assert got == idx
        # regularisation is stored per velocity index; push it joint by joint
        for j in range(self.m.nb):
            r = self.m.velocity_range(j)
            if len(r) and program.reg[r[0]] != 0.0:
                L.orc_regularize(self.h, C.c_int(j), C.c_double(float(program.reg[r[0]])))
        s = program.standing
        if s is not None:
            jt, jj = _c(s.joint_tasks, np.int32), _c(s.joints, np.int32)
            L.orc_set_standing(self.h, C.c_int(s.linmom_task), C.c_int(s.pelvis_task), C.c_int(s.pelvis_body),
                               C.c_int(len(s.joints)), _i(jt), _i(jj), _d(_c(s.joint_kp)), _d(_c(s.joint_kd)),
                               _d(_c(s.joint_ref)), C.c_double(s.com_kp), C.c_double(s.com_kd),
                               C.c_double(s.pelvis_kp), C.c_double(s.pelvis_kd), _d(_c(s.comref)))
        self.set_settings(program.settings)
        L.orc_finalize(self.h)
        dims = [C.c_int() for _ in range(4)]
        L.orc_dims(self.h, *[C.byref(d) for d in dims])
        self.nvar, self.nrows, self.ndes, self.ncontacts = [d.value for d in dims]

    def __del__(self):
        if getattr(self, "h", None):
            lib().orc_controller_destroy(self.h)
            self.h = None

    def set_settings(self, s, warm_start: Optional[int] = None):
        ws = s.warm_start if warm_start is None else warm_start
        lib().orc_set_settings(self.h, C.c_double(s.eps_abs), C.c_double(s.eps_rel), C.c_int(s.max_iter),
                               C.c_int(s.adaptive_rho_interval), C.c_int(s.check_termination), C.c_int(s.scaling),
                               C.c_int(ws), C.c_double(s.rho), C.c_double(s.sigma), C.c_double(s.alpha))

    def sync_contact_defaults(self):
        for idx, c in enumerate(self.program.contacts):
            lib().orc_set_contact_defaults(self.h, C.c_int(idx), C.c_double(c.weight), C.c_double(c.maxnormalforce))

    def task_rows(self, state: OracleState, task: int):
        dim = self.program.tasks[task].task.dimension
        J, b = np.zeros((dim, self.m.nv)), np.zeros(dim)
        lib().orc_task_rows(self.h, state.h, C.c_int(task), _d(J), _d(b))
        return J, b

    def standing_desireds(self, state: OracleState):
        des = np.zeros(self.ndes)
        lib().orc_standing_desireds(self.h, state.h, _d(des))
        return des

    def lifted_qp(self, q, v, desired=None, cweight=None, cmaxnf=None):
        q, v = _c(q), _c(v)
        desired = _c(self.program.default_desired() if desired is None else desired)
        cweight, cmaxnf = _c(cweight), _c(cmaxnf)
        n, m = self.nvar, self.nrows
        P, qq, A, l, u = np.zeros((n, n)), np.zeros(n), np.zeros((m, n)), np.zeros(m), np.zeros(m)
        lib().orc_lifted_qp(self.h, _d(q), _d(v), _d(desired), _d(cweight), _d(cmaxnf), _d(P), _d(qq), _d(A), _d(l),
                            _d(u))
        return P, qq, A, l, u

    def solve_batch(self, q, v, desired=None, cweight=None, cmaxnf=None, nthreads=0, return_lifted=False,
                    task_weight=None, contact_geometry=None):
        """Batched tick.  q [B,nq], v [B,nv]; desired [B,ndes] / [ndes] / None (task defaults); contact arrays
        [B,nc] / [nc] / None (ContactPoint values); task_weight [B,ntasks] per-tick scalar task weights and
        contact_geometry [B,nc,7] per-tick (position, normal, mu) -- the reference's Parameter-valued weights and contact
        frames (momentum.jl:107-110, contacts.jl:39,53-61)."""
        q, v = np.atleast_2d(_c(q)), np.atleast_2d(_c(v))
        B = q.shape[0]
        nv, nc = self.m.nv, self.ncontacts
        desired = self.program.default_desired() if desired is None else desired
        desired = _c(desired)
        dstride = 0 if desired.ndim == 1 else desired.shape[1]
        if cweight is None:
            cweight = np.array([c.weight for c in self.program.contacts])
        if cmaxnf is None:
            cmaxnf = np.array([c.maxnormalforce for c in self.program.contacts])
        cweight, cmaxnf = _c(cweight), _c(cmaxnf)
        if cweight.ndim != cmaxnf.ndim:
            if cweight.ndim == 1:
                cweight = np.ascontiguousarray(np.broadcast_to(cweight, cmaxnf.shape))
            else:
                cmaxnf = np.ascontiguousarray(np.broadcast_to(cmaxnf, cweight.shape))
        cstride = 0 if cweight.ndim == 1 else nc
        out = dict(tau=np.zeros((B, nv)), vd=np.zeros((B, nv)), wrenches=np.zeros((B, nc, 6)),
                   status=np.zeros(B, np.int32), iters=np.zeros(B, np.int32), res=np.zeros((B, 2)),
                   rho_updates=np.zeros(B, np.int32))
        xl = np.zeros((B, self.nvar)) if return_lifted else None
        tw = None if task_weight is None else _c(np.broadcast_to(task_weight, (B, len(self.program.tasks))))
        cg = None if contact_geometry is None else _c(np.broadcast_to(contact_geometry, (B, nc, 7)))
        lib().orc_set_tick_parameters(_d(tw), C.c_int64(0 if tw is None else tw.shape[1]), _d(cg),
                                      C.c_int64(0 if cg is None else nc * 7))
        secs = lib().orc_solve_batch(self.h, C.c_int64(B), _d(q), _d(v), _d(desired), C.c_int64(dstride), _d(cweight),
                                     _d(cmaxnf), C.c_int64(cstride), _d(out["tau"]), _d(out["vd"]),
                                     _d(out["wrenches"]), _i(out["status"]), _i(out["iters"]), _d(out["res"]),
                                     _i(out["rho_updates"]), _d(xl), C.c_int(nthreads))
        lib().orc_set_tick_parameters(None, C.c_int64(0), None, C.c_int64(0))
        out["seconds"] = secs
        if return_lifted:
            out["x_lifted"] = xl
        return out

    def reset(self):
        lib().orc_reset_workspaces(self.h)


def solve_dense_qp_batch(P, q, A, l, u, eps_abs=1e-8, eps_rel=1e-8, max_iter=20000, nthreads=0):
    """dense QPs min 1/2 x'Px + q'x, l <= Ax <= u; P [B,n,n], A [B,m,n]."""
    P, q, A, l, u = (_c(a) for a in (P, q, A, l, u))
    B, n = q.shape
    m = l.shape[1]
    x, y = np.zeros((B, n)), np.zeros((B, m))
    status, iters, res = np.zeros(B, np.int32), np.zeros(B, np.int32), np.zeros((B, 2))
    secs = lib().orc_solve_dense_qp_batch(C.c_int64(B), C.c_int(n), C.c_int(m), _d(P), _d(q), _d(A), _d(l), _d(u),
                                          C.c_double(eps_abs), C.c_double(eps_rel), C.c_int(max_iter), _d(x), _d(y),
                                          _i(status), _i(iters), _d(res), C.c_int(nthreads))
    return dict(x=x, y=y, status=status, iters=iters, res=res, seconds=secs)


def max_threads() -> int:
    return int(lib().orc_max_threads())
