"""TEST INFRASTRUCTURE ONLY: drives tests/emu/libqpc_emu.so, the single-thread CPU compilation of the CUDA kernel
bodies (csrc/kin.cuh, csrc/admm.cuh), so their arithmetic is checked against the oracle without a GPU."""
import ctypes as C
import os
import subprocess

import numpy as np

import qpc_loader

qpc_loader.load()
from qpcontrol_jl_b200 import _lib as L  # noqa: E402

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "libqpc_emu.so")
_CSRC = os.path.join(_HERE, "..", "..", "qpcontrol.jl_b200", "csrc")


def build():
    deps = [os.path.join(_HERE, "emu.cpp")] + [os.path.join(_CSRC, f) for f in os.listdir(_CSRC)
                                                if f.endswith((".h", ".cuh"))]
    if os.path.exists(_SO) and all(os.path.getmtime(d) <= os.path.getmtime(_SO) for d in deps):
        return _SO
    subprocess.run(["/usr/bin/g++", "-O2", "-march=x86-64-v3", "-fopenmp", "-fPIC", "-std=c++17", "-shared", "-o", _SO,
                    os.path.join(_HERE, "emu.cpp")], check=True)
    return _SO


class EmuController:
    def __init__(self, program):
        self.lib = L.load(build())
        self.h = L.Handles(self.lib, program, 0)

    def solve(self, q, v, desired=None, cw=None, cm=None, task_weight=None, contact_geometry=None,
              task_weight_matrix=None, time=None):
        h = self.h
        h.sync_defaults()
        q, v, desired, cw, cm, B = L._prep_host_inputs(h, q, v, desired, cw, cm)
        twm = None
        if task_weight_matrix is None:
            tw, cg = L._prep_tick_parameters(h, task_weight, contact_geometry, B)
        else:
            tw, cg, twm = L._prep_tick_parameters(h, task_weight, contact_geometry, B, task_weight_matrix)
        res = L._alloc_out(h, B)
        bi, bo = h.batch_in(q, v, desired, cw, cm, task_weight=tw, contact_geometry=cg,
                            task_weight_matrix=twm, time=L._prep_time(time, B)), L._batch_out(res)
        L.check(self.lib, self.lib.emu_solve_batch(h.ctrl, C.c_int64(B), C.byref(bi), C.byref(bo)), "emu_solve_batch")
        return res

    def solve_warp(self, q, v, desired=None, cw=None, cm=None, **warp_kw):
        """One tick with the ONE-WARP ADMM kernel body (csrc/admm_warp.cuh on CPU fibres) between the emulated assembly
        and inverse-dynamics stages; `fallback` (reason codes) is attached to the result."""
        h = self.h
        a = self.assemble(q, v, desired, cw, cm)
        warp_kw.setdefault("pbb_block", h.program.N)  # the product passes the contact-block structure of P (api.cu: warp_pflags)
        w = warp_solve_qp_batch(a["P"], a["q"], a["G"], a["lg"], a["lb"], a["ub"], settings=h.program.settings, **warp_kw)
        h.sync_defaults()
        q, v, desired, cw, cm, B = L._prep_host_inputs(h, q, v, desired, cw, cm)
        res = L._alloc_out(h, B)
        bi, bo = h.batch_in(q, v, desired, cw, cm), L._batch_out(res)
        x = np.ascontiguousarray(w["x"])
        L.check(self.lib, self.lib.emu_id_batch(h.ctrl, C.c_int64(B), C.byref(bi), L._p(x), C.byref(bo)), "emu_id_batch")
        res.status[:], res.iters[:], res.residuals[:], res.factorizations[:] = w["status"], w["iters"], w["res"], w["nfac"]
        res.fallback = w["fallback"]
        return res

    def forward_dynamics(self, q, v, tau, k=5e4, d=1e3, mu=0.8, v_eps=1e-2, ground_z=-1e9):
        """vd = M^-1 (tau - c + J'w) with the soft ground contact of kin.cuh (ground far below by default: no contact)."""
        h = self.h
        q, v, tau = (np.atleast_2d(L._c(a)) for a in (q, v, tau))
        B = q.shape[0]
        vd = np.zeros((B, h.nv)); fc = np.zeros((B, max(h.ncontacts, 1), 3))
        L.check(self.lib, self.lib.emu_forward_dynamics_batch(h.ctrl, C.c_int64(B), L._p(q), L._p(v), L._p(tau),
                                                               C.c_double(k), C.c_double(d), C.c_double(mu),
                                                               C.c_double(v_eps), C.c_double(ground_z), L._p(vd), L._p(fc)),
                "emu_forward_dynamics_batch")
        return vd, fc[:, :h.ncontacts]

    def assemble(self, q, v, desired=None, cw=None, cm=None, time=None):
        h = self.h
        h.sync_defaults()
        q, v, desired, cw, cm, B = L._prep_host_inputs(h, q, v, desired, cw, cm)
        out = dict(P=np.zeros((B, h.n, h.n)), q=np.zeros((B, h.n)), G=np.zeros((B, h.mg, h.n)),
                   lg=np.zeros((B, h.mg)), ug=np.zeros((B, h.mg)), lb=np.zeros((B, h.nbox)), ub=np.zeros((B, h.nbox)),
                   desired=np.zeros((B, h.ndes)))
        bi = h.batch_in(q, v, desired, cw, cm, time=L._prep_time(time, B))
        p = L._p
        L.check(self.lib, self.lib.emu_assemble_batch(h.ctrl, C.c_int64(B), C.byref(bi), p(out["P"]), p(out["q"]),
                                                      p(out["G"]), p(out["lg"]), p(out["ug"]), p(out["lb"]),
                                                      p(out["ub"]), p(out["desired"])), "emu_assemble_batch")
        return out


def solve_qp_batch(P, qv, G, lg, ug, lb=None, ub=None, settings=None, warm=None):
    """`warm` = dict(x, y, rho) of a previous solve: OSQP-style warm start (updated in place)."""
    lib = L.load(build())
    P, qv, G, lg, ug = (L._c(a) for a in (P, qv, G, lg, ug))
    B, n = qv.shape
    mg = lg.shape[1]
    nbox = 0 if lb is None else lb.shape[1]
    lb = L._c(lb) if nbox else np.zeros((B, 0))
    ub = L._c(ub) if nbox else np.zeros((B, 0))
    from qpcontrol_jl_b200 import OSQPSettings
    st = L.qpc_settings.from_py(settings or OSQPSettings())
    out = dict(x=np.zeros((B, n)), y=np.zeros((B, mg + nbox)), status=np.zeros(B, np.int32),
               iters=np.zeros(B, np.int32), res=np.zeros((B, 2)))
    p = L._p
    out["rho"] = np.zeros(B)
    if warm is not None:
        out["x"][:], out["y"][:], out["rho"][:] = warm["x"], warm["y"], warm["rho"]
    L.check(lib, lib.emu_solve_qp_batch_warm(C.c_int64(B), C.c_int32(n), C.c_int32(mg), C.c_int32(nbox), p(P), p(qv),
                                             p(G), p(lg), p(ug), p(lb), p(ub), C.byref(st), p(out["x"]), p(out["y"]),
                                             p(out["status"]), p(out["iters"]), p(out["res"]), p(out["rho"])),
            "emu_solve_qp_batch_warm")
    return out


_WSO = os.path.join(_HERE, "libqpc_warp_emu.so")


def build_warp():
    deps = [os.path.join(_HERE, "warp_emu.cpp")] + [os.path.join(_CSRC, f) for f in ("admm_warp.cuh", "admm.cuh",
                                                                                      "qpc_common.h", "qpc_program.h")]
    if os.path.exists(_WSO) and all(os.path.getmtime(d) <= os.path.getmtime(_WSO) for d in deps):
        return _WSO
    subprocess.run(["/usr/bin/g++", "-O2", "-march=x86-64-v3", "-ffp-contract=off", "-fPIC", "-std=c++17", "-shared",
                    "-Wno-unknown-pragmas", "-o", _WSO, os.path.join(_HERE, "warp_emu.cpp")], check=True)
    return _WSO


def warp_solve_qp_batch(P, qv, G, lg, lb, ub, settings=None, kappa=30.0, growth=1.35, first=25, check=25, aitken=25, paa_diag=True, pbb_block=0,
                        warm=None, debug=False):
    """The one-warp ADMM kernel body (csrc/admm_warp.cuh, MG = 24, NA = 21) run on 32 CPU fibres per QP."""
    lib = C.CDLL(build_warp())
    from qpcontrol_jl_b200 import OSQPSettings
    st = settings or OSQPSettings()
    P, qv, G, lg, lb, ub = (L._c(a) for a in (P, qv, G, lg, lb, ub))
    B, n = qv.shape
    mg, nbx = lg.shape[1], lb.shape[1]
    out = dict(x=np.zeros((B, n)), y=np.zeros((B, mg + nbx)), rho=np.zeros(B), status=np.zeros(B, np.int32),
               iters=np.zeros(B, np.int32), res=np.zeros((B, 2)), nfac=np.zeros(B, np.int32),
               fallback=np.zeros(B, np.int32), dbg=np.zeros(4096))
    if warm is not None:
        out["x"][:], out["y"][:], out["rho"][:] = warm["x"], warm["y"], warm["rho"]
    p = L._p
    rc = lib.emu_warp_solve_qp_batch(C.c_int64(B), C.c_int32(n), C.c_int32(mg), C.c_int32(nbx), p(P), p(qv), p(G), p(lg),
                                     p(lb), p(ub), C.c_double(st.rho), C.c_double(st.alpha), C.c_double(st.eps_abs),
                                     C.c_double(st.eps_rel), C.c_double(st.eps_prim_inf), C.c_int32(st.max_iter),
                                     C.c_int32(int(st.adaptive_rho)), C.c_double(st.adaptive_rho_tolerance),
                                     C.c_double(kappa), C.c_double(growth), C.c_int32(first), C.c_int32(check), C.c_int32(aitken),
                                     C.c_int32(int(paa_diag) | (int(pbb_block) << 8)), C.c_int32(0 if warm is None else 1), p(out["x"]),
                                     p(out["y"]), p(out["rho"]), p(out["status"]), p(out["iters"]), p(out["res"]),
                                     p(out["nfac"]), p(out["fallback"]), p(out["dbg"]) if debug else None)
    if rc != 0:
        raise RuntimeError("emu_warp_solve_qp_batch: unsupported shape")
    return out
