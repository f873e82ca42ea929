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

    def solve(self, q, v, desired=None, cw=None, cm=None, task_weight=None, contact_geometry=None):
        h = self.h
        h.sync_defaults()
        q, v, desired, cw, cm, B = L._prep_host_inputs(h, q, v, desired, cw, cm)
        tw, cg = L._prep_tick_parameters(h, task_weight, contact_geometry)
        res = L._alloc_out(h, B)
        bi, bo = h.batch_in(q, v, desired, cw, cm, task_weight=tw, contact_geometry=cg), L._batch_out(res)
        L.check(self.lib, self.lib.emu_solve_batch(h.ctrl, C.c_int64(B), C.byref(bi), C.byref(bo)), "emu_solve_batch")
        return res

    def assemble(self, q, v, desired=None, cw=None, cm=None):
        h = self.h
        h.sync_defaults()
        q, v, desired, cw, cm, B = L._prep_host_inputs(h, q, v, desired, cw, cm)
        out = dict(P=np.zeros((B, h.n, h.n)), q=np.zeros((B, h.n)), G=np.zeros((B, h.mg, h.n)),
                   lg=np.zeros((B, h.mg)), ug=np.zeros((B, h.mg)), lb=np.zeros((B, h.nbox)), ub=np.zeros((B, h.nbox)),
                   desired=np.zeros((B, h.ndes)))
        bi = h.batch_in(q, v, desired, cw, cm)
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
