#!/usr/bin/env python
"""What the rho adaptations of the one-warp ADMM kernel do, per adaptation point: how many solves reach it, how many change
the global rho (a full re-inversion is unavoidable), how many only move rows between the active / interior rho classes (and
how many rows), how many change nothing.  Runs the kernel body on CPU fibres with the statistics hook of admm_warp.cuh
(tools/emu_stats/stats_emu.cpp).  Evidence for DESIGN.md 2.4.   python tools/adapt_stats.py [n]"""
import ctypes as C
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
import qpc_loader
qpc_loader.load()
from qpcontrol_jl_b200 import OSQPSettings, scenarios
from emu import emu

so = "/tmp/libqpc_stats_emu.so"
subprocess.run(["g++", "-O2", "-std=c++17", "-fPIC", "-shared", "-I" + os.path.join(ROOT, "include"), "-Wno-unknown-pragmas",
                "-o", so, os.path.join(ROOT, "tools", "emu_stats", "stats_emu.cpp")], check=True)
lib = C.CDLL(so)
n = int(sys.argv[1]) if len(sys.argv) > 1 else 512


def solve(a, st, N):
    B, nn = a["q"].shape
    mg, nbx = a["lg"].shape[1], a["lb"].shape[1]
    out = dict(x=np.zeros((B, nn)), y=np.zeros((B, mg + nbx)), rho=np.zeros(B), status=np.zeros(B, np.int32),
               iters=np.zeros(B, np.int32), res=np.zeros((B, 2)), nfac=np.zeros(B, np.int32), fallback=np.zeros(B, np.int32))
    p = lambda x: x.ctypes.data_as(C.c_void_p)  # noqa: E731
    P, q, G, lg, lb, ub = (np.ascontiguousarray(a[k]) for k in ("P", "q", "G", "lg", "lb", "ub"))
    rc = lib.emu_warp_solve_qp_batch(
        C.c_int64(B), C.c_int32(nn), C.c_int32(mg), C.c_int32(nbx), p(P), p(q), p(G), p(lg), p(lb), p(ub), C.c_double(st.rho),
        C.c_double(st.alpha), C.c_double(st.eps_abs), C.c_double(st.eps_rel), C.c_double(st.eps_prim_inf), C.c_int32(st.max_iter),
        C.c_int32(int(st.adaptive_rho)), C.c_double(st.adaptive_rho_tolerance), C.c_double(30.0), C.c_double(1.35), C.c_int32(25),
        C.c_int32(25), C.c_int32(25), C.c_int32(1 | (N << 8)), C.c_int32(0), p(out["x"]), p(out["y"]), p(out["rho"]),
        p(out["status"]), p(out["iters"]), p(out["res"]), p(out["nfac"]), p(out["fallback"]), None)
    assert rc == 0
    buf = (C.c_int * 3000000)()
    m = lib.emu_adapt_stats(buf, 3000000)
    return out, np.array(buf[:m]).reshape(-1, 3)


for name, st in (("notebook", OSQPSettings.standing_notebook()), ("test_suite", OSQPSettings.test_suite())):
    mech, low, ctrl, qnom = scenarios.atlas_standing(st)
    q, v = scenarios.atlas_random_states(mech, qnom, n, seed=3)
    a = emu.EmuController(low.program).assemble(q, v)
    out, stats = solve(a, st, low.program.N)
    r = out["rho"] / st.rho
    print(f"{name}: {n} solves, iterations mean {out['iters'].mean():.1f}, factorisations mean {out['nfac'].mean():.2f}; final / initial "
          f"global rho: unchanged on {np.mean(np.abs(r - 1) < 1e-9):.2f}, 5 % / 95 % quantiles {np.quantile(r, 0.05):.2f} / {np.quantile(r, 0.95):.2f}")
    for it in sorted(set(stats[:, 0]))[:6]:
        s = stats[stats[:, 0] == it]
        only = (s[:, 1] == 0) & (s[:, 2] > 0)
        print(f"  adaptation at {it}: reached by {len(s)}; global rho change {np.mean(s[:, 1]):.2f}; class changes only {np.mean(only):.2f} "
              f"(rows: mean {s[only, 2].mean() if only.any() else 0:.1f}, max {s[only, 2].max() if only.any() else 0}); nothing "
              f"{np.mean((s[:, 1] == 0) & (s[:, 2] == 0)):.2f}")
