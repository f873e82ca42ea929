#!/usr/bin/env python
"""Where does the gathered-pivot-row inversion of the one-warp kernel (admm_warp.cuh: factor, `gather`) stop being accurate
enough?  Atlas standing ticks at eps_abs = eps_rel in {1e-5 .. 1e-8}; run once with QPC_WARP_GATHER_EPS=0 (always gather) and
once with QPC_WARP_GATHER_EPS=1 (never).   python tools/gather_eps_probe.py"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import qpc_loader
qpc = qpc_loader.load()
from qpcontrol_jl_b200 import OSQPSettings, scenarios

B = 16384
for eps in (1e-5, 1e-6, 1e-7, 1e-8):
    st = OSQPSettings(eps_abs=eps, eps_rel=eps, max_iter=5000, adaptive_rho_interval=25)
    mech, low, ctrl, qnom = scenarios.atlas_standing(st)
    q, v = scenarios.atlas_random_states(mech, qnom, B, seed=3)
    dev = low.finalize()
    dev.set_profiling(True)
    for _ in range(3):
        res = ctrl(q, v, check=False)
    ms = dev.stage_times()
    acc = np.mean((res.status == 1) | (res.status == 2))
    print(f"gather_eps={os.environ.get('QPC_WARP_GATHER_EPS', 'default')} eps={eps:g}: admm {ms[1]:.3f} ms, iters mean {res.iters.mean():.1f} "
          f"max {res.iters.max()}, accepted {acc:.5f}, status 1 frac {np.mean(res.status == 1):.5f}, residual max {res.residuals.max(0)}")
