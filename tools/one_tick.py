#!/usr/bin/env python
"""One (or a few) Atlas standing ticks on cuda:0, unchunked (profiling mode: one stream, events around the kernels) --
the command ncu wraps for the per-kernel captures under profiles/.   python tools/one_tick.py [notebook|test_suite] [B] [reps]"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import qpc_loader
qpc = qpc_loader.load()
from qpcontrol_jl_b200 import OSQPSettings, scenarios

name = sys.argv[1] if len(sys.argv) > 1 else "notebook"
B = int(sys.argv[2]) if len(sys.argv) > 2 else 16384
reps = int(sys.argv[3]) if len(sys.argv) > 3 else 2
st = OSQPSettings.standing_notebook() if name == "notebook" else OSQPSettings.test_suite()
mech, low, ctrl, qnom = scenarios.atlas_standing(st)
q, v = scenarios.atlas_random_states(mech, qnom, B, seed=3)
dev = low.finalize()
dev.set_profiling(True)
for _ in range(reps):
    res = ctrl(q, v, check=False)
ms = dev.stage_times()
print(f"{name} B={B} warp={dev.admm_warp()} stage ms asm/admm/id = {ms[0]:.3f} / {ms[1]:.3f} / {ms[2]:.3f}; iters mean {res.iters.mean():.1f} "
      f"max {res.iters.max()} nfac {res.factorizations.mean():.2f} accepted {np.mean((res.status == 1) | (res.status == 2)):.5f}")
it = np.sort(res.iters)[::-1]
print("largest iteration counts:", it[:12], "status of those:", res.status[np.argsort(res.iters)[::-1][:12]])
