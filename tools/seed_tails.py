#!/usr/bin/env python
"""Development aid: ADMM stage time and iteration tail of the 16,384-state workload for the seeds bench.py gives ranks 0..7."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import qpc_loader
qpc = qpc_loader.load()
from qpcontrol_jl_b200 import OSQPSettings, scenarios
st = OSQPSettings.standing_notebook()
mech, low, ctrl, qnom = scenarios.atlas_standing(st)
dev = low.finalize(); dev.set_profiling(True)
for r in range(8):
    q, v = scenarios.atlas_random_states(mech, qnom, 16384, seed=3 + 1000 * r)
    for _ in range(3): res = ctrl(q, v, check=False)
    ms = dev.stage_times()
    it = np.sort(res.iters)[::-1]; top_idx = np.argsort(res.iters)[::-1][:3].tolist()
    print(f"rank {r}: admm {ms[1]:.3f} ms asm {ms[0]:.3f} iters mean {res.iters.mean():.1f} top {it[:6].tolist()} >400: {(res.iters>400).sum()} status!=1: {(res.status!=1).sum()} nfac {res.factorizations.mean():.2f} top_idx {top_idx}")
