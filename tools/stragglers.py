#!/usr/bin/env python
"""Development aid: indices / iteration counts of the slowest solves of the 16,384-state workload on the device."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import qpc_loader
qpc = qpc_loader.load()
from qpcontrol_jl_b200 import OSQPSettings, scenarios
for name, st in (("notebook", OSQPSettings.standing_notebook()), ("test_suite", OSQPSettings.test_suite())):
    mech, low, ctrl, qnom = scenarios.atlas_standing(st)
    q, v = scenarios.atlas_random_states(mech, qnom, 16384, seed=3)
    res = ctrl(q, v, check=False)
    o = np.argsort(res.iters)[::-1][:16]
    print(name, "idx", o.tolist())
    print(name, "iters", res.iters[o].tolist(), "status", res.status[o].tolist(), "nfac", res.factorizations[o].tolist())
    print(name, "hist: >100:", int((res.iters > 100).sum()), ">200:", int((res.iters > 200).sum()), ">400:", int((res.iters > 400).sum()), ">1000:", int((res.iters > 1000).sum()),
          "sum iters", int(res.iters.sum()), "sum of iters above 400:", int(res.iters[res.iters > 400].sum()))
