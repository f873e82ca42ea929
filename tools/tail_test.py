import os, sys
sys.path.insert(0, '/root/repo')
import numpy as np
import qpc_loader
qpc = qpc_loader.load()
from qpcontrol_jl_b200 import OSQPSettings, scenarios
for mi in (5000, 400, 100, 25):
    st = OSQPSettings.standing_notebook(); st.max_iter = mi
    mech, low, ctrl, qnom = scenarios.atlas_standing(st)
    q, v = scenarios.atlas_random_states(mech, qnom, 16384, seed=3)
    dev = low.finalize(); dev.set_profiling(True)
    for _ in range(3): res = ctrl(q, v, check=False)
    ms = dev.stage_times()
    print(f"max_iter {mi}: admm {ms[1]:.3f} ms iters mean {res.iters.mean():.1f} nfac {res.factorizations.mean():.2f}")
# uniform easy instance
st = OSQPSettings.standing_notebook()
mech, low, ctrl, qnom = scenarios.atlas_standing(st)
q, v = scenarios.atlas_random_states(mech, qnom, 16384, seed=3)
q[:] = q[1]; v[:] = v[1]
dev = low.finalize(); dev.set_profiling(True)
for _ in range(3): res = ctrl(q, v, check=False)
ms = dev.stage_times()
print(f"uniform instance: admm {ms[1]:.3f} ms iters mean {res.iters.mean():.1f} nfac {res.factorizations.mean():.2f}")
