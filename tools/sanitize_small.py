"""Small workload for compute-sanitizer (memcheck / racecheck / synccheck): every kernel variant once."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import qpc_loader
qpc = qpc_loader.load()
from qpcontrol_jl_b200 import OSQPSettings, _lib, scenarios

st = OSQPSettings(eps_abs=1e-5, eps_rel=1e-5, max_iter=60)
mech, low, ctrl, qnom = scenarios.atlas_standing(st)
q, v = scenarios.atlas_random_states(mech, qnom, 4, seed=3)
r = ctrl(q, v, check=False)
low.set_warm_start(True)
q1, v1, r2 = ctrl.simulate(q, v, 2e-3, 2, check=False)
low.set_warm_start(False)
cg = np.array([list(c.position) + list(c.normal) + [c.mu] for c in low.program.contacts])
r3 = ctrl.lowlevel(q, v, check=False, task_weight=np.array([e.weight for e in low.program.tasks]), contact_geometry=cg)
print("atlas", r.status, r2.status, r3.status)
# register tiles; shared-memory kernel with 128- and 512-thread CTAs; per-CTA global scratch with 512-thread CTAs
for n, m in ((30, 30), (68, 71), (2, 3), (40, 110), (60, 100), (100, 100)):
    P, qv, A, l, u = scenarios.synthetic_qps(2, n, m, seed=5)
    out = _lib.solve_qp_batch_host(P, qv, A, l, u, settings=st)
    print("dense", n, m, out["status"], out["iters"])
mech2, low2, task = scenarios.acrobot_point_task(OSQPSettings(max_iter=60))
qa, va, da = scenarios.acrobot_random_inputs(mech2, 8, seed=2)
print("acrobot", low2(qa, va, da, check=False).status)
# round 2: the plant kernel, a program with a device-side SE3PDController (its own assembly instantiation), matrix weights
sys.path.insert(0, os.path.join(ROOT, "tests"))
qp, vp, rp = low.simulate_plant(q[:2], v[:2], 2e-3, 1, ground_z=-10.0, substeps=2, check=False)
print("plant", rp.status)
import test_se3pd_device as T3
mech3, low3, ctrl3, qnom3, se3, off3 = T3.make("piecewise", True, OSQPSettings(eps_abs=1e-5, eps_rel=1e-5, max_iter=60))
q3, v3 = scenarios.atlas_random_states(mech3, qnom3, 3, seed=4)
print("se3pd", low3(q3, v3, time=np.array([0.1, 0.7, 1.9]), check=False).status)
