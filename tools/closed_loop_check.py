#!/usr/bin/env python
"""Development aid: the closed loop with a plant (qpc_simulate_batch) for a few Atlas robots; prints CoM height, |v| and
accepted fraction over time."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
import qpc_loader
qpc = qpc_loader.load()
from qpcontrol_jl_b200 import OSQPSettings, scenarios, center_of_mass_host
import util

st = OSQPSettings.standing_notebook()
mech, low, ctrl, qnom = scenarios.atlas_standing(st)
B = 64
rng = np.random.default_rng(1)
q = np.tile(qnom, (B, 1)); v = np.zeros((B, mech.nv))
fj = mech.findjoint("pelvis_to_world"); o = int(mech.qoff[fj])
mask = np.ones(mech.nq, bool); mask[o:o + 7] = False
q[:, mask] += rng.normal(0, 0.02, (B, int(mask.sum())))
fk = util.forward_kinematics(mech, qnom)
zs = []
for c in low.program.contacts:
    R, p = fk[c.body]
    zs.append((R @ np.asarray(c.position) + p)[2])
k, d = float(sys.argv[1]) if len(sys.argv) > 1 else 5e4, float(sys.argv[2]) if len(sys.argv) > 2 else 1e3
pen = mech.total_mass * 9.81 / (len(zs) * k)
ground = min(zs) - 0.0 + pen * 0  # start with the soles on the ground (they sink by the static penetration)
print("contact z at nominal", np.round(zs, 4), "ground", ground, "static penetration", pen, "mass", mech.total_mass)
dev = low.finalize()
dev.set_warm_start(True)
dt, sub = 2e-3, int(sys.argv[3]) if len(sys.argv) > 3 else 8
T = 0.0
for block in range(10):
    q, v, res = dev.simulate_host(q, v, dt, 125, ground_z=ground, substeps=sub, stiffness=k, damping=d)
    T += 125 * dt
    com = np.array([center_of_mass_host(mech, q[i]) for i in range(B)])
    acc = np.mean((res.status == 1) | (res.status == 2))
    print(f"t={T:.2f}s com z mean {com[:,2].mean():.4f} min {com[:,2].min():.4f} |v|inf max {np.abs(v).max():.3e} pelvis z {q[:,o+6].mean():.4f} accepted {acc:.3f} iters {res.iters.mean():.1f} quat w min {np.abs(q[:,o]).min():.4f} argmax|v| {np.unravel_index(np.abs(v).argmax(), v.shape)} v0 top {np.round(np.sort(np.abs(v[0]))[::-1][:4],3)} idx {np.argsort(np.abs(v[0]))[::-1][:4]} |tau|max {np.abs(res.tau).max():.1f}", flush=True)
