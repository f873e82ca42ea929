#!/usr/bin/env python
"""Development aid: dumps the reduced problem the one-warp ADMM kernel builds for instance 0 (QPC_WARP_DEBUG) and compares
it with the numpy prototype's reduction of the same assembled QP; prints hand-back reason codes (QPC_WARP_NOFALLBACK)."""
import os, sys
os.environ["QPC_WARP_DEBUG"] = "/tmp/warp_dbg.bin"
os.environ["QPC_WARP_NOFALLBACK"] = "1"
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests")); sys.path.insert(0, os.path.join(ROOT, "tools"))
import numpy as np
import qpc_loader
qpc = qpc_loader.load()
from qpcontrol_jl_b200 import OSQPSettings, scenarios
from warp_proto import reduce_qp, solve

np.set_printoptions(linewidth=220, precision=4)
st = OSQPSettings.standing_notebook() if len(sys.argv) < 2 or sys.argv[1] != "tight" else OSQPSettings.test_suite()
mech, low, ctrl, qnom = scenarios.atlas_standing(st)
B = 8
q, v = scenarios.atlas_random_states(mech, qnom, B, seed=3)
dev = low.finalize()
res = ctrl(q, v, check=False)
print("status", res.status, "iters", res.iters, "nfac", res.factorizations)
print("residuals", res.residuals)
a = dev.assemble_host(q, v)
ME, NA = 3, 21
d = np.fromfile("/tmp/warp_dbg.bin")
H = d[:1024].reshape(32, 32); h = d[1024:1056]; A3 = d[1056:1056 + ME * 32].reshape(ME, 32); b3 = d[1056 + ME * 32:1056 + ME * 33]
xa0 = d[1056 + ME * 33:1056 + ME * 33 + NA]; W = d[1056 + ME * 33 + NA:1056 + ME * 33 + NA + NA * 32].reshape(NA, 32); cs = d[1056 + ME * 33 + NA + NA * 32]
r = reduce_qp(a["P"][0], a["q"][0], a["G"][0], a["lg"][0], 32)
def rel(x, y): return np.abs(x - y).max() / max(np.abs(y).max(), 1e-300)
print("H rel", rel(H, r["H"]), "h rel", rel(h, r["h"]), "xa0 rel", rel(xa0, r["xa0"]), "W rel", rel(W, r["W"]), "cs", cs, np.trace(r["H"]) / 32)
# A3 rows span the same space up to an orthogonal transform: compare projectors and the particular solution
Pd, Pr = A3.T @ A3, r["A3"].T @ r["A3"]
print("A3 projector rel", rel(Pd, Pr), "A3'b3 rel", rel(A3.T @ b3, r["A3"].T @ r["b3"]), "A3 A3' - I", np.abs(A3 @ A3.T - np.eye(ME)).max())
print("H sym", np.abs(H - H.T).max(), "finite", np.isfinite(d).all())
x, stt, it, nf, nad, nj = solve(a["P"][0], a["q"][0], a["G"][0], a["lg"][0], a["lb"][0], a["ub"][0], eps_abs=st.eps_abs, eps_rel=st.eps_rel, max_iter=st.max_iter)
print("proto: status", stt, "iters", it, "nfac", nf)
