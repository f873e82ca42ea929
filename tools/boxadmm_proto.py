"""Development tool (not product, not oracle): numpy prototype of the box-condensed multi-rate ADMM.

All general rows of the controller QP are equalities and all inequalities are boxes on the trailing nbx variables
(reference src/contacts.jl:64-65 are the only inequalities QPControl ever builds).  The OSQP iteration then has a fast
nonlinear part (the box rows) and a slow linear part (free variables' proximal centre, equality duals).  Here the slow
part is refreshed by a full OSQP iteration every `outer` iterations; in between only the nbx x nbx block
S = (H0 + sigma + rho cb^2)^-1 of the KKT inverse is applied.  Fixed points are OSQP's.
"""
import sys, os
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests")); sys.path.insert(0, os.path.join(ROOT, "tools"))
import numpy as np
from admm_proto import ruiz, admm, atlas_qps, stack, INF


def boxadmm(P, q, G, b, lb, ub, eps_abs=1e-5, eps_rel=1e-5, max_iter=5000, rho0=0.1, sigma=1e-6, alpha=1.6,
            outer=25, tol=5.0, scaling=10, delta=1e-2, adapt=True, alpha_slow=None, stats=None):
    n, mg, nb = P.shape[0], G.shape[0], lb.shape[0]
    nf = n - nb
    Eb = np.zeros((nb, n)); Eb[:, nf:] = np.eye(nb)
    A = np.vstack([G, Eb])
    Ps, qs, As, D, E, c = ruiz(P, q, A, scaling)
    Gs, cb = As[:mg], np.diag(As[mg:, nf:]).copy()
    bs = b * E[:mg]
    ls, us = np.maximum(lb, -INF) * E[mg:], np.minimum(ub, INF) * E[mg:]
    a_s = alpha if alpha_slow is None else alpha_slow
    # ---- setup: partial elimination of the L = (free x, equality rows) block ------------------------------------
    KLL = np.block([[Ps[:nf, :nf] + sigma * np.eye(nf), Gs[:, :nf].T], [Gs[:, :nf], -delta * np.eye(mg)]])
    KLB = np.vstack([Ps[:nf, nf:], Gs[:, nf:]])
    W = np.linalg.inv(KLL)
    T = W @ KLB
    H0 = Ps[nf:, nf:] - KLB.T @ T
    rho = rho0
    S = np.linalg.inv(H0 + np.diag(sigma + rho * cb * cb)); nfac = 1
    xf, yg = np.zeros(nf), np.zeros(mg)
    xB, z, y = np.zeros(nb), np.zeros(nb), np.zeros(nb)
    rL = np.concatenate([sigma * xf - qs[:nf], bs - delta * yg])
    gL = T.T @ rL
    for it in range(1, max_iter + 1):
        rB = sigma * xB - qs[nf:] + cb * (rho * z - y)
        xt = S @ (rB - gL)
        zt = cb * xt
        xB_prev, y_prev = xB, y
        xB = alpha * xt + (1 - alpha) * xB
        zr = alpha * zt + (1 - alpha) * z
        zn = np.clip(zr + y / rho, ls, us)
        y = y + rho * (zr - zn); z = zn
        if it % outer and it != max_iter:
            continue
        # ---- full OSQP iteration for the slow rows -----------------------------------------------------------------
        tL = W @ rL - T @ xt
        xf = a_s * tL[:nf] + (1 - a_s) * xf
        yg = yg + a_s * (tL[nf:] - yg)
        rL = np.concatenate([sigma * xf - qs[:nf], bs - delta * yg])
        gL = T.T @ rL
        # ---- OSQP residuals on the full problem ----------------------------------------------------------------------
        x = np.concatenate([xf, xB]); yy = np.concatenate([yg, y]); zz = np.concatenate([bs, z])
        Ax, Aty, Px = As @ x, As.T @ yy, Ps @ x
        rp = np.abs((Ax - zz) / E).max()
        rd = np.abs((Px + qs + Aty) / D / c).max()
        ps = max(np.abs(zz / E).max(), np.abs(Ax / E).max())
        ds = max(np.abs(Px / D).max(), np.abs(Aty / D).max(), np.abs(qs / D).max()) / c
        if stats is not None:
            stats.append((it, rp, rd, rho))
        if rp < eps_abs + eps_rel * ps and rd < eps_abs + eps_rel * ds:
            return D * x, E * yy / c, 1, it, nfac, rp, rd
        if adapt:
            prn = np.abs(Ax - zz).max() / (max(np.abs(zz).max(), np.abs(Ax).max()) + 1e-10)
            drn = np.abs(Px + qs + Aty).max() / (max(np.abs(Px).max(), np.abs(Aty).max(), np.abs(qs).max()) + 1e-10)
            rn = np.clip(rho * np.sqrt(prn / (drn + 1e-10)), 1e-6, 1e6)
            if rn > rho * tol or rn < rho / tol:
                rho = rn
                S = np.linalg.inv(H0 + np.diag(sigma + rho * cb * cb)); nfac += 1
    return D * x, E * yy / c, -2, max_iter, nfac, rp, rd


if __name__ == "__main__":
    B = int(sys.argv[1]) if len(sys.argv) > 1 else 32
    tight = len(sys.argv) > 2 and sys.argv[2] == "tight"
    kw0 = dict(eps_abs=1e-8, eps_rel=1e-16, max_iter=20000) if tight else {}
    mech, low, q, v, a = atlas_qps(B)
    ref = [admm(*stack(a, i), eps_abs=1e-10, eps_rel=1e-16, max_iter=40000) for i in range(B)]
    base = [admm(*stack(a, i), **kw0) for i in range(B)]
    def report(name, rs):
        its = np.array([r[3] for r in rs]); nf = np.array([r[4] for r in rs]); st = np.array([r[2] for r in rs])
        err = np.array([np.abs(r[0] - rr[0]).max() / max(1.0, np.abs(rr[0]).max()) for r, rr in zip(rs, ref)])
        print(f"{name:34s} iters mean {its.mean():7.1f} med {np.median(its):6.0f} max {its.max():6d} nfac {nf.mean():.2f} "
              f"ok {np.mean(st == 1):.3f} xerr med {np.median(err):.1e} max {err.max():.1e}")
    report("osqp (current device form)", base)
    for delta in (1e-2, 1e-4, 1e-6):
        for outer in (25, 10):
            for a_s in (1.6, 1.0):
                rs = [boxadmm(a["P"][i], a["q"][i], a["G"][i], a["lg"][i], a["lb"][i], a["ub"][i], delta=delta,
                              outer=outer, alpha_slow=a_s, **kw0) for i in range(B)]
                report(f"box delta={delta:g} outer={outer} a_slow={a_s}", rs)
