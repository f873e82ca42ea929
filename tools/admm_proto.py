"""Development tool (not product, not oracle): numpy prototype of the device ADMM on the condensed Atlas QPs, used to
study iteration / factorisation counts of algorithmic variants before they are written as CUDA.  Needs tests/emu."""
import sys, os
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
import qpc_loader
qpc = qpc_loader.load()

INF = 1e20

def ruiz(P, q, A, iters=10):
    n, m = P.shape[0], A.shape[0]
    D, E, c = np.ones(n), np.ones(m), 1.0
    P, q, A = P.copy(), q.copy(), A.copy()
    lim = lambda v: np.minimum(np.where(v < 1e-4, 1.0, v), 1e4)
    for _ in range(iters):
        cn = np.maximum(np.abs(P).max(0), np.abs(A).max(0) if m else 0)
        rn = np.abs(A).max(1) if m else np.zeros(0)
        d, e = 1 / np.sqrt(lim(cn)), 1 / np.sqrt(lim(rn))
        P = P * d[:, None] * d[None, :]; A = A * e[:, None] * d[None, :]; q = q * d
        D *= d; E *= e
        ct = 1.0 / max(lim(np.abs(P).max(0).mean()), lim(np.abs(q).max()))
        P *= ct; q *= ct; c *= ct
    return P, q, A, D, E, c

def admm(P, q, A, l, u, eps_abs=1e-5, eps_rel=1e-5, max_iter=5000, rho0=0.1, sigma=1e-6, alpha=1.6, interval=25,
         tol=5.0, scaling=10, dual_mask=None, kform=False, adapt=True, eq_boost=1e3):
    n, m = P.shape[0], A.shape[0]
    Ps, qs, As, D, E, c = ruiz(P, q, A, scaling)
    ls, us = np.maximum(l, -INF) * E, np.minimum(u, INF) * E
    eq = (us - ls) < 1e-4
    def rhovec(r):
        return np.where(eq, eq_boost * r, r)
    rho = rho0; rv = rhovec(rho)
    def factor(rv):
        S = Ps + sigma * np.eye(n) + As.T @ (rv[:, None] * As)
        Sinv = np.linalg.inv(S)
        if kform:
            M1 = Sinv @ As.T
            return ("k", Sinv, M1, As @ M1, As @ Sinv)
        return ("s", np.linalg.cholesky(S))
    F = factor(rv); nfac = 1
    x, z, y = np.zeros(n), np.zeros(m), np.zeros(m)
    for it in range(1, max_iter + 1):
        w = rv * z - y
        if F[0] == "k":
            xt = F[1] @ (sigma * x - qs) + F[2] @ w
            zt = F[4] @ (sigma * x - qs) + F[3] @ w
        else:
            rhs = sigma * x - qs + As.T @ w
            L = F[1]
            xt = np.linalg.solve(L.T, np.linalg.solve(L, rhs))
            zt = As @ xt
        x = alpha * xt + (1 - alpha) * x
        zr = alpha * zt + (1 - alpha) * z
        zn = np.clip(zr + y / rv, ls, us)
        y = y + rv * (zr - zn); z = zn
        if it % interval and it != max_iter:
            continue
        Ax, Aty, Px = As @ x, As.T @ y, Ps @ x
        rp = np.abs((Ax - z) / E).max() if m else 0.0
        rdv = (Px + qs + Aty) / D / c
        rd = np.abs(rdv).max()
        ps = max(np.abs(z / E).max(), np.abs(Ax / E).max()) if m else 0
        if dual_mask is None:
            ds = max(np.abs(Px / D).max(), np.abs(Aty / D).max(), np.abs(qs / D).max()) / c
        else:
            ds = max(np.abs(Px / D)[dual_mask].max(), np.abs(Aty / D)[dual_mask].max(), np.abs(qs / D)[dual_mask].max()) / c
        if rp < eps_abs + eps_rel * ps and rd < eps_abs + eps_rel * ds:
            return D * x, E * y / c, 1, it, nfac, rp, rd
        if adapt:
            prn = np.abs(Ax - z).max() / (max(np.abs(z).max(), np.abs(Ax).max()) + 1e-10)
            drn = np.abs(Px + qs + Aty).max() / (max(np.abs(Px).max(), np.abs(Aty).max(), np.abs(qs).max()) + 1e-10)
            rn = np.clip(rho * np.sqrt(prn / (drn + 1e-10)), 1e-6, 1e6)
            if rn > rho * tol or rn < rho / tol:
                rho = rn; rv = rhovec(rho); F = factor(rv); nfac += 1
    return D * x, E * y / c, -2, max_iter, nfac, rp, rd

def atlas_qps(B, seed=3, settings=None):
    from emu import emu
    st = settings or qpc.OSQPSettings.standing_notebook()
    mech, low, ctrl, qnom = qpc.scenarios.atlas_standing(st)
    q, v = qpc.scenarios.atlas_random_states(mech, qnom, B, seed=seed)
    a = emu.EmuController(low.program).assemble(q, v)
    return mech, low, q, v, a

def stack(a, i):
    """condensed QP i as (P, q, A, l, u) with the box rows appended to A"""
    P, qv, G = a["P"][i], a["q"][i], a["G"][i]
    n, nb = P.shape[0], a["lb"].shape[1]
    Eb = np.zeros((nb, n)); Eb[:, n - nb:] = np.eye(nb)
    return P, qv, np.vstack([G, Eb]), np.concatenate([a["lg"][i], a["lb"][i]]), np.concatenate([a["ug"][i], a["ub"][i]])

if __name__ == "__main__":
    B = int(sys.argv[1]) if len(sys.argv) > 1 else 32
    mech, low, q, v, a = atlas_qps(B)
    for name, kw in [("base", {}), ("kform", dict(kform=True)), ("tight1e-8", dict(eps_abs=1e-8, eps_rel=1e-16, max_iter=20000)),
                     ("tight1e-8 kform", dict(eps_abs=1e-8, eps_rel=1e-16, max_iter=20000, kform=True))]:
        its, nf, st = [], [], []
        for i in range(B):
            r = admm(*stack(a, i), **kw)
            its.append(r[3]); nf.append(r[4]); st.append(r[2])
        print(f"{name:20s} iters mean {np.mean(its):7.1f} med {np.median(its):6.0f} max {np.max(its):6d}  nfac {np.mean(nf):.2f}  ok {np.mean(np.array(st)==1):.3f}")
