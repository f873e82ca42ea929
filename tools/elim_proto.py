"""Development tool (not product, not oracle): numpy prototype of the OSQP iteration with the diagonal-cost free
variables eliminated from the KKT system (sigma = 0 on them): KKT [[P_kk + sigma + rho_b cb^2, G_k'],[G_k, -(1/rho + M)]],
M = G_f P_ff^-1 G_f'.  Compares iteration counts / accuracy with the full form on Atlas QPs."""
import sys, os
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests")); sys.path.insert(0, os.path.join(ROOT, "tools"))
import numpy as np
from admm_proto import admm, atlas_qps, stack, INF

lim = lambda v: np.minimum(np.where(v < 1e-4, 1.0, v), 1e4)

def elim_admm(P, q, G, lg, ug, lb, ub, nel, eps_abs=1e-5, eps_rel=1e-5, max_iter=5000, rho0=0.1, sigma=1e-6, alpha=1.6,
              interval=25, tol=5.0, scaling=10):
    n, mg, nb = P.shape[0], G.shape[0], lb.shape[0]
    nk = n - nel
    pf = np.diag(P)[:nel].copy(); Gf = G[:, :nel]; qf = q[:nel]
    M = (Gf / pf) @ Gf.T
    cg = (Gf / pf) @ qf
    Pk, qk, Gk = P[nel:, nel:].copy(), q[nel:].copy(), G[:, nel:].copy()
    # Ruiz on the FULL problem (x_f columns included), exactly OSQP's; the reduced matrices are then built from the scaled data
    from admm_proto import ruiz
    Ebx = np.zeros((nb, n)); Ebx[:, n - nb:] = np.eye(nb)
    Ps, qs, As, D, Efull, c = ruiz(P, q, np.vstack([G, Ebx]), scaling)
    Eg, Eb, Dk = Efull[:mg], Efull[mg:], D[nel:]
    pfs = np.diag(Ps)[:nel]; Gfs = As[:mg, :nel]
    M = (Gfs / pfs) @ Gfs.T
    Pk, qk, Gk = Ps[nel:, nel:], qs[nel:], As[:mg, nel:]
    cb = np.diag(As[mg:, n - nb:]).copy()
    cg = cg  # unscaled shift, scaled below through Eg
    A = np.vstack([Gk, np.hstack([np.zeros((nb, nk - nb)), np.diag(cb)])])
    E = np.concatenate([Eg, Eb])
    l = np.concatenate([(np.maximum(lg, -INF)) * Eg, np.maximum(lb, -INF) * Eb])
    u = np.concatenate([(np.minimum(ug, INF)) * Eg, np.minimum(ub, INF) * Eb])
    m = mg + nb
    eq = (u - l) < 1e-4
    rhovec = lambda r: np.where(eq, 1e3 * r, r)
    Mfull = np.zeros((m, m)); Mfull[:mg, :mg] = M
    rho = rho0; rv = rhovec(rho)
    def factor(rv):
        K = np.block([[Pk + sigma * np.eye(nk), A.T], [A, -np.diag(1 / rv) - Mfull]])
        return np.linalg.inv(K)
    Z = factor(rv); nfac = 1
    x, z, y = np.zeros(nk), np.zeros(m), np.zeros(m)
    shift = np.zeros(mg); w = np.zeros(mg); Df = D[:nel]; qfs = qs[:nel]
    for it in range(1, max_iter + 1):
        rhs = np.concatenate([sigma * x - qk, z - y / rv])
        t = Z @ rhs
        xt, nu = t[:nk], t[nk:]
        zt = z + (nu - y) / rv
        x = alpha * xt + (1 - alpha) * x
        w = alpha * nu[:mg] + (1 - alpha) * w
        zr = alpha * zt + (1 - alpha) * z
        zn = np.clip(zr + y / rv, l, u)
        y = y + rv * (zr - zn); z = zn
        if it % interval and it != max_iter:
            continue
        xfs = -(qfs + Gfs.T @ w) / pfs
        Ax = A @ x; implied = Ax[:mg] - (M @ w + (Gfs / pfs) @ qfs); Ax[:mg] += Gfs @ xfs
        if os.environ.get("CORR"):
            dcor = Ax[:mg] - implied            # what the iteration believes vs what is measured
            shift_new = dcor
            l[:mg] += shift - shift_new; u[:mg] += shift - shift_new; z[:mg] += shift - shift_new; shift = shift_new
            Ax[:mg] = implied + shift_new       # measured activity ...
            zc = z.copy(); zc[:mg] += shift_new  # ... compared with the unshifted target
        else:
            zc = z
        Aty = A.T @ y; Px = Pk @ x
        Pxf = -(qfs + Gfs.T @ w); Atyf = Gfs.T @ y[:mg]          # P_ff x_f and G_f' y_g with x_f = -P_ff^-1 (q_f + G_f' w)
        rp = np.abs((Ax - zc) / E).max()
        rd = max(np.abs((Px + qk + Aty) / Dk / c).max(), np.abs((Pxf + qfs + Atyf) / Df / c).max())
        ps = max(np.abs(z / E).max(), np.abs(Ax / E).max())
        ds = max(np.abs(Px / Dk).max(), np.abs(Aty / Dk).max(), np.abs(qk / Dk).max(),
                 np.abs(Pxf / Df).max(), np.abs(Atyf / Df).max(), np.abs(qfs / Df).max()) / c
        if rp < eps_abs + eps_rel * ps and rd < eps_abs + eps_rel * ds:
            break
        prn = np.abs(Ax - z).max() / (max(np.abs(z).max(), np.abs(Ax).max()) + 1e-10)
        drn = max(np.abs(Px + qk + Aty).max(), np.abs(Pxf + qfs + Atyf).max()) / (max(np.abs(Px).max(), np.abs(Aty).max(), np.abs(qk).max(), np.abs(Pxf).max(), np.abs(Atyf).max(), np.abs(qfs).max()) + 1e-10)
        rn = np.clip(rho * np.sqrt(prn / (drn + 1e-10)), 1e-6, 1e6)
        if rn > rho * tol or rn < rho / tol:
            rho = rn; rv = rhovec(rho); Z = factor(rv); nfac += 1
    yg = Eg * w / c
    xf = -(qf + Gf.T @ yg) / pf
    st = 1 if it < max_iter or (rp < eps_abs + eps_rel * ps and rd < eps_abs + eps_rel * ds) else -2
    return np.concatenate([xf, Dk * x]), None, st, it, nfac, rp, rd

if __name__ == "__main__":
    B = int(sys.argv[1]) if len(sys.argv) > 1 else 32
    tight = len(sys.argv) > 2 and sys.argv[2] == "tight"
    kw0 = dict(eps_abs=1e-8, eps_rel=1e-16, max_iter=20000) if tight else {}
    mech, low, q, v, a = atlas_qps(B)
    n, nb = a["P"].shape[1], a["lb"].shape[1]
    nel = n - nb
    ref = [admm(*stack(a, i), eps_abs=1e-10, eps_rel=1e-16, max_iter=40000) for i in range(B)]
    base = [admm(*stack(a, i), **kw0) for i in range(B)]
    def report(name, rs):
        its = np.array([r[3] for r in rs]); nf = np.array([r[4] for r in rs]); st = np.array([r[2] for r in rs])
        err = np.array([np.abs(r[0] - rr[0]).max() / max(1.0, np.abs(rr[0]).max()) for r, rr in zip(rs, ref)])
        print(f"{name:34s} iters mean {its.mean():7.1f} med {np.median(its):6.0f} max {its.max():6d} nfac {nf.mean():.2f} "
              f"ok {np.mean(st == 1):.3f} xerr med {np.median(err):.1e} max {err.max():.1e}")
    report("osqp full form (current)", base)
    rs = [elim_admm(a["P"][i], a["q"][i], a["G"][i], a["lg"][i], a["ug"][i], a["lb"][i], a["ub"][i], nel, **kw0) for i in range(B)]
    report(f"eliminated nel={nel}", rs)
