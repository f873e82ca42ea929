"""Development tool (not product, not oracle): numpy prototype of the one-warp-per-QP ADMM (csrc/admm_warp.cuh).

The device QP  min 1/2 x'Px + q'x  s.t.  G x = b,  lb <= x_b <= ub  (x = (x_a, x_b), x_b the box variables = friction
cone multipliers) is reduced before iterating:
  * QR of G_a (mg x na, full column rank):  x_a = xa0 - W x_b, and the remaining me = mg - na rows A3 x_b = b3
  * reduced QP in x_b alone: H = P_bb + W'P_aa W, h = q_b - W'(P_aa xa0 + q_a)
  * ADMM with the box as the only split constraint, the equalities A3 handled exactly inside the x-update, sigma = 0:
    x~ = T (rho .* z - y) + t0,  T = K^-1 - K^-1 A3'(A3 K^-1 A3')^-1 A3 K^-1,  K = H + diag(rho)  -- a 32 x 32 operator
  * per-row rho: rows whose z sits on a bound get kappa rho, interior rows rho / kappa, l = u rows 1e3 rho (OSQP's
    rho_vec idea extended to the active set), re-evaluated on a geometric schedule together with OSQP's residual
    balancing rule.
Usage: warp_proto.py B loose|tight [masks]"""
import sys, os
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests")); sys.path.insert(0, os.path.join(ROOT, "tools"))
import numpy as np
from admm_proto import admm, stack
import qpc_loader
qpc = qpc_loader.load()


def reduce_qp(P, q, G, b, nb):
    n, mg = P.shape[0], G.shape[0]
    na = n - nb
    Ga, Gb = G[:, :na], G[:, na:]
    Q, R = np.linalg.qr(Ga, mode="complete")
    R1 = R[:na]
    Q1, Q2 = Q[:, :na], Q[:, na:]
    Ri = np.linalg.inv(R1)
    W = Ri @ (Q1.T @ Gb)
    xa0 = Ri @ (Q1.T @ b)
    A3, b3 = Q2.T @ Gb, Q2.T @ b
    Paa, Pbb, qa, qb = P[:na, :na], P[na:, na:], q[:na], q[na:]
    H = Pbb + W.T @ Paa @ W
    h = qb - W.T @ (Paa @ xa0 + qa)
    if A3.shape[0]:
        U, S, Vt = np.linalg.svd(A3, full_matrices=False)
        A3, b3 = Vt, (U.T @ b3) / S
    return dict(W=W, xa0=xa0, A3=A3, b3=b3, H=H, h=h, cond=np.linalg.cond(R1))


def solve(P, q, G, lg, lb, ub, eps_abs=1e-5, eps_rel=1e-5, max_iter=5000, rho0=0.1, alpha=1.6, check=5, tol=5.0,
          kappa=30.0, first=25, growth=1.35, eq_boost=1e3, eps_pinf=1e-4, aitken=25, cap=10**9, minstep=25):
    nb = lb.shape[0]; n = P.shape[0]; na = n - nb
    r = reduce_qp(P, q, G, lg, nb); H, h, A3, b3, W, xa0 = r["H"], r["h"], r["A3"], r["b3"], r["W"], r["xa0"]
    eq = (ub - lb) < 1e-4
    cs = np.trace(H) / nb; rho = rho0 * cs
    actv = np.zeros(nb, bool)
    def factor(rho, actv):
        rv = np.where(eq, eq_boost * rho, np.where(actv, rho * kappa, rho / kappa))
        K = H + np.diag(rv); Ki = np.linalg.inv(K); KA = Ki @ A3.T; Si = np.linalg.inv(A3 @ KA)
        T = Ki - KA @ Si @ KA.T; t0 = -T @ h + KA @ Si @ b3
        return rv, T, t0
    rv, T, t0 = factor(rho, actv); nfac = 1; nadapt = 0
    z = np.zeros(nb); y = np.zeros(nb); bn = np.abs(lg).max()
    nxt = first
    it = 0; sprev = None; dprev = None; njump = 0; rp_last = None
    while it < max_iter:
        it += 1
        v = rv * z - y; xt = T @ v + t0
        zr = alpha * xt + (1 - alpha) * z
        zn = np.clip(zr + y / rv, lb, ub); yn = y + rv * (zr - zn)
        rdv = yn - y - rv * (xt - z); dy = yn - y; zold = z; z, y = zn, yn
        adapt = it >= nxt
        if it % check and it != max_iter and not adapt: continue
        rp = np.abs(xt - z).max(); rd = np.abs(rdv).max()
        xa = xa0 - W @ xt; x = np.concatenate([xa, xt]); Px = P @ x
        Aty = -(Px + q); Aty[na:] += rdv
        ps = max(bn, np.abs(xt).max(), np.abs(z).max()); ds = max(np.abs(Px).max(), np.abs(Aty).max(), np.abs(q).max())
        pok = rp < eps_abs + eps_rel * ps
        if pok and rd < eps_abs + eps_rel * ds:
            return x, 1, it, nfac, nadapt, njump
        # Extrapolation of slowly converging solves (every `aitken` iterations, between adaptations): once the active set
        # has settled the iteration is affine, s+ = M s + c on s = (z, y); when one real mode dominates, successive
        # increments are parallel, d_k = r d_{k-1}, and the limit is s + d r / (1 - r) (Aitken).  r -> 1 is the dual drift
        # of a wrongly active row.  The step is cut so that no clipped row reaches its release point and no interior row
        # leaves the box: the active set, hence the affine regime, is preserved; ADMM then continues from the new point.
        if aitken and not adapt and it % aitken == 0 and it < max_iter:
            sv = np.concatenate([z, y])
            if sprev is not None:
                d = sv - sprev
                if dprev is not None:
                    dd = d @ d; pp = dprev @ dprev; dp = d @ dprev
                    if dd > 0 and pp > 0:
                        cosv = dp / np.sqrt(dd * pp); rr = dp / pp
                        if cosv > 1 - 1e-4 and 0 < rr < 1 - 1e-12:
                            gain = rr / (1 - rr)
                            dz = d[:nb]; dyv = d[nb:]
                            clipped = ((z <= lb) | (z >= ub)) & ~eq
                            m1 = clipped & (y * dyv < 0)
                            if m1.any(): gain = min(gain, 0.9 * np.min(-y[m1] / dyv[m1]))
                            inter = ~clipped & ~eq
                            m2 = inter & (dz > 0)
                            if m2.any(): gain = min(gain, 0.9 * np.min((ub[m2] - z[m2]) / dz[m2]))
                            m3 = inter & (dz < 0)
                            if m3.any(): gain = min(gain, 0.9 * np.min((lb[m3] - z[m3]) / dz[m3]))
                            if gain >= 1.0:
                                sn = sv + gain * d
                                z = np.clip(sn[:nb], lb, ub); y = sn[nb:]
                                njump += 1
                                sprev = None; dprev = None
                                continue
                dprev = d
            sprev = sv
        # primal infeasibility certificate of {A3 x = b3, lb <= x <= ub}: dy + A3'mu = 0, u'dy+ + l'dy- + b3'mu < 0
        ndy = np.abs(dy).max()
        if not pok and ndy > eps_pinf:
            mu = -A3 @ dy
            sup = ub @ np.maximum(dy, 0) + lb @ np.minimum(dy, 0) + b3 @ mu
            if sup < -eps_pinf * ndy and np.abs(dy + A3.T @ mu).max() < eps_pinf * ndy:
                return x, -3, it, nfac, nadapt, njump
        if adapt:
            nadapt += 1; nxt = max(min(int(np.ceil(nxt * growth)), nxt + cap), nxt + minstep)
            # rho floor: the explicit inverse T = (H + diag(rho))^-1 carries a rounding floor ~ eps_mach |x| lambda_max / rho_row
            # on the primal residual; keep it below eps_abs (lambda_max <= trace(H) = nb cs)
            rho_floor = kappa * 2.2e-16 * max(np.abs(xt).max(), np.abs(z).max(), 1.0) * nb * cs / eps_abs
            rn = np.clip(rho * np.sqrt((rp / (ps + 1e-300)) / (rd / (ds + 1e-300) + 1e-300)), max(1e-6 * cs, rho_floor), 1e6 * cs)  # no 1e-10 guards: they break scale invariance at tight tolerances
            na_ = (z <= lb) | (z >= ub)
            stalled = (not pok) and rp_last is not None and rp >= 0.5 * rp_last and rn > 1.5 * rho
            rp_last = rp
            big = rn > rho * tol or rn < rho / tol or stalled
            if big or (kappa != 1.0 and (na_ != actv).any()):
                if big: rho = rn
                actv = na_
                rv, T, t0 = factor(rho, actv); nfac += 1
                sprev = None; dprev = None
    return x, -2, max_iter, nfac, nadapt, njump


if __name__ == "__main__":
    B = int(sys.argv[1]) if len(sys.argv) > 1 else 64
    mode = sys.argv[2] if len(sys.argv) > 2 else "loose"
    masks = len(sys.argv) > 3 and sys.argv[3].startswith("masks")
    from emu import emu
    st = qpc.OSQPSettings.standing_notebook()
    mech, low, ctrl, qnom = qpc.scenarios.atlas_standing(st)
    q, v = qpc.scenarios.atlas_random_states(mech, qnom, B, seed=4 if masks else 3)
    cm = None
    if masks:
        p = float(sys.argv[3][5:] or 0.75)
        cm = qpc.scenarios.contact_masks(B, 8, p=p)
    a = emu.EmuController(low.program).assemble(q, v, cm=cm)
    kw = dict(eps_abs=1e-5, eps_rel=1e-5, max_iter=5000) if mode == "loose" else dict(eps_abs=1e-8, eps_rel=1e-16, max_iter=20000)
    cache = f"/tmp/warp_ref_{B}_{sys.argv[3] if masks else 'all'}.npy"
    if os.path.exists(cache):
        ref = np.load(cache)
    else:
        rr = [admm(*stack(a, i), eps_abs=1e-10, eps_rel=1e-16, max_iter=40000) for i in range(B)]
        ref = np.stack([np.concatenate([r[0], [r[2]]]) for r in rr]); np.save(cache, ref)
    xr, sr = ref[:, :-1], ref[:, -1]
    nb = a["lb"].shape[1]; na = a["P"].shape[1] - nb
    def err(x, i):
        e1 = np.abs(x[:na] - xr[i, :na]).max() / max(1.0, np.abs(xr[i, :na]).max())
        w = a["G"][i][:, na:] @ x[na:]; wr = a["G"][i][:, na:] @ xr[i, na:]
        return max(e1, np.abs(w - wr).max() / max(1.0, np.abs(wr).max()))
    print("reference statuses:", dict(zip(*np.unique(sr, return_counts=True))))
    base = [admm(*stack(a, i), **kw) for i in range(B)]
    print(f"{'osqp form':30s} iters mean {np.mean([r[3] for r in base]):7.1f} max {np.max([r[3] for r in base]):6d} nfac {np.mean([r[4] for r in base]):.2f}")
    variants = [(f"k{k:g} f{f} g{g:g}", dict(kappa=k, first=f, growth=g)) for k in (10.0, 30.0, 50.0) for f, g in ((25, 2.0), (20, 2.0), (10, 3.0))]
    for name, kk in variants:
        rs = [solve(a["P"][i], a["q"][i], a["G"][i], a["lg"][i], a["lb"][i], a["ub"][i], **{**kw, **kk}) for i in range(B)]
        its = np.array([r[2] for r in rs]); nf = np.array([r[3] for r in rs]); nad = np.array([r[4] for r in rs]); stt = np.array([r[1] for r in rs])
        okb = (stt == 1) & (sr == 1)
        es = np.array([err(r[0], i) for i, r in enumerate(rs)])[okb]
        cost = its + 42 * nf + 3 * nad
        print(f"{name:18s} cost {cost.mean():6.1f} iters mean {its.mean():6.1f} med {np.median(its):5.0f} p90 {np.percentile(its,90):5.0f} max {its.max():6d} nfac {nf.mean():.2f} "
              f"status {dict(zip(*np.unique(stt, return_counts=True)))} agree {np.mean((stt==1)==(sr==1)):.3f} err med {np.median(es):.1e} max {es.max():.1e}", flush=True)
