"""Development tool (not product, not oracle): batched numpy version of tools/warp_proto.py for rho-adaptation studies."""
import sys, os
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests")); sys.path.insert(0, os.path.join(ROOT, "tools"))
import numpy as np
from admm_proto import admm, atlas_qps, stack
from warp_proto import reduce_qp


def reduce_all(a):
    B = a["P"].shape[0]; nb = a["lb"].shape[1]
    R = [reduce_qp(a["P"][i], a["q"][i], a["G"][i], a["lg"][i], nb) for i in range(B)]
    return {k: np.stack([r[k] for r in R]) for k in ("W", "xa0", "A3", "b3", "H", "h")}


def factor(H, h, A3, b3, rv):
    K = H + np.einsum("bi,ij->bij", rv, np.eye(H.shape[1]))
    Ki = np.linalg.inv(K)
    KA = Ki @ A3.transpose(0, 2, 1)
    S = A3 @ KA
    Si = np.linalg.inv(S)
    T = Ki - KA @ Si @ KA.transpose(0, 2, 1)
    t0 = -(T @ h[..., None])[..., 0] + (KA @ Si @ b3[..., None])[..., 0]
    return T, t0


def run(a, red, eps_abs=1e-5, eps_rel=1e-5, max_iter=5000, rho0=0.1, alpha=1.6, check=25, tol=5.0, eq_boost=1e3,
        adapt_at=None, rule="osqp", cs_mode="trace", rho_clip=(1e-6, 1e6), verbose=False):
    P, q, lb, ub, lg = a["P"], a["q"], a["lb"], a["ub"], a["lg"]
    H, h, A3, b3, W, xa0 = (red[k] for k in ("H", "h", "A3", "b3", "W", "xa0"))
    B, nb = lb.shape; na = P.shape[1] - nb
    eq = (ub - lb) < 1e-4
    cs = np.trace(H, axis1=1, axis2=2) / nb if cs_mode == "trace" else np.ones(B)
    rho = rho0 * cs
    rv = np.where(eq, eq_boost * rho[:, None], rho[:, None])
    T, t0 = factor(H, h, A3, b3, rv)
    nfac = np.ones(B, int); iters = np.zeros(B, int); status = np.zeros(B, int)
    z = np.zeros((B, nb)); y = np.zeros((B, nb)); xout = np.zeros((B, P.shape[1]))
    bn = np.abs(lg).max(1)
    act = np.ones(B, bool)
    if adapt_at is None:
        adapt_at = set(range(25, max_iter + 1, 25))
    for it in range(1, max_iter + 1):
        v = rv * z - y
        xt = (T @ v[..., None])[..., 0] + t0
        zr = alpha * xt + (1 - alpha) * z
        zn = np.clip(zr + y / rv, lb, ub)
        yn = y + rv * (zr - zn)
        rdv = yn - y - rv * (xt - z)
        z = np.where(act[:, None], zn, z); y = np.where(act[:, None], yn, y)
        if it % check and it != max_iter and it not in adapt_at:
            continue
        rp = np.abs(xt - z).max(1); rd = np.abs(rdv).max(1)
        xa = xa0 - (W @ xt[..., None])[..., 0]
        x = np.concatenate([xa, xt], 1)
        Px = (P @ x[..., None])[..., 0]
        Aty = -(Px + q); Aty[:, na:] += rdv
        ps = np.maximum(bn, np.maximum(np.abs(xt).max(1), np.abs(z).max(1)))
        ds = np.maximum(np.abs(Px).max(1), np.maximum(np.abs(Aty).max(1), np.abs(q).max(1)))
        if it % check == 0 or it == max_iter:
            ok = act & (rp < eps_abs + eps_rel * ps) & (rd < eps_abs + eps_rel * ds)
            xout[ok] = x[ok]; iters[ok] = it; status[ok] = 1; act &= ~ok
            if not act.any():
                break
        if it in adapt_at:
            if rule == "osqp":   # reduced-problem norms
                Hx = (H @ xt[..., None])[..., 0]
                prn = rp / (np.maximum(np.abs(xt).max(1), np.abs(z).max(1)) + 1e-10)
                drn = rd / (np.maximum(np.abs(Hx).max(1), np.maximum(np.abs(y).max(1), np.abs(h).max(1))) + 1e-10)
            elif rule == "full":  # full-problem norms (the termination normalisers)
                prn = rp / (ps + 1e-10); drn = rd / (ds + 1e-10)
            elif rule == "plain":  # plain residual balancing
                prn = rp; drn = rd
            rn = np.clip(rho * np.sqrt(prn / (drn + 1e-10)), rho_clip[0] * cs, rho_clip[1] * cs)
            ch = act & ((rn > rho * tol) | (rn < rho / tol))
            if ch.any():
                rho = np.where(ch, rn, rho)
                rv = np.where(eq, eq_boost * rho[:, None], rho[:, None])
                Tn, t0n = factor(H[ch], h[ch], A3[ch], b3[ch], rv[ch])
                T[ch] = Tn; t0[ch] = t0n; nfac[ch] += 1
    iters[act] = max_iter; status[act] = -2; xout[act] = x[act]
    return xout, status, iters, nfac, rho


if __name__ == "__main__":
    B = int(sys.argv[1]) if len(sys.argv) > 1 else 256
    mode = sys.argv[2] if len(sys.argv) > 2 else "loose"
    mech, low, q, v, a = atlas_qps(B)
    red = reduce_all(a)
    kw = dict(eps_abs=1e-5, eps_rel=1e-5, max_iter=5000) if mode == "loose" else dict(eps_abs=1e-8, eps_rel=1e-16, max_iter=20000)
    cache = f"/tmp/warp_ref_{B}.npy"
    if os.path.exists(cache):
        xr = np.load(cache)
    else:
        xr = np.stack([admm(*stack(a, i), eps_abs=1e-10, eps_rel=1e-16, max_iter=40000)[0] for i in range(B)]); np.save(cache, xr)
    nb = a["lb"].shape[1]; na = a["P"].shape[1] - nb
    def err(x):
        e1 = np.abs(x[:, :na] - xr[:, :na]).max(1) / np.maximum(1.0, np.abs(xr[:, :na]).max(1))
        w = (a["G"][:, :, na:] @ x[:, na:, None])[..., 0]; wr = (a["G"][:, :, na:] @ xr[:, na:, None])[..., 0]
        return np.maximum(e1, np.abs(w - wr).max(1) / np.maximum(1.0, np.abs(wr).max(1)))
    variants = [
        ("osqp-rule int25", dict()),
        ("full-rule int25", dict(rule="full")),
        ("plain-rule int25", dict(rule="plain")),
        ("osqp-rule tol2", dict(tol=2.0)),
        ("osqp-rule at 10,25,50,100,+100", dict(adapt_at=set([10, 25, 50, 100] + list(range(200, 20001, 100))))),
        ("full-rule at 10,25,50,100,+100", dict(rule="full", adapt_at=set([10, 25, 50, 100] + list(range(200, 20001, 100))))),
        ("osqp-rule rho0=.01", dict(rho0=0.01)),
        ("full-rule rho0=.01", dict(rule="full", rho0=0.01)),
    ]
    for name, kk in variants:
        x, st, its, nf, rho = run(a, red, **{**kw, **kk})
        e = err(x)
        print(f"{name:34s} iters mean {its.mean():7.1f} med {np.median(its):6.0f} p99 {np.percentile(its,99):6.0f} max {its.max():6d} nfac {nf.mean():.2f} ok {np.mean(st==1):.3f} err med {np.median(e):.2e} p99 {np.percentile(e,99):.2e} max {e.max():.2e}", flush=True)
