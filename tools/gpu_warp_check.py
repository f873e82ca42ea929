#!/usr/bin/env python
"""GPU development check of the one-warp-per-QP ADMM kernel (csrc/admm_warp.cuh): against the register-tile kernel and
the oracle, statuses with contact masks, and stage timings.  Evidence / debugging aid; the pass/fail gates are in tests/.
   python tools/gpu_warp_check.py [--n 512] [--big 16384]"""
import argparse
import os
import sys
import time
import traceback

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import qpc_loader  # noqa: E402

qpc = qpc_loader.load()
from oracle import oracle as orc  # noqa: E402
from qpcontrol_jl_b200 import OSQPSettings, scenarios  # noqa: E402
import parity  # noqa: E402


def stats(name, res):
    st, cnt = np.unique(res.status, return_counts=True)
    print(f"  {name:10s} status {dict(zip(st.tolist(), cnt.tolist()))} iters mean {res.iters.mean():.1f} med {np.median(res.iters):.0f} "
          f"max {res.iters.max()} nfac {res.factorizations.mean():.2f} res max {res.residuals.max(0)}", flush=True)


def cmp(a, b, what="tau"):
    k = ((a.status == 1) | (a.status == 2)) & ((b.status == 1) | (b.status == 2))
    e = parity.rel_err(getattr(a, what)[k], getattr(b, what)[k])
    return f"{what} rel err med {np.median(e):.2e} p99 {np.percentile(e, 99):.2e} max {e.max():.2e}"


def section(title, fn):
    print(f"== {title}", flush=True)
    try:
        fn()
    except Exception:
        traceback.print_exc()
        sys.stdout.flush()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--n", type=int, default=512)
    ap.add_argument("--big", type=int, default=16384)
    args = ap.parse_args()
    n = args.n

    def ab(settings, masks=None, seed=3, with_oracle=False):
        mech, low, ctrl, qnom = scenarios.atlas_standing(settings)
        q, v = scenarios.atlas_random_states(mech, qnom, n, seed=seed)
        dev = low.finalize()
        cw = cm = None
        if masks is not None:
            cm = scenarios.contact_masks(n, 8, p=masks, seed=seed)
            cw = np.full((n, 8), 1e-3)
        print("  admm_warp:", dev.admm_warp())
        rw = ctrl(q, v, cw, cm, check=False)
        stats("warp", rw)
        dev.set_admm_warp(False)
        dev.set_admm_elimination(False)
        rr = ctrl(q, v, cw, cm, check=False)
        stats("reg(full)", rr)
        dev.set_admm_warp(True)
        print("  warp vs reg:", cmp(rw, rr, "tau"), "|", cmp(rw, rr, "vdot"), "|", cmp(rw, rr, "wrenches"))
        acc_w, acc_r = (rw.status == 1) | (rw.status == 2), (rr.status == 1) | (rr.status == 2)
        print(f"  accept agree {np.mean(acc_w == acc_r):.4f}; status pairs (warp, reg):",
              {tuple(p): int(c) for p, c in zip(*np.unique(np.stack([rw.status, rr.status], 1), axis=0, return_counts=True))})
        if with_oracle:
            oc = orc.OracleController(low.program)
            t0 = time.time()
            ref = oc.solve_batch(q, v, cweight=cw, cmaxnf=cm) if masks is not None else oc.solve_batch(q, v)
            print(f"  oracle {time.time() - t0:.1f}s status {dict(zip(*[x.tolist() for x in np.unique(ref['status'], return_counts=True)]))}")
            ok = ((ref["status"] == 1) | (ref["status"] == 2)) & acc_w
            for what, key in (("tau", "tau"), ("vdot", "vd"), ("wrenches", "wrenches")):
                e = parity.rel_err(getattr(rw, what)[ok], ref[key][ok])
                print(f"  warp vs oracle {what}: med {np.median(e):.2e} max {e.max():.2e}")
            print(f"  accept agree with oracle {np.mean(acc_w == ((ref['status'] == 1) | (ref['status'] == 2))):.4f}")
            a, b = parity.active_sets(rw.wrenches[ok], low.program), parity.active_sets(ref["wrenches"][ok], low.program)
            print(f"  active sets identical {np.mean(np.all(a == b, axis=1)):.4f}")

    section("notebook settings (eps 1e-5), config 3", lambda: ab(OSQPSettings.standing_notebook()))
    section("test-suite settings (eps 1e-8), config 3, vs oracle", lambda: ab(OSQPSettings.test_suite(), with_oracle=True))
    section("test-suite settings, contact masks p=0.75 (config 4), vs oracle",
            lambda: ab(OSQPSettings.test_suite(), masks=0.75, seed=4, with_oracle=True))
    section("notebook settings, contact masks p=0.4 (many infeasible)", lambda: ab(OSQPSettings.standing_notebook(), masks=0.4, seed=4))

    def timing():
        B = args.big
        for name, st in (("notebook", OSQPSettings.standing_notebook()), ("test_suite", OSQPSettings.test_suite())):
            mech, low, ctrl, qnom = scenarios.atlas_standing(st)
            q, v = scenarios.atlas_random_states(mech, qnom, B, seed=3)
            dev = low.finalize()
            dev.set_profiling(True)
            for warp in (True, False):
                dev.set_admm_warp(warp)
                for rep in range(3):
                    res = ctrl(q, v, check=False)
                ms = dev.stage_times()
                print(f"  {name:10s} warp={warp} B={B} stage ms asm/admm/id = {ms[0]:.3f} / {ms[1]:.3f} / {ms[2]:.3f}; iters mean "
                      f"{res.iters.mean():.1f} max {res.iters.max()} nfac {res.factorizations.mean():.2f} accepted "
                      f"{np.mean((res.status == 1) | (res.status == 2)):.5f} fallback-ish {np.sum(res.status == -99)}", flush=True)

    section("stage timings", timing)


if __name__ == "__main__":
    main()
