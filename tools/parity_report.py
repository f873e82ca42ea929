#!/usr/bin/env python
"""Large-sample parity statistics of the CUDA path against the CPU oracle (test-suite OSQP settings, eps_abs 1e-8),
beyond what the unit tests sample.  Writes a JSON summary (profiles/) -- evidence, not a test: the pass/fail gates live
in tests/.   python tools/parity_report.py --out profiles/<name>.json [--n 4096]"""
import argparse
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import qpc_loader  # noqa: E402

qpc = qpc_loader.load()
from oracle import oracle as orc  # noqa: E402
from qpcontrol_jl_b200 import OSQPSettings, scenarios  # noqa: E402
import parity  # noqa: E402


def compare(res, ref, program):
    ok_ref = (ref["status"] == 1) | (ref["status"] == 2)
    ok_res = (res.status == 1) | (res.status == 2)
    k = ok_ref & ok_res
    et, ev = parity.rel_err(res.tau[k], ref["tau"][k]), parity.rel_err(res.vdot[k], ref["vd"][k])
    out = {"instances": int(len(ok_ref)), "accepted_by_oracle": int(ok_ref.sum()), "accepted_by_device": int(ok_res.sum()),
           "accept_reject_identical": bool(np.array_equal(ok_ref, ok_res)),
           "status_identical_frac": float(np.mean(res.status == ref["status"])),
           "tau_rel_err": {"max": float(et.max(initial=0)), "median": float(np.median(et)) if et.size else 0.0,
                           "frac_below_1e-5": float(np.mean(et < 1e-5)) if et.size else 1.0},
           "vdot_rel_err_max": float(ev.max(initial=0)),
           "floating_torques_exactly_zero": bool(np.all(res.tau[:, :6] == 0.0)) if program.floating_body >= 0 else None,
           "device_iters_mean": float(res.iters.mean()), "oracle_iters_mean": float(ref["iters"].mean())}
    pairs, counts = np.unique(np.stack([ref["status"], res.status], 1), axis=0, return_counts=True)
    out["status_pairs_oracle_device_count"] = [[int(a), int(b), int(c)] for (a, b), c in zip(pairs, counts)]
    dis = ok_ref != ok_res
    out["disagreeing"] = [{"oracle": int(ref["status"][i]), "device": int(res.status[i]), "oracle_iters": int(ref["iters"][i]),
                           "device_iters": int(res.iters[i]), "oracle_res": [float(x) for x in ref["res"][i]],
                           "device_res": [float(x) for x in res.residuals[i]]} for i in np.where(dis)[0][:12]]
    if res.wrenches.shape[1]:
        ew = parity.rel_err(res.wrenches[k], ref["wrenches"][k])
        out["wrench_rel_err_max"] = float(ew.max(initial=0))
        a, b = parity.active_sets(res.wrenches[k], program), parity.active_sets(ref["wrenches"][k], program)
        out["active_contact_sets_identical_frac"] = float(np.mean(np.all(a == b, axis=1)))
    return out


def bench_settings_accuracy(n):
    """bench.py's settings (notebook: eps 1e-5, max_iter 5000) are loose enough that two correct OSQP runs differ by
    ~1e-5..1e-4 in tau, so parity at those settings is stated against the converged solution (oracle at the test-suite
    tolerances): distance of (a) the device fast path (free variables eliminated), (b) the device full system, (c) the
    CPU oracle at the same loose settings from it.  The fast path must be no further away than OSQP itself is."""
    nb = OSQPSettings.standing_notebook()
    mech, low, ctrl, qnom = scenarios.atlas_standing(nb)
    q, v = scenarios.atlas_random_states(mech, qnom, n, seed=3)
    oc = orc.OracleController(low.program)
    oc.set_settings(OSQPSettings.test_suite(), warm_start=0)
    truth = oc.solve_batch(q, v)
    oc.set_settings(nb, warm_start=0)
    oc.reset()
    loose = oc.solve_batch(q, v)
    dev = low.finalize()
    fast = ctrl(q, v, check=False)          # the default solver: the one-warp kernel on the reduced problem
    nel = dev.admm_eliminated()
    dev.set_admm_elimination(False)
    dev.set_admm_warp(False)
    full = ctrl(q, v, check=False)          # the register-tile kernel on the full KKT system (round 1's formulation)
    dev.set_admm_elimination(True)
    dev.set_admm_warp(True)
    ok = (truth["status"] == 1) & ((fast.status == 1) | (fast.status == 2)) & ((full.status == 1) | (full.status == 2)) & \
        ((loose["status"] == 1) | (loose["status"] == 2))

    def stats(tau, wr):
        et, ew = parity.rel_err(tau[ok], truth["tau"][ok]), parity.rel_err(wr[ok], truth["wrenches"][ok])
        return {"tau_rel_err_median": float(np.median(et)), "tau_rel_err_p99": float(np.quantile(et, 0.99)),
                "tau_rel_err_max": float(et.max()), "wrench_rel_err_median": float(np.median(ew)),
                "wrench_rel_err_max": float(ew.max())}
    return {"instances": int(n), "compared": int(ok.sum()), "eliminated_variables": int(nel),
            "settings": "eps_abs = eps_rel = 1e-5, max_iter 5000 (Standing controller.ipynb:66-71), cold start",
            "reference": "oracle at eps_abs 1e-8 / eps_rel 1e-16 (test/runtests.jl:35-43)",
            "device_default_solver": dict(stats(fast.tau, fast.wrenches), iters_mean=float(fast.iters.mean()),
                                          one_warp_kernel=bool(dev.admm_warp())),
            "device_kkt_register_tile_kernel": dict(stats(full.tau, full.wrenches), iters_mean=float(full.iters.mean())),
            "cpu_oracle_same_settings": dict(stats(loose["tau"], loose["wrenches"]), iters_mean=float(loose["iters"].mean())),
            "accept_reject_fast_vs_full_identical": bool(np.array_equal((fast.status == 1) | (fast.status == 2),
                                                                        (full.status == 1) | (full.status == 2)))}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--out", default=None)
    ap.add_argument("--n", type=int, default=4096)
    args = ap.parse_args()
    n = args.n
    st = OSQPSettings.test_suite()
    report = {"settings": "test/runtests.jl:35-43 (eps_abs 1e-8, eps_rel 1e-16, max_iter 20000)", "tolerance": 1e-5}
    mech, low, ctrl, qnom = scenarios.atlas_standing(st)
    oc = orc.OracleController(low.program)
    oc.set_settings(st, warm_start=0)
    q, v = scenarios.atlas_random_states(mech, qnom, n, seed=3)
    report["config3_atlas_random_states"] = compare(ctrl(q, v, check=False), oc.solve_batch(q, v), low.program)
    print(json.dumps(report["config3_atlas_random_states"]), flush=True)
    q, v = scenarios.atlas_random_states(mech, qnom, n, seed=4)
    cm = scenarios.contact_masks(n, len(low.program.contacts), seed=4)
    cw = np.full_like(cm, 1e-3)
    oc.reset()
    report["config4_atlas_contact_masks"] = compare(ctrl(q, v, contact_weight=cw, contact_maxnormalforce=cm, check=False),
                                                    oc.solve_batch(q, v, cweight=cw, cmaxnf=cm), low.program)
    print(json.dumps(report["config4_atlas_contact_masks"]), flush=True)
    # harsher masks (p = 0.4, at least one contact): many infeasible instances -- the accept/reject decision must agree
    cm2 = scenarios.contact_masks(n, len(low.program.contacts), p=0.4, min_enabled=1, seed=14)
    oc.reset()
    report["atlas_sparse_contact_masks"] = compare(ctrl(q, v, contact_weight=cw, contact_maxnormalforce=cm2, check=False),
                                                   oc.solve_batch(q, v, cweight=cw, cmaxnf=cm2), low.program)
    print(json.dumps(report["atlas_sparse_contact_masks"]), flush=True)
    mech2, low2, task = scenarios.acrobot_point_task()
    qa, va, da = scenarios.acrobot_random_inputs(mech2, 4 * n, seed=2)
    report["config2_acrobot_point_task"] = compare(low2(qa, va, da, check=False),
                                                   orc.OracleController(low2.program).solve_batch(qa, va, desired=da),
                                                   low2.program)
    print(json.dumps(report["config2_acrobot_point_task"]), flush=True)
    report["bench_settings_vs_converged_solution"] = bench_settings_accuracy(min(n, 2048))
    print(json.dumps(report["bench_settings_vs_converged_solution"]), flush=True)
    if args.out:
        json.dump(report, open(args.out, "w"), indent=1)


if __name__ == "__main__":
    main()
