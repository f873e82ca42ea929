#!/usr/bin/env python
"""Measurements of the BASELINE.json configs that are not bench.py's headline line,
on one B200, device-resident inputs, CUDA events on the launching stream, L2 flushed between timed ticks.

    python tools/bench_configs.py [--out profiles/<name>.json] [--quick]

config 2: Acrobot PointAccelerationTask demo arm, 1,048,576 instances (hard point task, regularisation 1e-6).
config 4: Atlas standing controller, 65,536 instances with per-instance active contact sets (p = 0.75 per point,
          >= 3 enabled; disabled = maxnormalforce 0); infeasible instances stay in the batch and are counted.
config 5: synthetic dense QPs, n = m in {30, 68(71), 100, 143(178), 200}, batch sized to <= 2 GB of QP data; reports
          solves/s and the executed-flop rate W(n, m, K, R) / time against the measured DFMA peak.
Not part of the bench.py contract; results are recorded under profiles/ and quoted in DESIGN.md.
"""
import argparse
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import qpc_loader  # noqa: E402

qpc = qpc_loader.load()
from qpcontrol_jl_b200 import OSQPSettings, _lib, scenarios  # noqa: E402
import ctypes as C  # noqa: E402


def W(n, m, K, R):
    return R * (n * (n + 1) * m + n ** 3 / 3.0) + K * (4 * n * m + 2 * n * n + 12 * (n + m)) + \
        np.ceil(K / 25.0) * (2 * n * m + 2 * n * n)


def timed(fn, steps, warmup, torch, flush):
    for _ in range(warmup):
        flush.zero_()
        fn()
    torch.cuda.synchronize()
    ms = []
    for _ in range(steps):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        fn()
        e1.record()
        torch.cuda.synchronize()
        ms.append(e0.elapsed_time(e1))
    return float(np.median(ms))


def tick_config(torch, low, q, v, desired, cw, cm, steps, warmup, flush):
    cuda = torch.device("cuda", 0)
    dev = low.finalize()
    B = q.shape[0]
    nv, nc = dev.dims["nv"], dev.dims["ncontacts"]
    dev.reserve(B)
    dev.h.sync_defaults()
    t = lambda a: None if a is None else torch.from_numpy(np.ascontiguousarray(a)).to(cuda)  # noqa: E731
    dq, dv, dd, dcw, dcm = t(q), t(v), t(desired), t(cw), t(cm)
    out = dict(tau=torch.empty(B, nv, dtype=torch.float64, device=cuda),
               vdot=torch.empty(B, nv, dtype=torch.float64, device=cuda),
               wrench=torch.empty(B, max(nc, 1), 6, dtype=torch.float64, device=cuda),
               status=torch.empty(B, dtype=torch.int32, device=cuda), iters=torch.empty(B, dtype=torch.int32, device=cuda),
               residuals=torch.empty(B, 2, dtype=torch.float64, device=cuda),
               factorizations=torch.empty(B, dtype=torch.int32, device=cuda))
    stream = torch.cuda.current_stream().cuda_stream
    fn = lambda: dev.solve_device(B, dq, dv, out, desired=dd, contact_weight=dcw, contact_maxnormalforce=dcm,  # noqa: E731
                                  stream=stream)
    ms = timed(fn, steps, warmup, torch, flush)
    dev.set_profiling(True)  # stage split of one more tick (the library's CUDA events on the same stream)
    fn()
    stage = [float(t_) for t_ in dev.stage_times()]
    dev.set_profiling(False)
    st = out["status"].cpu().numpy()
    it = out["iters"].cpu().numpy()
    return dict(batch=B, ms_per_tick=ms, stage_ms=dict(assemble=stage[0], admm=stage[1], inverse_dynamics=stage[2]), solves_per_s=B / (ms * 1e-3), iters_mean=float(it.mean()),
                iters_max=int(it.max()), accepted_frac=float(np.mean((st == 1) | (st == 2))),
                status_counts={int(k): int(c) for k, c in zip(*np.unique(st, return_counts=True))},
                qp_dims=dict(n=dev.dims["n"], mg=dev.dims["mg"], nbox=dev.dims["nbox"]))


def dense_config(torch, n, m, B, steps, warmup, flush, peak_tf):
    cuda = torch.device("cuda", 0)
    nb = min(B, 2048)
    P, qv, A, l, u = scenarios.synthetic_qps(nb, n, m, seed=5)
    rep = (B + nb - 1) // nb
    tile = lambda a: torch.from_numpy(a).to(cuda).repeat((rep,) + (1,) * (a.ndim - 1))[:B].contiguous()  # noqa: E731
    dP, dq, dA, dl, du = (tile(a) for a in (P, qv, A, l, u))
    x = torch.empty(B, n, dtype=torch.float64, device=cuda)
    y = torch.empty(B, m, dtype=torch.float64, device=cuda)
    status = torch.empty(B, dtype=torch.int32, device=cuda)
    iters = torch.empty(B, dtype=torch.int32, device=cuda)
    res = torch.empty(B, 2, dtype=torch.float64, device=cuda)
    st = _lib.qpc_settings.from_py(OSQPSettings(eps_abs=1e-8, eps_rel=1e-8, max_iter=20000))
    lib = _lib.load()
    stream = torch.cuda.current_stream().cuda_stream
    p = lambda t_: C.c_void_p(t_.data_ptr())  # noqa: E731

    def fn():
        _lib.check(lib, lib.qpc_solve_qp_batch(C.c_int32(0), C.c_int64(B), C.c_int32(n), C.c_int32(m), C.c_int32(0), p(dP),
                                               p(dq), p(dA), p(dl), p(du), None, None, C.byref(st), p(x), p(y), p(status),
                                               p(iters), p(res), C.c_int32(_lib.DEVICE_PTRS), C.c_void_p(stream)),
                   "qpc_solve_qp_batch")
    ms = timed(fn, steps, warmup, torch, flush)
    it = iters.cpu().numpy().astype(np.float64)
    stt = status.cpu().numpy()
    K = float(it.mean())
    flops = float(W(n, m, K, 1.0 + K / 200.0))  # R is not returned by this entry point: ~1 refactorisation per 200 its
    return dict(n=n, m=m, batch=B, ms=ms, solves_per_s=B / (ms * 1e-3), iters_mean=K,
                solved_frac=float(np.mean(stt == 1)), tflops_W=flops * B / (ms * 1e-3) / 1e12,
                frac_of_dfma_peak=flops * B / (ms * 1e-3) / 1e12 / peak_tf, kernel="register tile" if n + m <= 144 else
                "shared-memory / global-scratch fallback")


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--out", default=None)
    ap.add_argument("--quick", action="store_true")
    ap.add_argument("--only-dense", action="store_true", help="config 5 only (development: A/B of the large-QP kernels)")
    ap.add_argument("--only-acrobot", action="store_true", help="config 2 only")
    ap.add_argument("--dense", default=None, help="n,m,B: one size of config 5 (with --only-dense)")
    args = ap.parse_args()
    import torch
    if not torch.cuda.is_available():
        raise SystemExit("needs a CUDA device")
    torch.cuda.set_device(0)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
    steps, warmup = (2, 1) if args.quick else (5, 3)
    result = {"device": torch.cuda.get_device_name(0), "fp64_dfma_peak_tflops": _lib.measure_fp64_peak(0)}

    if args.only_dense:
        sizes = [tuple(int(t) for t in args.dense.split(","))] if args.dense else [(100, 100, 4096), (143, 178, 1024),
                                                                                  (200, 200, 512)]
        for n, m, B in sizes:
            print("config 5", json.dumps(dense_config(torch, n, m, B, 1 if args.dense else 2, 1, flush, result["fp64_dfma_peak_tflops"])), flush=True)
        return 0

    if args.only_acrobot:
        mech, low, task = scenarios.acrobot_point_task()
        q, v, des = scenarios.acrobot_random_inputs(mech, 1 << 20, seed=2)
        print("config 2", json.dumps(tick_config(torch, low, q, v, des, None, None, 3, 2, flush)), flush=True)
        return 0

    # ---- config 1: single Atlas instance -- CPU oracle on one thread (cold / warm-started repeat calls, what the
    # notebook's @benchmark measures, Standing controller.ipynb:149) next to the device latency of a batch of one -----
    from oracle import oracle as orc
    st1 = OSQPSettings.standing_notebook()
    mech, low, ctrl, qnom = scenarios.atlas_standing(st1)
    q1 = np.tile(qnom, (64, 1))
    v1 = np.zeros((64, mech.nv))
    oc = orc.OracleController(low.program)
    oc.set_settings(st1, warm_start=0)
    oc.solve_batch(q1[:4], v1[:4], nthreads=1)
    cold = oc.solve_batch(q1, v1, nthreads=1)
    oc.set_settings(st1, warm_start=1)
    oc.reset()
    oc.solve_batch(q1[:1], v1[:1], nthreads=1)  # first call of the workspace: cold
    warm = oc.solve_batch(q1, v1, nthreads=1)   # 64 repeat calls on the same state, each warm-started from the last
    dev = low.finalize()
    import torch as _t
    r1 = tick_config(torch, low, q1[:1], v1[:1], None, None, None, 20, 5, flush)
    result["config1_atlas_single_instance"] = {
        "cpu_oracle_1_thread_cold_us_per_solve": 1e6 * cold["seconds"] / 64, "cpu_cold_iters": float(cold["iters"].mean()),
        "cpu_oracle_1_thread_warm_repeat_us_per_solve": 1e6 * warm["seconds"] / 64,
        "cpu_warm_iters": float(warm["iters"].mean()),
        "gpu_batch_of_one_latency_us": 1e3 * r1["ms_per_tick"], "gpu_iters": r1["iters_mean"],
        "note": "nominal notebook state, eps 1e-5; the CPU numbers are the oracle port (Julia/OSQP.jl cannot run here)"}
    print("config 1", json.dumps(result["config1_atlas_single_instance"]), flush=True)

    # ---- config 2 --------------------------------------------------------------------------------------------------
    mech, low, task = scenarios.acrobot_point_task()
    B2 = 65536 if args.quick else 1 << 20
    q, v, des = scenarios.acrobot_random_inputs(mech, B2, seed=2)
    r = tick_config(torch, low, q, v, des, None, None, steps, warmup, flush)
    r["hbm_bytes_per_solve"] = 72
    r["hbm_gbs_algorithmic"] = 72 * r["solves_per_s"] / 1e9
    result["config2_acrobot_point_task"] = r
    print("config 2", json.dumps(r), flush=True)

    # ---- config 4 --------------------------------------------------------------------------------------------------
    mech, low, ctrl, qnom = scenarios.atlas_standing(OSQPSettings.standing_notebook())
    B4 = 8192 if args.quick else 65536
    q, v = scenarios.atlas_random_states(mech, qnom, B4, seed=4)
    cm = scenarios.contact_masks(B4, len(low.program.contacts), seed=4)
    cw = np.full_like(cm, 1e-3)
    r = tick_config(torch, low, q, v, None, cw, cm, steps, warmup, flush)
    result["config4_atlas_contact_masks"] = r
    print("config 4", json.dumps(r), flush=True)

    # ---- config 5 --------------------------------------------------------------------------------------------------
    # n in {30, 50, 68, 100, 143, 200} x batch in {1k, 16k, 128k, 1M}, bounded to <= 24 GB of QP data
    sweep = []
    for n, m in ((30, 30), (50, 50), (68, 71), (100, 100), (143, 178), (200, 200)):
        for B in (1024, 16384, 131072, 1 << 20):
            if 8.0 * B * (n * n + m * n + 3 * n + 3 * m) <= 24e9 and (n + m <= 144 or B <= (16384 if n <= 100 else 1024)):
                sweep.append((n, m, B))
    if args.quick:
        sweep = [(30, 30, 2048), (68, 71, 2048)]
    result["config5_dense_qp_sweep"] = []
    for n, m, B in sweep:
        r = dense_config(torch, n, m, B, max(steps // 2, 1), 1, flush, result["fp64_dfma_peak_tflops"])
        result["config5_dense_qp_sweep"].append(r)
        print("config 5", json.dumps(r), flush=True)
    if args.out:
        with open(args.out, "w") as f:
            json.dump(result, f, indent=1)
    return 0


if __name__ == "__main__":
    sys.exit(main())
