#!/usr/bin/env python
"""bench.py -- Atlas standing QP-control solves/sec on N B200 (BASELINE.json `metric`), one process per GPU.

    python bench.py --gpus 1 --steps K --warmup W                 # this framework (CUDA, sm_100a)
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P \
        bench.py --gpus N --steps K --warmup W                    # N ranks, instances sharded, no data-path collective
    python bench.py --impl reference ...                          # the CPU restatement of the reference on host cores

One "step" = one control tick (state -> kinematics/dynamics terms -> QP assembly -> ADMM solve -> inverse dynamics
-> torques) for a batch of `--batch` randomised Atlas states PER GPU (BASELINE config 3: 16,384 on one B200; the
OSQP settings are the Atlas notebook's, reference notebooks/Standing controller.ipynb:66-71).  Weak scaling: every
rank owns its own contiguous block of instances (seed offset by rank); nothing is exchanged on the data path
(torch.distributed is used only for the barrier and the max-over-ranks of the device time).

`value`     solves/s with the inputs resident in HBM, timed with CUDA events on the launching stream (per step, L2
            flushed between steps), max over ranks.
`e2e`       the same metric through the reference-facing call with HOST buffers (qpc_solve_batch, QPC_HOST_PTRS):
            pinned host q/v in, tau/vdot/wrenches/status/iters/residuals out, all transfers inside the timed region
            (inputs and the small outputs by chunked async copies; tau/vdot/wrenches are written into the pinned output
            buffers by the epilogue kernel itself -- the same bytes over PCIe, counted in d2h_bytes_per_step).
`roofline`  dominant kernel (the ADMM solve; the one-warp-per-QP kernel for the standing program): the FLOPs its
            algorithm executes (reduction + R inversions + K iterations, 2 per FMA) x batch / its
            CUDA-event duration, against the fp64 DFMA peak measured live on the same device.  SURVEY.md 8(d)'s
            W(n,m,K,R) at the canonical KKT dims is reported beside it (`survey_formula`): that formula prices a KKT
            iteration the kernel no longer performs, so it is context, not the fraction.
`config4_strong_split`  BASELINE config 4 (65,536 states with per-instance contact sets, ONE global batch cut into
            contiguous shards, one per rank): strong scaling, so the N-GPU efficiency measures the machine and not
            the per-rank seeds.
`cpu_baseline` / `--impl reference`: oracle/ (CPU fp64 restatement of RBD + Parametron + OSQP, lifted sparse form,
            OpenMP over instances) on the box's host cores -- "port", because Julia / OSQP.jl cannot run here.
"""
import argparse
import json
import os
import subprocess
import sys
import tempfile
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

METRIC = "atlas_standing_qp_control_solves_per_sec"
UNIT = "solves/s"


def _env_int(name, default):
    try:
        return int(os.environ.get(name, default))
    except ValueError:
        return default


def warp_kernel_flops(K, R, MG=24, NA=21, NB=32):
    """FLOPs (2 per FMA) the one-warp ADMM kernel executes per solve (csrc/admm_warp.cuh), all 32 lanes:
    Householder QR on rotating registers NA (4 MG + MG/32), back-substitution 2 NA (NA - 1), reduced Hessian NA (NB + 1);
    per factorisation the 32-pivot Gauss-Jordan NB (NB - 1), the rank-ME correction 2 ME NB and t0 / rho fold 2 NB; per
    iteration NB FMAs per lane; residual checks are lane-local."""
    ME = MG - NA
    setup = NA * (4 * MG + MG / 32.0) + 2 * NA * (NA - 1) + NA * (NB + 1)
    factor = NB * (NB - 1) + 2 * ME * NB + 2 * NB
    return 2.0 * 32.0 * (setup + R * factor + K * NB)


def algorithmic_flops(n, m, K, R):
    """W(n,m,K,R) = R [n(n+1) m + n^3/3] + K [4nm + 2n^2 + 12(n+m)] + ceil(K/25) [2nm + 2n^2]."""
    return R * (n * (n + 1) * m + n ** 3 / 3.0) + K * (4 * n * m + 2 * n * n + 12 * (n + m)) + \
        np.ceil(K / 25.0) * (2 * n * m + 2 * n * n)


class ClockSampler:
    """SM clock and throttle reasons sampled every ~3 ms through NVML (in-process thread) while the warm-up and the timed
    region run -- the timed region is tens of milliseconds, shorter than the start-up of an nvidia-smi process; nvidia-smi
    (-lms 20) is the fallback when pynvml is missing."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")
    REASONS = ((0x8, "hw_slowdown"), (0x40, "hw_thermal_slowdown"), (0x20, "sw_thermal_slowdown"), (0x4, "sw_power_cap"))

    def __init__(self, gpu_index):
        self.gpu = gpu_index
        self.proc = None
        self.path = None
        self.thread = None
        self.rows = []

    def _nvml_loop(self, nv, handle):
        while not self._stop:
            try:
                sm = nv.nvmlDeviceGetClockInfo(handle, nv.NVML_CLOCK_SM)
                try:
                    mask = nv.nvmlDeviceGetCurrentClocksEventReasons(handle)
                except Exception:
                    mask = nv.nvmlDeviceGetCurrentClocksThrottleReasons(handle)
                self.rows.append((float(sm), int(mask)))
            except Exception:
                pass
            time.sleep(0.003)

    def start(self):
        try:
            import threading
            import pynvml as nv
            nv.nvmlInit()
            # CUDA_VISIBLE_DEVICES may renumber the devices: go through the PCI bus id of the torch device
            import torch
            bus = getattr(torch.cuda.get_device_properties(self.gpu), "pci_bus_id", None)
            handle = None
            if bus is not None:
                for i in range(nv.nvmlDeviceGetCount()):
                    h = nv.nvmlDeviceGetHandleByIndex(i)
                    if nv.nvmlDeviceGetPciInfo(h).bus == bus:
                        handle = h
                        break
            if handle is None:
                handle = nv.nvmlDeviceGetHandleByIndex(self.gpu)
            self.smax = float(nv.nvmlDeviceGetMaxClockInfo(handle, nv.NVML_CLOCK_SM))
            self._stop = False
            self.thread = threading.Thread(target=self._nvml_loop, args=(nv, handle), daemon=True)
            self.thread.start()
            return
        except Exception:
            self.thread = None
        try:
            fd, self.path = tempfile.mkstemp(prefix="qpc_clocks_", suffix=".csv")
            os.close(fd)
            self.f = open(self.path, "w")
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.gpu), f"--query-gpu={self.Q}",
                                          "--format=csv,noheader,nounits", "-lms", "20"], stdout=self.f,
                                         stderr=subprocess.DEVNULL)
        except Exception:
            self.proc = None

    def stop(self):
        out = {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        if self.thread is not None:
            self._stop = True
            self.thread.join(timeout=2)
            if self.rows:
                reasons = sorted({nm for _, mask in self.rows for bit, nm in self.REASONS if mask & bit})
                out.update(sm_mhz=float(np.median([r[0] for r in self.rows])), sm_max_mhz=self.smax, reasons=reasons,
                           samples=len(self.rows), source="nvml")
            return out
        if self.proc is None:
            return out
        try:
            self.proc.terminate()
            self.proc.wait(timeout=5)
        except Exception:
            pass
        try:
            self.f.close()
            rows = [r.strip().split(",") for r in open(self.path) if r.strip()]
            os.unlink(self.path)
        except Exception:
            return out
        sm, smax, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in rows:
            if len(r) < 8:
                continue
            try:
                sm.append(float(r[1]))
                smax.append(float(r[2]))
            except ValueError:
                continue
            for k, nm in enumerate(names):
                if r[4 + k].strip().lower().startswith("active"):
                    reasons.add(nm)
        if sm:
            out.update(sm_mhz=float(np.median(sm)), sm_max_mhz=float(max(smax)), reasons=sorted(reasons),
                       samples=len(sm), source="nvidia-smi")
        return out


def host_cores():
    try:
        return max(1, len(os.sched_getaffinity(0)))
    except AttributeError:
        return max(1, os.cpu_count() or 1)


def build_workload(batch, rank, settings_name):
    import qpc_loader
    qpc = qpc_loader.load()
    settings = getattr(qpc.OSQPSettings, settings_name)()
    return qpc, settings


SETTINGS = {"notebook": ("standing_notebook", "eps_abs=eps_rel=1e-5 max_iter=5000 adaptive_rho_interval=25 cold start "
                         "(notebooks/Standing controller.ipynb:66-71)"),
            "test_suite": ("test_suite", "eps_abs=1e-8 eps_rel=1e-16 max_iter=20000 adaptive_rho_interval=25 cold start "
                           "(test/runtests.jl:35-43: the settings at which parity <= 1e-5 is defined and tested)")}


def run_reference(args, rank, world):
    """The reference arm: oracle/ on the host cores (all threads), on a bounded sample of the same workload."""
    if rank != 0:
        return 0
    import qpc_loader
    qpc = qpc_loader.load()
    from oracle import oracle as orc
    settings = getattr(qpc.OSQPSettings, SETTINGS[args.settings][0])()
    mech, low, ctrl, qnom = qpc.scenarios.atlas_standing(settings)
    sample = args.cpu_sample
    q, v = qpc.scenarios.atlas_random_states(mech, qnom, sample, seed=3)
    oc = orc.OracleController(low.program)
    oc.set_settings(settings, warm_start=0)
    cores = host_cores()  # torchrun exports OMP_NUM_THREADS=1: ask for every core this process may run on
    for _ in range(args.warmup):
        oc.reset()
        oc.solve_batch(q, v, nthreads=cores)
    secs = []
    for _ in range(args.steps):
        oc.reset()
        r = oc.solve_batch(q, v, nthreads=cores)
        secs.append(r["seconds"])
    total = float(np.sum(secs))
    value = sample * args.steps / total
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": 1e3 * total / args.steps, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": "atlas_standing_randomised_states (BASELINE config 3)", "osqp": SETTINGS[args.settings][1],
                   "batch_per_step": sample},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": "port",
                         "sample": f"{sample} instances per step (seed 3), lifted sparse-LDL OSQP restatement, "
                                   f"OpenMP over instances; Julia/QPControl.jl cannot run on this box"},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
        "iters_mean": float(r["iters"].mean()),
    }
    print(json.dumps(line), flush=True)
    return 0


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--batch", type=int, default=16384, help="instances per GPU per step")
    ap.add_argument("--cpu-sample", type=int, default=4096, help="instances per step of the CPU arm")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--masks", action="store_true", help="BASELINE config 4: per-instance active contact sets")
    ap.add_argument("--settings", default="notebook", choices=sorted(SETTINGS),
                    help="OSQP settings: the Atlas notebook's (headline) or the reference test suite's (parity tolerance)")
    ap.add_argument("--config4-batch", type=int, default=65536, help="global batch of the strong-split config 4 leg (0 = skip)")
    ap.add_argument("--as-rank", type=int, default=None, help="development: use the seeds rank R would use (1 GPU)")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else args.warmup

    import qpc_loader
    from importlib import import_module
    qpc_loader.load()
    sharding = import_module("qpcontrol_jl_b200.sharding")
    rank, world, local = sharding.env_rank_world()
    if args.impl == "reference":
        return run_reference(args, rank, world)

    import torch
    import torch.distributed as dist
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; the hot path has no CPU fallback (use --impl reference for the "
                         "CPU arm)")
    torch.cuda.set_device(local)
    if world > 1:
        sharding.init_process_group("nccl", local)
    qpc = qpc_loader.load()
    from qpcontrol_jl_b200 import _lib

    settings = getattr(qpc.OSQPSettings, SETTINGS[args.settings][0])()
    mech, low, ctrl, qnom = qpc.scenarios.atlas_standing(settings, device=local)
    B = args.batch
    # each rank owns its own block of instances: seed offset = rank (rank 0 = BASELINE config 3's seed 3)
    seed_rank = rank if args.as_rank is None else args.as_rank
    q, v = qpc.scenarios.atlas_random_states(mech, qnom, B, seed=3 + 1000 * seed_rank)
    cw = cm = None
    if args.masks:
        cm = qpc.scenarios.contact_masks(B, len(low.program.contacts), seed=4 + 1000 * rank)
        cw = np.full_like(cm, 1e-3)
    dev = low.finalize()
    dims = dev.dims
    nv, nc = dims["nv"], dims["ncontacts"]
    dev.reserve(B)
    dev.h.sync_defaults()

    cuda = torch.device("cuda", local)
    dq, dv = torch.from_numpy(q).to(cuda), torch.from_numpy(v).to(cuda)
    dcw = None if cw is None else torch.from_numpy(cw).to(cuda)
    dcm = None if cm is None else torch.from_numpy(cm).to(cuda)
    out = dict(tau=torch.empty(B, nv, dtype=torch.float64, device=cuda),
               vdot=torch.empty(B, nv, dtype=torch.float64, device=cuda),
               wrench=torch.empty(B, nc, 6, dtype=torch.float64, device=cuda),
               status=torch.empty(B, dtype=torch.int32, device=cuda),
               iters=torch.empty(B, dtype=torch.int32, device=cuda),
               residuals=torch.empty(B, 2, dtype=torch.float64, device=cuda),
               factorizations=torch.empty(B, dtype=torch.int32, device=cuda))
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=cuda)  # > 126 MB L2
    stream = torch.cuda.current_stream()

    def step_device():
        dev.solve_device(B, dq, dv, out, contact_weight=dcw, contact_maxnormalforce=dcm, stream=stream.cuda_stream)

    def barrier():
        sharding.barrier(cuda)

    # ---- device-resident throughput ("value") -----------------------------------------------------------------------
    sampler = ClockSampler(local)  # started before the warm-up: the timed region is tens of milliseconds
    sampler.start()
    for _ in range(args.warmup):
        flush.zero_()
        step_device()
    torch.cuda.synchronize()
    launches0 = dev.launch_count()
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
    barrier()
    t_wall0 = time.perf_counter()
    for k in range(args.steps):
        flush.zero_()  # L2 flush between timed steps, outside the per-step events
        ev[k][0].record(stream)
        step_device()
        ev[k][1].record(stream)
    barrier()
    wall = time.perf_counter() - t_wall0
    ms_steps = [a.elapsed_time(b) for a, b in ev]
    ms_total = float(np.sum(ms_steps))
    launches = dev.launch_count() - launches0

    # stage split + roofline of the dominant kernel (ADMM), CUDA events recorded by the library on the same stream
    dev.set_profiling(True)
    stage = np.zeros(3)
    nprof = min(args.steps, 5)
    for _ in range(nprof):
        flush.zero_()
        step_device()
        stage += np.array(dev.stage_times())
    stage /= nprof
    dev.set_profiling(False)
    clocks = sampler.stop()

    status = out["status"].cpu().numpy()
    iters = out["iters"].cpu().numpy().astype(np.float64)
    nfac = out["factorizations"].cpu().numpy().astype(np.float64)
    accepted = float(np.mean((status == 1) | (status == 2)))
    tau_first = out["tau"][:min(args.cpu_sample, B)].cpu().numpy()  # compared with the CPU arm below

    # ---- per-batch latency of smaller batches (BASELINE metric: "per-batch latency us"), device-resident, rank 0's view ---
    latency = {}
    for lb in (1, 64, 444, 4096):
        if lb > B:
            continue
        ms_l = []
        for _ in range(7):
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record(stream)
            dev.solve_device(lb, dq, dv, out, contact_weight=dcw, contact_maxnormalforce=dcm, stream=stream.cuda_stream)
            b.record(stream)
            torch.cuda.synchronize()
            ms_l.append(a.elapsed_time(b))
        latency[str(lb)] = 1e3 * float(np.median(ms_l[2:]))

    # ---- end to end through the host-pointer C-ABI call ("e2e") ---------------------------------------------------------
    hq = torch.from_numpy(q).pin_memory()
    hv = torch.from_numpy(v).pin_memory()
    hcw = None if cw is None else torch.from_numpy(cw).pin_memory()
    hcm = None if cm is None else torch.from_numpy(cm).pin_memory()
    hres = qpc.BatchResult(tau=torch.empty(B, nv, dtype=torch.float64).pin_memory().numpy(),
                           vdot=torch.empty(B, nv, dtype=torch.float64).pin_memory().numpy(),
                           wrenches=torch.empty(B, nc, 6, dtype=torch.float64).pin_memory().numpy(),
                           status=torch.empty(B, dtype=torch.int32).pin_memory().numpy(),
                           iters=torch.empty(B, dtype=torch.int32).pin_memory().numpy(),
                           residuals=torch.empty(B, 2, dtype=torch.float64).pin_memory().numpy())
    h2d = hq.numel() * 8 + hv.numel() * 8 + (0 if hcw is None else 2 * hcw.numel() * 8)
    d2h = sum(a.nbytes for a in (hres.tau, hres.vdot, hres.wrenches, hres.status, hres.iters, hres.residuals))

    def step_host():
        dev.solve_host_into(hq.numpy(), hv.numpy(), hres, None if hcw is None else hcw.numpy(),
                            None if hcm is None else hcm.numpy())

    for _ in range(args.warmup):
        step_host()
    barrier()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        step_host()  # synchronous: returns after the D2H copies completed
    torch.cuda.synchronize()
    e2e_s = time.perf_counter() - t0
    e2e_ok = float(np.mean((hres.status == 1) | (hres.status == 2)))

    # ---- sequential ticks: warm-started closed loop on the device (reported beside the headline) ----
    seq = None
    if not args.masks:
        nseq = 25
        sq = torch.from_numpy(np.tile(qnom, (B, 1))).to(cuda)  # the notebook's nominal stance, upper body displaced
        sp = low.program.standing
        qj = torch.tensor([mech.qoff[j] for j in sp.joints], device=cuda)
        g = torch.Generator(device=cuda).manual_seed(7 + rank)
        sq[:, qj] += 0.1 * torch.randn(B, len(sp.joints), dtype=torch.float64, device=cuda, generator=g)
        sv = torch.zeros(B, nv, dtype=torch.float64, device=cuda)
        dev.set_warm_start(True)
        dev.reset_warm_start()
        dev.step_device(B, sq, sv, 2e-3, 5, out, stream=stream.cuda_stream)  # first ticks are cold / settling
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        dev.step_device(B, sq, sv, 2e-3, nseq, out, stream=stream.cuda_stream)
        e1.record(stream)
        torch.cuda.synchronize()
        seq_ms = e0.elapsed_time(e1)
        dev.set_warm_start(False)
        seq_ok = out["status"].cpu().numpy()
        seq = {"ticks": nseq, "dt": 2e-3, "ms_per_tick": seq_ms / nseq, "solves_per_s": B * nseq / (seq_ms * 1e-3),
               "iters_mean_last_tick": float(out["iters"].cpu().numpy().mean()),
               "accepted_frac_last_tick": float(np.mean((seq_ok == 1) | (seq_ok == 2))),
               "note": "warm-started closed loop (qpc_step_batch, device pointers), per rank; not the headline metric"}

    # ---- BASELINE config 4, strong split: ONE global batch of per-instance contact sets, contiguous shard per rank -------
    c4 = None
    if args.config4_batch > 0 and not args.masks:
        G4 = args.config4_batch
        lo4, hi4 = sharding.shard_range(G4, rank, world)
        n4 = hi4 - lo4
        q4, v4 = qpc.scenarios.atlas_random_states(mech, qnom, G4, seed=4)  # every rank draws the same global batch
        cm4 = qpc.scenarios.contact_masks(G4, len(low.program.contacts), seed=4)
        dq4, dv4 = torch.from_numpy(q4[lo4:hi4].copy()).to(cuda), torch.from_numpy(v4[lo4:hi4].copy()).to(cuda)
        dcm4 = torch.from_numpy(cm4[lo4:hi4].copy()).to(cuda)
        dcw4 = torch.full_like(dcm4, 1e-3)
        out4 = dict(tau=torch.empty(n4, nv, dtype=torch.float64, device=cuda),
                    vdot=torch.empty(n4, nv, dtype=torch.float64, device=cuda),
                    wrench=torch.empty(n4, nc, 6, dtype=torch.float64, device=cuda),
                    status=torch.empty(n4, dtype=torch.int32, device=cuda),
                    iters=torch.empty(n4, dtype=torch.int32, device=cuda),
                    residuals=torch.empty(n4, 2, dtype=torch.float64, device=cuda),
                    factorizations=torch.empty(n4, dtype=torch.int32, device=cuda))
        dev.reserve(n4)

        def step4():
            dev.solve_device(n4, dq4, dv4, out4, contact_weight=dcw4, contact_maxnormalforce=dcm4, stream=stream.cuda_stream)

        for _ in range(3):
            flush.zero_()
            step4()
        n4steps = max(3, min(args.steps, 10))
        ev4 = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(n4steps)]
        barrier()
        for k in range(n4steps):
            flush.zero_()
            ev4[k][0].record(stream)
            step4()
            ev4[k][1].record(stream)
        barrier()
        ms4 = float(np.sum([a.elapsed_time(b) for a, b in ev4]))
        ms4_max = float(sharding.max_over_ranks([ms4], cuda)[0])
        st4 = out4["status"].cpu().numpy()
        acc4 = float(sharding.max_over_ranks([-float(np.mean((st4 == 1) | (st4 == 2)))], cuda)[0])
        c4 = {"workload": "atlas_standing_varying_contact_sets (BASELINE config 4), each contact on w.p. 0.75, seed 4",
              "global_batch": G4, "shard": "contiguous blocks (sharding.shard_range), no collective", "scaling": "strong",
              "steps": n4steps, "ms_per_step": ms4_max / n4steps, "value": G4 * n4steps / (ms4_max * 1e-3), "unit": UNIT,
              "accepted_frac_min_over_ranks": -acc4, "iters_mean_rank0": float(out4["iters"].cpu().numpy().mean())}
        del dq4, dv4, dcm4, dcw4, out4

    # ---- max over ranks -----------------------------------------------------------------------------------------------------
    ms_total_max, e2e_s_max, admm_ms = [float(x) for x in sharding.max_over_ranks([ms_total, e2e_s, stage[1]], cuda)]
    total_solves = B * world * args.steps
    value = total_solves / (ms_total_max * 1e-3)
    e2e_value = total_solves / e2e_s_max

    if rank == 0:
        K, R = float(iters.mean()), float(nfac.mean())
        peak_tf = _lib.measure_fp64_peak(local)
        n_c, m_c = 68, 71  # SURVEY.md 8(d) canonical condensed dims
        warp = dev.admm_warp()
        nel_x = 0 if warp else dev.admm_eliminated()  # KKT fast path: diagonal-cost free variables eliminated in the solver
        n_x, m_x = dims["n"] - nel_x, dims["mg"] + dims["nbox"]
        w_canon = float(algorithmic_flops(n_c, m_c, K, R))
        if warp:
            w_exec = float(warp_kernel_flops(K, R, MG=dims["mg"], NA=dims["n"] - dims["nbox"]))
            kernel_name = "qpc_admm_warp_kernel<24,21>"
            dims_solved = {"reduced_variables": dims["nbox"], "eliminated_by_qr": dims["n"] - dims["nbox"],
                           "equality_rows": dims["mg"], "box_rows": dims["nbox"]}
        else:
            w_exec = float(algorithmic_flops(n_x, m_x, K, R))
            kernel_name = "qpc_admm_reg_kernel"
            dims_solved = {"n": n_x, "m": m_x, "eliminated": nel_x}
        ach = w_exec * B / (admm_ms * 1e-3) / 1e12
        ach_canon = w_canon * B / (admm_ms * 1e-3) / 1e12
        peaks = {}
        try:
            peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        except Exception:
            pass
        hbm_peak = float(peaks.get("hbm_gbs", 6650.0))
        bytes_per_solve = 1632.0  # q,v,maxnormalforce in; tau,vdot,wrenches,status,iters,res out
        hbm_ach = bytes_per_solve * B / (float(np.mean(ms_steps)) * 1e-3) / 1e9
        traffic = None
        tf = os.path.join(ROOT, "profiles", "admm_warp_traffic.json" if warp else "admm_traffic.json")
        if os.path.exists(tf):
            try:
                traffic = json.load(open(tf)).get("dram_bytes_per_launch")
            except Exception:
                traffic = None
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms_total_max / args.steps, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": ("atlas_standing_varying_contact_sets (BASELINE config 4)" if args.masks else
                                    "atlas_standing_randomised_states (BASELINE config 3)"),
                       "batch_per_gpu": B, "global_batch": B * world, "parallelism": f"instance-shard x{world}, "
                       "no collective", "osqp": SETTINGS[args.settings][1],
                       "qp_dims_solved": dims_solved, "l2": "flushed between steps (256 MiB memset)",
                       "model": "atlas-topology 36-DoF humanoid, synthetic inertias (qpcontrol_jl_b200.mechanism.atlas_like)"},
            "per_batch_latency_us": 1e3 * ms_total_max / args.steps,
            "ms_per_step_each": [round(float(x), 3) for x in ms_steps],
            "latency_us_by_batch": latency,
            "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": int(d2h),
                    "ms_per_step": 1e3 * e2e_s_max / args.steps, "accepted_frac": e2e_ok},
            "gpu_launches": int(launches),
            "clocks": clocks,
            "roofline": {"bound": "fp64", "kernel": kernel_name, "achieved": ach, "peak": peak_tf,
                         "unit": "TFLOP/s", "frac": ach / peak_tf if peak_tf else None, "traffic": traffic,
                         "peak_source": "measured live: dependent-chain-free DFMA loop on this device "
                                        "(MEASURED_PEAKS.json carries no fp64 figure)",
                         "flops_per_solve": w_exec,
                         "flops_definition": ("FLOPs the kernel's algorithm executes (bench.py: warp_kernel_flops; DESIGN.md 2.3)"
                                              if warp else "SURVEY.md 8(d) W(n,m,K,R) at the dims the kernel iterates on"),
                         "survey_formula": {"dims": {"n": n_c, "m": m_c}, "flops_per_solve": w_canon, "achieved": ach_canon,
                                            "frac": ach_canon / peak_tf if peak_tf else None,
                                            "note": "SURVEY.md 8(d) W(68,71,K,R) with this kernel's K and R: prices a "
                                                    "KKT-system iteration; context only when the one-warp kernel runs"},
                         "iters_mean": K, "factorizations_mean": R,
                         "kernel_ms": admm_ms,
                         "hbm": {"achieved": hbm_ach, "peak": hbm_peak, "unit": "GB/s", "frac": hbm_ach / hbm_peak,
                                 "bytes_per_solve": bytes_per_solve}},
            "stage_ms": {"assemble": float(stage[0]), "admm": float(stage[1]), "inverse_dynamics": float(stage[2])},
            "accepted_frac": accepted, "iters_max": float(iters.max()), "wall_s_timed_region": wall,
            "sequential_ticks": seq,
            "config4_strong_split": c4,
        }
        if not args.no_cpu_baseline and world == 1:
            from oracle import oracle as orc
            ncores = host_cores()
            sample = min(args.cpu_sample, B)
            oc = orc.OracleController(low.program)
            oc.set_settings(settings, warm_start=0)
            kw = {}
            if cm is not None:
                kw = dict(cweight=cw[:sample], cmaxnf=cm[:sample])
            oc.solve_batch(q[:64], v[:64], nthreads=ncores)
            oc.reset()
            r = oc.solve_batch(q[:sample], v[:sample], nthreads=ncores, **kw)
            line["cpu_baseline"] = {"value": sample / r["seconds"], "unit": UNIT, "cores": ncores,
                                    "kind": "port", "sample": f"first {sample} instances of rank 0's batch, "
                                    "oracle/ lifted sparse-LDL OSQP restatement, OpenMP over instances, cold start",
                                    "iters_mean": float(r["iters"].mean())}
            tau_dev = tau_first[:sample]
            ok = (r["status"] == 1) & (status[:sample] == 1)
            err = np.abs(tau_dev[ok] - r["tau"][ok]).max(1) / np.maximum(1.0, np.abs(r["tau"][ok]).max(1))
            line["cpu_baseline"]["tau_rel_diff_median"] = float(np.median(err)) if err.size else None
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
