"""ctypes binding of libqpcontrol_b200.so (C ABI: include/qpcontrol_b200.h).

There is no CPU fallback: `load()` raises if the shared library has not been built, and `DeviceController` raises if
the CUDA runtime reports no device.  The same binding code can drive another library exporting the setup half of the
ABI (the kernel-body emulation under tests/emu uses that, tests only).
"""
from __future__ import annotations

import ctypes as C
import os
from typing import Optional

import numpy as np

from .program import MATRIX_WEIGHT
from .controller import BatchResult
from .program import OSQPSettings, Program

_HERE = os.path.dirname(os.path.abspath(__file__))
# QPC_LIB_PATH: developer override used to A/B two builds of the same CUDA library on the GPU box
LIB_PATH = os.environ.get("QPC_LIB_PATH") or os.path.join(_HERE, "csrc", "libqpcontrol_b200.so")

_dp = C.POINTER(C.c_double)
_ip = C.POINTER(C.c_int32)

HOST_PTRS, DEVICE_PTRS = 0, 1


class qpc_settings(C.Structure):
    _fields_ = [("rho", C.c_double), ("sigma", C.c_double), ("alpha", C.c_double), ("eps_abs", C.c_double),
                ("eps_rel", C.c_double), ("eps_prim_inf", C.c_double), ("eps_dual_inf", C.c_double),
                ("adaptive_rho_tolerance", C.c_double), ("max_iter", C.c_int32), ("scaling", C.c_int32),
                ("adaptive_rho", C.c_int32), ("adaptive_rho_interval", C.c_int32), ("check_termination", C.c_int32),
                ("reserved", C.c_int32 * 3)]

    @staticmethod
    def from_py(s: OSQPSettings) -> "qpc_settings":
        out = qpc_settings()
        for f in ("rho", "sigma", "alpha", "eps_abs", "eps_rel", "eps_prim_inf", "eps_dual_inf",
                  "adaptive_rho_tolerance", "max_iter", "scaling", "adaptive_rho", "adaptive_rho_interval",
                  "check_termination"):
            setattr(out, f, getattr(s, f))
        return out


class qpc_batch_in(C.Structure):
    _fields_ = [("q", C.c_void_p), ("v", C.c_void_p), ("desired", C.c_void_p), ("desired_stride", C.c_int64),
                ("contact_weight", C.c_void_p), ("contact_maxnormalforce", C.c_void_p), ("contact_stride", C.c_int64),
                ("task_weight", C.c_void_p), ("task_weight_stride", C.c_int64),
                ("contact_geometry", C.c_void_p), ("contact_geometry_stride", C.c_int64),
                ("task_weight_matrix", C.c_void_p), ("task_weight_matrix_stride", C.c_int64),
                ("time", C.c_void_p), ("time_stride", C.c_int64)]


class qpc_interp_piece(C.Structure):
    _fields_ = [("break_start", C.c_double), ("x0", C.c_double), ("xf", C.c_double), ("y0", C.c_double * 4),
                ("dy", C.c_double * 3), ("angle", C.c_double), ("coeffs", C.c_double * 6), ("ncoeffs", C.c_int32),
                ("reserved", C.c_int32)]

    @staticmethod
    def array(pieces):
        arr = (qpc_interp_piece * len(pieces))()
        for a, d in zip(arr, pieces):
            a.break_start, a.x0, a.xf, a.angle = d["break_start"], d["x0"], d["xf"], d["angle"]
            y0 = np.zeros(4)
            y0[:len(d["y0"])] = d["y0"]
            a.y0[:] = y0.tolist()
            a.dy[:] = np.asarray(d["dy"], dtype=np.float64).tolist()
            a.ncoeffs = len(d["coeffs"])
            a.coeffs[:] = (list(map(float, d["coeffs"])) + [0.0] * 6)[:6]
        return arr


class qpc_contact_model(C.Structure):
    _fields_ = [("stiffness", C.c_double), ("damping", C.c_double), ("mu", C.c_double), ("v_eps", C.c_double),
                ("ground_z", C.c_double)]


class qpc_batch_out(C.Structure):
    _fields_ = [("tau", C.c_void_p), ("vdot", C.c_void_p), ("wrench", C.c_void_p), ("status", C.c_void_p),
                ("iters", C.c_void_p), ("residuals", C.c_void_p), ("factorizations", C.c_void_p)]


SETUP_SYMBOLS = ["qpc_version", "qpc_last_error", "qpc_device_count", "qpc_default_settings", "qpc_mechanism_create",
                 "qpc_mechanism_destroy", "qpc_mechanism_dims", "qpc_controller_create", "qpc_controller_destroy",
                 "qpc_add_contact", "qpc_set_contact_params", "qpc_add_task", "qpc_set_task_desired", "qpc_regularize",
                 "qpc_standing_setup", "qpc_set_settings", "qpc_finalize", "qpc_controller_dims",
                 "qpc_controller_weight_matrix_doubles", "qpc_add_se3pd", "qpc_se3pd_update"]
COMPUTE_SYMBOLS = ["qpc_solve_batch", "qpc_reserve", "qpc_launch_count", "qpc_assemble_batch", "qpc_solve_qp_batch",
                   "qpc_set_profiling", "qpc_stage_times", "qpc_measure_fp64_peak", "qpc_set_warm_start",
                   "qpc_reset_warm_start", "qpc_step_batch", "qpc_set_admm_elimination", "qpc_admm_eliminated",
                   "qpc_set_admm_warp", "qpc_admm_warp", "qpc_solve_batch_multi", "qpc_pin_host_buffer",
                   "qpc_unpin_host_buffer", "qpc_simulate_batch"]

_libs = {}


def load(path: Optional[str] = None):
    path = path or LIB_PATH
    if path in _libs:
        return _libs[path]
    if not os.path.exists(path):
        raise RuntimeError(f"{path} not found: build the CUDA library first (python -c 'import __graft_entry__ as g; "
                           f"g.build()' or make -C qpcontrol.jl_b200/csrc). There is no CPU fallback.")
    lib = C.CDLL(path)
    lib.qpc_last_error.restype = C.c_char_p
    lib.qpc_mechanism_create.restype = C.c_void_p
    lib.qpc_controller_create.restype = C.c_void_p
    if hasattr(lib, "qpc_launch_count"):
        lib.qpc_launch_count.restype = C.c_int64
    _libs[path] = lib
    return lib


def _c(a, dtype=np.float64):
    return None if a is None else np.ascontiguousarray(a, dtype=dtype)


def _p(a):
    # data_as keeps a reference to the array alive for as long as the pointer object lives
    return None if a is None else a.ctypes.data_as(C.c_void_p)


def check(lib, code: int, what: str):
    if code < 0:
        raise RuntimeError(f"{what} failed ({code}): {lib.qpc_last_error().decode()}")
    return code


class Handles:
    """Mechanism + controller handles built from a recorded Program through the setup half of the C ABI."""

    def __init__(self, lib, program: Program, device: int = 0):
        self.lib = lib
        self.program = program
        m = program.mechanism
        arrs = [_c(m.parent, np.int32), _c(m.jtype, np.int32), _c(m.axis), _c(m.X_R), _c(m.X_p), _c(m.mass),
                _c(m.com), _c(m.inertia_origin()), _c(m.gravity)]
        self.mech = C.c_void_p(lib.qpc_mechanism_create(C.c_int32(m.nb), *[_p(a) for a in arrs]))
        if not self.mech:
            raise RuntimeError("qpc_mechanism_create failed: " + lib.qpc_last_error().decode())
        st = qpc_settings.from_py(program.settings)
        self.ctrl = C.c_void_p(lib.qpc_controller_create(self.mech, C.c_int32(program.N),
                                                         C.c_int32(program.floating_body), C.byref(st)))
        if not self.ctrl:
            raise RuntimeError("qpc_controller_create failed: " + lib.qpc_last_error().decode())
        for kind, idx in program.events:
            if kind == "contact":
                c = program.contacts[idx]
                got = check(lib, lib.qpc_add_contact(self.ctrl, C.c_int32(c.body), _p(_c(c.position)),
                                                     _p(_c(c.normal)), C.c_double(c.mu)), "qpc_add_contact")
                assert got == idx
            else:
                e = program.tasks[idx]
                t = e.task
                got = check(lib, lib.qpc_add_task(self.ctrl, C.c_int32(t.kind), C.c_int32(t.source),
                                                  C.c_int32(t.target), C.c_int32(t.frame), _p(_c(np.array(t.point))),
                                                  C.c_int32(t.joint), C.c_int32(e.mode), C.c_double(e.weight),
                                                  _p(_c(e.W))), "qpc_add_task")
                assert got == idx
        for j in range(m.nb):
            r = m.velocity_range(j)
            if len(r) and program.reg[r[0]] != 0.0:
                check(lib, lib.qpc_regularize(self.ctrl, C.c_int32(j), C.c_double(float(program.reg[r[0]]))),
                      "qpc_regularize")
        s = program.standing
        if s is not None:
            check(lib, lib.qpc_standing_setup(
                self.ctrl, C.c_int32(s.linmom_task), C.c_int32(s.pelvis_task), C.c_int32(s.pelvis_body),
                C.c_int32(len(s.joints)), _p(_c(s.joint_tasks, np.int32)), _p(_c(s.joints, np.int32)),
                _p(_c(s.joint_kp)), _p(_c(s.joint_kd)), _p(_c(s.joint_ref)), C.c_double(s.com_kp),
                C.c_double(s.com_kd), C.c_double(s.pelvis_kp), C.c_double(s.pelvis_kd), _p(_c(s.comref))),
                "qpc_standing_setup")
        for i, sp in enumerate(program.se3pd):
            got = check(lib, lib.qpc_add_se3pd(self.ctrl, C.c_int32(sp.task), C.c_int32(sp.controller.base),
                                               C.c_int32(sp.controller.body), *self._se3pd_args(sp.controller)),
                        "qpc_add_se3pd")
            assert got == i
        check(lib, lib.qpc_finalize(self.ctrl, C.c_int32(device)), "qpc_finalize")
        self.sync_defaults()
        dims = [C.c_int32() for _ in range(7)]
        check(lib, lib.qpc_controller_dims(self.ctrl, *[C.byref(d) for d in dims]), "qpc_controller_dims")
        self.nq, self.nv, self.ndes, self.ncontacts, self.n, self.mg, self.nbox = [d.value for d in dims]

    @staticmethod
    def _se3pd_args(ctl):
        """gains and the two trajectory components of an SE3PDController in the layout of qpc_add_se3pd"""
        from .se3pd import compile_trajectory, gains_matrix
        traj = ctl.trajectory
        if not (hasattr(traj, "angular") and hasattr(traj, "linear")):
            raise TypeError("the device evaluates SE3Trajectory references (angular + linear components)")
        pa, pwa, enda = compile_trajectory(traj.angular, True)
        pl, pwl, endl = compile_trajectory(traj.linear, False)
        return (_p(gains_matrix(ctl.gains)), C.c_int32(len(pa)), qpc_interp_piece.array(pa), C.c_int32(int(pwa)),
                C.c_double(enda), C.c_int32(len(pl)), qpc_interp_piece.array(pl), C.c_int32(int(pwl)), C.c_double(endl))

    def sync_defaults(self):
        """Push the current `setdesired!` values and ContactPoint.weight / .maxnormalforce (mutable between ticks in
        the reference), the SE3PDControllers' trajectory / gains Refs (se3pdcontroller.jl:4-6) and the solver settings.
        The library re-uploads its tables only when something changed."""
        lib, pr = self.lib, self.program
        # nothing to do when nothing changed since the last push: ~40 ctypes calls saved per tick (0.1 ms of a 2.3 ms tick)
        snap = (tuple((c.weight, c.maxnormalforce) for c in pr.contacts),
                tuple(np.asarray(e.task.desired, dtype=np.float64).tobytes() for e in pr.tasks),
                tuple(getattr(pr.settings, f) for f in ("rho", "sigma", "alpha", "eps_abs", "eps_rel", "eps_prim_inf",
                                                        "eps_dual_inf", "adaptive_rho_tolerance", "max_iter", "scaling",
                                                        "adaptive_rho", "adaptive_rho_interval", "check_termination")),
                tuple((id(sp.controller.trajectory), id(sp.controller.gains)) for sp in pr.se3pd))
        if not pr.se3pd and snap == getattr(self, "_pushed", None):
            return
        for i, sp in enumerate(pr.se3pd):
            check(lib, lib.qpc_se3pd_update(self.ctrl, C.c_int32(i), *self._se3pd_args(sp.controller)), "qpc_se3pd_update")
        for i, c in enumerate(pr.contacts):
            check(lib, lib.qpc_set_contact_params(self.ctrl, C.c_int32(i), C.c_double(c.weight),
                                                  C.c_double(c.maxnormalforce)), "qpc_set_contact_params")
        for i, e in enumerate(pr.tasks):
            check(lib, lib.qpc_set_task_desired(self.ctrl, C.c_int32(i), _p(_c(e.task.desired))),
                  "qpc_set_task_desired")
        st = qpc_settings.from_py(pr.settings)
        check(lib, lib.qpc_set_settings(self.ctrl, C.byref(st)), "qpc_set_settings")
        self._pushed = snap

    def close(self):
        if getattr(self, "ctrl", None):
            self.lib.qpc_controller_destroy(self.ctrl)
            self.ctrl = None
        if getattr(self, "mech", None):
            self.lib.qpc_mechanism_destroy(self.mech)
            self.mech = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # ---- argument marshalling shared by every compute entry point ----------------------------------------------
    def batch_in(self, q, v, desired, cw, cm, ptr=lambda a: a.ctypes.data, keep=None, task_weight=None,
                 contact_geometry=None, task_weight_matrix=None, time=None):
        """Builds a qpc_batch_in from arrays (numpy for host pointers; `ptr` extracts the address).  task_weight
        [B, ntasks] / [ntasks] and contact_geometry [B, ncontacts, 7] / [ncontacts, 7] are the per-tick Parameters of
        the reference (task weights; contact position, normal, mu)."""
        bi = qpc_batch_in()
        bi.q, bi.v = ptr(q), ptr(v)
        bi.desired = None if desired is None else ptr(desired)
        bi.desired_stride = 0 if desired is None or desired.ndim == 1 else desired.shape[1]
        if (cw is None) != (cm is None):
            raise ValueError("give both contact_weight and contact_maxnormalforce or neither")
        bi.contact_weight = None if cw is None else ptr(cw)
        bi.contact_maxnormalforce = None if cm is None else ptr(cm)
        bi.contact_stride = 0 if cw is None or cw.ndim == 1 else cw.shape[1]
        if task_weight is not None:
            bi.task_weight = ptr(task_weight)
            bi.task_weight_stride = 0 if task_weight.ndim == 1 else task_weight.shape[1]
        if contact_geometry is not None:
            bi.contact_geometry = ptr(contact_geometry)
            bi.contact_geometry_stride = 0 if contact_geometry.ndim == 2 else contact_geometry.shape[1] * 7
        if task_weight_matrix is not None:
            bi.task_weight_matrix = ptr(task_weight_matrix)
            bi.task_weight_matrix_stride = 0 if task_weight_matrix.ndim == 1 else task_weight_matrix.shape[1]
        if time is not None:  # [B] controller times, or one value (shape () / (1,)) for the whole batch
            bi.time = ptr(time)
            bi.time_stride = 1 if time.ndim == 1 and time.shape[0] > 1 else 0
        return bi


def _prep_host_inputs(h: Handles, q, v, desired, cw, cm):
    q, v = np.atleast_2d(_c(q)), np.atleast_2d(_c(v))
    B = q.shape[0]
    if q.shape != (B, h.nq) or v.shape != (B, h.nv):
        raise ValueError(f"q must be [B,{h.nq}] and v [B,{h.nv}]")
    desired = _c(desired)
    if desired is not None and desired.shape[-1] != h.ndes:
        raise ValueError(f"desired must have {h.ndes} columns")
    cw, cm = _c(cw), _c(cm)
    if cw is not None and cm is None:
        cm = np.array([c.maxnormalforce for c in h.program.contacts])
    if cm is not None and cw is None:
        cw = np.array([c.weight for c in h.program.contacts])
    if cw is not None and cw.ndim != cm.ndim:
        if cw.ndim == 1:
            cw = np.ascontiguousarray(np.broadcast_to(cw, cm.shape))
        else:
            cm = np.ascontiguousarray(np.broadcast_to(cm, cw.shape))
    _check_rows(B, desired=(desired, 1), contact_weight=(cw, 1), contact_maxnormalforce=(cm, 1))
    if cw is not None and cw.shape[-1] != h.ncontacts:
        raise ValueError(f"contact arrays must have {h.ncontacts} columns")
    return q, v, desired, cw, cm, B


def _require(a, shape, name, dtype=np.float64):
    """Caller-owned host buffer handed to the C ABI as is: right shape, dtype and C-contiguous, or an error (the library
    would otherwise read or write out of bounds)."""
    if not isinstance(a, np.ndarray) or a.dtype != dtype or a.shape != tuple(shape) or not a.flags.c_contiguous:
        raise ValueError(f"{name} must be a C-contiguous {np.dtype(dtype).name} array of shape {tuple(shape)}; got "
                         f"{getattr(a, 'dtype', type(a))} {getattr(a, 'shape', None)}")


def _require_dev(t, shape, name, at_least=False):
    """Device tensor handed to the C ABI: contiguous, with the exact shape -- or, with `at_least`, the same row layout and
    at least as many rows (a larger tensor may back a smaller batch).  Nothing is read or written when a row is empty
    (e.g. the wrench output of a controller without contacts), so any tensor passes then."""
    sh = tuple(t.shape)
    need = 1
    for d in shape:
        need *= int(d)
    if at_least:
        ok = need == 0 or (len(sh) == len(shape) and sh[1:] == tuple(shape[1:]) and sh[0] >= shape[0])
    else:
        ok = sh == tuple(shape)
    if not ok or not t.is_contiguous():
        raise ValueError(f"{name} must be a contiguous device tensor of shape {tuple(shape)}; got {sh}")


def _prep_time(time, B):
    """Controller time(s) of the tick: None, one value for the batch, or one per instance -> contiguous float64 array"""
    if time is None:
        return None
    t = np.ascontiguousarray(np.atleast_1d(np.asarray(time, dtype=np.float64)))
    if t.ndim != 1 or t.shape[0] not in (1, B):
        raise ValueError(f"time must be a scalar or have one entry per instance ({B}); got shape {t.shape}")
    return t


def _check_rows(B, **arrays):
    """Per-instance arrays must have exactly B rows: the C ABI copies B * stride doubles from them (a shorter array would
    be read out of bounds).  `base_ndim` = the number of dimensions of ONE row (a broadcast row has that many)."""
    for name, (a, base_ndim) in arrays.items():
        if a is None or a.ndim == base_ndim:
            continue
        if a.ndim != base_ndim + 1 or a.shape[0] != B:
            raise ValueError(f"{name} must have one row per instance ({B}) or be a single broadcast row; got {a.shape}")


def _prep_tick_parameters(h: Handles, task_weight, contact_geometry, B=None, task_weight_matrix=None):
    tw, cg, twm = _c(task_weight), _c(contact_geometry), _c(task_weight_matrix)
    if tw is not None and tw.shape[-1] != len(h.program.tasks):
        raise ValueError(f"task_weight must have {len(h.program.tasks)} columns (one per task, addtask! order)")
    if cg is not None and cg.shape[-2:] != (h.ncontacts, 7):
        raise ValueError(f"contact_geometry must be [..., {h.ncontacts}, 7] (position, normal, mu per contact)")
    if twm is not None:
        nw = sum(e.task.dimension ** 2 for e in h.program.tasks if e.mode == MATRIX_WEIGHT)
        if twm.shape[-1] != nw:
            raise ValueError(f"task_weight_matrix must have {nw} columns: the dim x dim weights (row-major) of the "
                             "matrix-weighted tasks, concatenated in addtask! order")
    if B is not None:
        _check_rows(B, task_weight=(tw, 1), contact_geometry=(cg, 2), task_weight_matrix=(twm, 1))
    if task_weight_matrix is None:
        return tw, cg
    return tw, cg, twm


def _alloc_out(h: Handles, B):
    return BatchResult(tau=np.zeros((B, h.nv)), vdot=np.zeros((B, h.nv)), wrenches=np.zeros((B, h.ncontacts, 6)),
                       status=np.zeros(B, np.int32), iters=np.zeros(B, np.int32), residuals=np.zeros((B, 2)),
                       factorizations=np.zeros(B, np.int32))


def _batch_out(res: BatchResult, ptr=lambda a: a.ctypes.data):
    bo = qpc_batch_out()
    bo.tau, bo.vdot, bo.wrench = ptr(res.tau), ptr(res.vdot), ptr(res.wrenches)
    bo.status, bo.iters, bo.residuals = ptr(res.status), ptr(res.iters), ptr(res.residuals)
    if getattr(res, "factorizations", None) is not None:
        bo.factorizations = ptr(res.factorizations)
    return bo


def solve_host_multi(devs, q, v, desired=None, contact_weight=None, contact_maxnormalforce=None, task_weight=None,
                     contact_geometry=None, time=None) -> BatchResult:
    """One batch over several devices from this process (qpc_solve_batch_multi): `devs` are DeviceControllers of replicas
    of one program, finalized on different devices; contiguous shards, one host thread per device inside the library."""
    h = devs[0].h
    for d in devs:
        d.h.sync_defaults()
    q, v, desired, cw, cm, B = _prep_host_inputs(h, q, v, desired, contact_weight, contact_maxnormalforce)
    tw, cg = _prep_tick_parameters(h, task_weight, contact_geometry, B)
    res = _alloc_out(h, B)
    bi, bo = h.batch_in(q, v, desired, cw, cm, task_weight=tw, contact_geometry=cg, time=_prep_time(time, B)), _batch_out(res)
    arr = (C.c_void_p * len(devs))(*[d.h.ctrl for d in devs])
    check(devs[0].lib, devs[0].lib.qpc_solve_batch_multi(arr, C.c_int32(len(devs)), C.c_int64(B), C.byref(bi), C.byref(bo)),
          "qpc_solve_batch_multi")
    return res


def pin_host_buffer(a: np.ndarray):
    """Page-lock a numpy array used as a QPC_HOST_PTRS buffer (qpc_pin_host_buffer); unpin before it is freed."""
    lib = load()
    check(lib, lib.qpc_pin_host_buffer(C.c_void_p(a.ctypes.data), C.c_int64(a.nbytes)), "qpc_pin_host_buffer")


def unpin_host_buffer(a: np.ndarray):
    lib = load()
    check(lib, lib.qpc_unpin_host_buffer(C.c_void_p(a.ctypes.data)), "qpc_unpin_host_buffer")


class DeviceController:
    """The CUDA controller behind `MomentumBasedController.finalize()`."""

    def __init__(self, program: Program, device: int = 0):
        self.lib = load()
        ndev = self.lib.qpc_device_count()
        if ndev <= 0:
            raise RuntimeError("no CUDA device visible: the qpcontrol_b200 hot path has no CPU fallback")
        self.h = Handles(self.lib, program, device)
        self.program = program
        self.device = device

    @property
    def dims(self):
        return dict(nq=self.h.nq, nv=self.h.nv, ndes=self.h.ndes, ncontacts=self.h.ncontacts, n=self.h.n,
                    mg=self.h.mg, nbox=self.h.nbox)

    def reserve(self, B: int):
        check(self.lib, self.lib.qpc_reserve(self.h.ctrl, C.c_int64(B)), "qpc_reserve")

    def launch_count(self) -> int:
        return int(self.lib.qpc_launch_count(self.h.ctrl))

    def set_profiling(self, on: bool):
        check(self.lib, self.lib.qpc_set_profiling(self.h.ctrl, C.c_int32(int(on))), "qpc_set_profiling")

    def stage_times(self):
        """{assembly, ADMM, inverse dynamics} milliseconds of the last tick (CUDA events on the launching stream)."""
        ms = (C.c_double * 3)()
        check(self.lib, self.lib.qpc_stage_times(self.h.ctrl, ms), "qpc_stage_times")
        return [ms[0], ms[1], ms[2]]

    def set_warm_start(self, on: bool):
        """OSQP's implicit warm start between ticks (previous x, y, rho of the same batch slot); off = cold starts."""
        check(self.lib, self.lib.qpc_set_warm_start(self.h.ctrl, C.c_int32(int(on))), "qpc_set_warm_start")

    def reset_warm_start(self):
        check(self.lib, self.lib.qpc_reset_warm_start(self.h.ctrl), "qpc_reset_warm_start")

    def set_admm_elimination(self, on: bool):
        """Allow (default) or forbid the ADMM fast path that eliminates the diagonal-cost free variables."""
        check(self.lib, self.lib.qpc_set_admm_elimination(self.h.ctrl, C.c_int32(int(on))), "qpc_set_admm_elimination")

    def admm_eliminated(self) -> int:
        """Number of variables the next tick eliminates from the KKT system (0 = full system)."""
        return int(self.lib.qpc_admm_eliminated(self.h.ctrl))

    def set_admm_warp(self, on: bool):
        """Allow (default) or forbid the one-warp-per-QP ADMM kernel on the reduced problem."""
        check(self.lib, self.lib.qpc_set_admm_warp(self.h.ctrl, C.c_int32(int(on))), "qpc_set_admm_warp")

    def admm_warp(self) -> bool:
        """True when the next tick runs the one-warp-per-QP ADMM kernel."""
        return bool(self.lib.qpc_admm_warp(self.h.ctrl))

    def simulate_host(self, q, v, dt: float, nticks: int, ground_z: float, substeps: int = 4, stiffness: float = 5e4,
                      damping: float = 1e3, mu: float = 0.8, v_eps: float = 1e-2, contact_weight=None,
                      contact_maxnormalforce=None, desired=None, time=None):
        """`nticks` control ticks of period dt with a PLANT between them (qpc_simulate_batch): forward dynamics under a
        soft ground contact at z = ground_z, `substeps` integration steps per tick.  Returns (q, v, last tick's result)."""
        h = self.h
        h.sync_defaults()
        q, v, desired, cw, cm, B = _prep_host_inputs(h, q, v, desired, contact_weight, contact_maxnormalforce)
        q, v = q.copy(), v.copy()
        res = _alloc_out(h, B)
        bi, bo = h.batch_in(q, v, desired, cw, cm, time=_prep_time(time, B)), _batch_out(res)  # time: of the first tick
        plant = qpc_contact_model(stiffness, damping, mu, v_eps, ground_z)
        check(self.lib, self.lib.qpc_simulate_batch(h.ctrl, C.c_int64(B), C.c_void_p(q.ctypes.data),
                                                    C.c_void_p(v.ctypes.data), C.byref(bi), C.byref(bo), C.byref(plant),
                                                    C.c_double(dt), C.c_int32(substeps), C.c_int32(nticks),
                                                    C.c_int32(HOST_PTRS), None), "qpc_simulate_batch")
        return q, v, res

    def step_host(self, q, v, dt: float, nsteps: int, desired=None, contact_weight=None, contact_maxnormalforce=None,
                  task_weight=None, contact_geometry=None, time=None):
        """`nsteps` closed-loop ticks on the device (qpc_step_batch); returns (q, v, result of the last tick)."""
        h = self.h
        h.sync_defaults()
        q, v, desired, cw, cm, B = _prep_host_inputs(h, q, v, desired, contact_weight, contact_maxnormalforce)
        tw, cg = _prep_tick_parameters(h, task_weight, contact_geometry)
        q, v = q.copy(), v.copy()
        res = _alloc_out(h, B)
        bi, bo = h.batch_in(q, v, desired, cw, cm, task_weight=tw, contact_geometry=cg, time=_prep_time(time, B)), _batch_out(res)
        check(self.lib, self.lib.qpc_step_batch(h.ctrl, C.c_int64(B), C.c_void_p(q.ctypes.data),
                                                C.c_void_p(v.ctypes.data), C.byref(bi), C.byref(bo), C.c_double(dt),
                                                C.c_int32(nsteps), C.c_int32(HOST_PTRS), None), "qpc_step_batch")
        return q, v, res

    def step_device(self, B: int, q, v, dt: float, nsteps: int, out: Optional[dict] = None, contact_weight=None,
                    contact_maxnormalforce=None, stream: int = 0):
        """Closed-loop ticks on device tensors, in place and asynchronous on `stream`."""
        ptr = lambda t: t.data_ptr()  # noqa: E731
        bi = qpc_batch_in()
        bi.q, bi.v = ptr(q), ptr(v)
        bi.contact_weight = None if contact_weight is None else ptr(contact_weight)
        bi.contact_maxnormalforce = None if contact_maxnormalforce is None else ptr(contact_maxnormalforce)
        bi.contact_stride = 0 if contact_weight is None or contact_weight.dim() == 1 else contact_weight.shape[1]
        bo = qpc_batch_out()
        for name in ("tau", "vdot", "wrench", "status", "iters", "residuals", "factorizations"):
            t = None if out is None else out.get(name)
            setattr(bo, name, None if t is None else ptr(t))
        check(self.lib, self.lib.qpc_step_batch(self.h.ctrl, C.c_int64(B), C.c_void_p(ptr(q)), C.c_void_p(ptr(v)),
                                                C.byref(bi), C.byref(bo), C.c_double(dt), C.c_int32(nsteps),
                                                C.c_int32(DEVICE_PTRS), C.c_void_p(stream)), "qpc_step_batch")

    def solve_host(self, q, v, desired=None, contact_weight=None, contact_maxnormalforce=None, task_weight=None,
                   contact_geometry=None, task_weight_matrix=None, time=None) -> BatchResult:
        """Host numpy buffers in, host numpy buffers out (H2D + kernels + D2H inside the call).  `time`: the functor's t
        (scalar or [B]) for controllers with device-side SE3PDControllers."""
        h = self.h
        h.sync_defaults()
        q, v, desired, cw, cm, B = _prep_host_inputs(h, q, v, desired, contact_weight, contact_maxnormalforce)
        twm = None
        if task_weight_matrix is None:
            tw, cg = _prep_tick_parameters(h, task_weight, contact_geometry, B)
        else:
            tw, cg, twm = _prep_tick_parameters(h, task_weight, contact_geometry, B, task_weight_matrix)
        res = _alloc_out(h, B)
        bi, bo = h.batch_in(q, v, desired, cw, cm, task_weight=tw, contact_geometry=cg, task_weight_matrix=twm,
                            time=_prep_time(time, B)), _batch_out(res)
        check(self.lib, self.lib.qpc_solve_batch(h.ctrl, C.c_int64(B), C.byref(bi), C.byref(bo), C.c_int32(HOST_PTRS),
                                                 None), "qpc_solve_batch")
        return res

    def solve_host_into(self, q, v, res: BatchResult, contact_weight=None, contact_maxnormalforce=None):
        """As `solve_host`, writing into caller-owned (e.g. pinned) buffers: no allocation, no marshalling copies.
        The buffers must already be C-contiguous float64 of the right shape (checked, never copied)."""
        h = self.h
        h.sync_defaults()  # cheap: unchanged values do not re-upload the program
        B = q.shape[0]
        _require(q, (B, h.nq), "q")
        _require(v, (B, h.nv), "v")
        for name, a in (("contact_weight", contact_weight), ("contact_maxnormalforce", contact_maxnormalforce)):
            if a is not None:
                _require(a, (B, h.ncontacts) if a.ndim == 2 else (h.ncontacts,), name)
        _require(res.tau, (B, h.nv), "res.tau")
        _require(res.vdot, (B, h.nv), "res.vdot")
        _require(res.wrenches, (B, h.ncontacts, 6), "res.wrenches")
        _require(res.status, (B,), "res.status", np.int32)
        _require(res.iters, (B,), "res.iters", np.int32)
        _require(res.residuals, (B, 2), "res.residuals")
        bi, bo = h.batch_in(q, v, None, contact_weight, contact_maxnormalforce), _batch_out(res)
        check(self.lib, self.lib.qpc_solve_batch(h.ctrl, C.c_int64(B), C.byref(bi), C.byref(bo), C.c_int32(HOST_PTRS),
                                                 None), "qpc_solve_batch")
        return res

    def solve_device(self, B: int, q, v, out: dict, desired=None, contact_weight=None, contact_maxnormalforce=None,
                     stream: int = 0):
        """Device pointers in/out (torch CUDA tensors or anything with `.data_ptr()`), asynchronous on `stream`
        (a raw cudaStream_t value).  `out` maps tau / vdot / wrench / status / iters / residuals to tensors."""
        h = self.h
        ptr = lambda t: t.data_ptr()  # noqa: E731
        _require_dev(q, (B, h.nq), "q", at_least=True)  # a larger tensor may back a smaller batch: the first B rows are used
        _require_dev(v, (B, h.nv), "v", at_least=True)
        for name, t, cols in (("desired", desired, h.ndes), ("contact_weight", contact_weight, h.ncontacts),
                              ("contact_maxnormalforce", contact_maxnormalforce, h.ncontacts)):
            if t is not None:
                _require_dev(t, (B, cols) if t.dim() == 2 else (cols,), name, at_least=t.dim() == 2)
        for name, shape in (("tau", (B, h.nv)), ("vdot", (B, h.nv)), ("wrench", (B, h.ncontacts, 6)), ("status", (B,)),
                            ("iters", (B,)), ("residuals", (B, 2)), ("factorizations", (B,))):
            if out.get(name) is not None:
                _require_dev(out[name], shape, name, at_least=True)
        bi = qpc_batch_in()
        bi.q, bi.v = ptr(q), ptr(v)
        bi.desired = None if desired is None else ptr(desired)
        bi.desired_stride = 0 if desired is None or desired.dim() == 1 else desired.shape[1]
        bi.contact_weight = None if contact_weight is None else ptr(contact_weight)
        bi.contact_maxnormalforce = None if contact_maxnormalforce is None else ptr(contact_maxnormalforce)
        bi.contact_stride = 0 if contact_weight is None or contact_weight.dim() == 1 else contact_weight.shape[1]
        bo = qpc_batch_out()
        for name in ("tau", "vdot", "wrench", "status", "iters", "residuals", "factorizations"):
            t = out.get(name)
            setattr(bo, name, None if t is None else ptr(t))
        check(self.lib, self.lib.qpc_solve_batch(h.ctrl, C.c_int64(B), C.byref(bi), C.byref(bo),
                                                 C.c_int32(DEVICE_PTRS), C.c_void_p(stream)), "qpc_solve_batch")

    def assemble_host(self, q, v, desired=None, contact_weight=None, contact_maxnormalforce=None, task_weight=None,
                      contact_geometry=None, time=None):
        """Stage-level entry point: the condensed QP of every instance."""
        h = self.h
        h.sync_defaults()
        q, v, desired, cw, cm, B = _prep_host_inputs(h, q, v, desired, contact_weight, contact_maxnormalforce)
        tw, cg = _prep_tick_parameters(h, task_weight, contact_geometry)
        out = dict(P=np.zeros((B, h.n, h.n)), q=np.zeros((B, h.n)), G=np.zeros((B, h.mg, h.n)),
                   lg=np.zeros((B, h.mg)), ug=np.zeros((B, h.mg)), lb=np.zeros((B, h.nbox)), ub=np.zeros((B, h.nbox)),
                   desired=np.zeros((B, h.ndes)))
        bi = h.batch_in(q, v, desired, cw, cm, task_weight=tw, contact_geometry=cg, time=_prep_time(time, B))
        check(self.lib, self.lib.qpc_assemble_batch(h.ctrl, C.c_int64(B), C.byref(bi), _p(out["P"]), _p(out["q"]),
                                                    _p(out["G"]), _p(out["lg"]), _p(out["ug"]), _p(out["lb"]),
                                                    _p(out["ub"]), _p(out["desired"]), C.c_int32(HOST_PTRS), None),
              "qpc_assemble_batch")
        return out


def measure_fp64_peak(device: int = 0) -> float:
    lib = load()
    tf = C.c_double()
    check(lib, lib.qpc_measure_fp64_peak(C.c_int32(device), C.byref(tf)), "qpc_measure_fp64_peak")
    return tf.value


def solve_qp_batch_host(P, qv, G, lg, ug, lb=None, ub=None, settings: Optional[OSQPSettings] = None, device: int = 0):
    """Raw batched dense QPs through qpc_solve_qp_batch with host buffers."""
    lib = load()
    if lib.qpc_device_count() <= 0:
        raise RuntimeError("no CUDA device visible: the qpcontrol_b200 hot path has no CPU fallback")
    P, qv, G, lg, ug = (_c(a) for a in (P, qv, G, lg, ug))
    B, n = qv.shape
    mg = lg.shape[1]
    nbox = 0 if lb is None else lb.shape[1]
    lb = _c(lb) if nbox else np.zeros((B, 0))
    ub = _c(ub) if nbox else np.zeros((B, 0))
    st = qpc_settings.from_py(settings or OSQPSettings())
    out = dict(x=np.zeros((B, n)), y=np.zeros((B, mg + nbox)), status=np.zeros(B, np.int32),
               iters=np.zeros(B, np.int32), res=np.zeros((B, 2)))
    check(lib, lib.qpc_solve_qp_batch(C.c_int32(device), C.c_int64(B), C.c_int32(n), C.c_int32(mg), C.c_int32(nbox),
                                      _p(P), _p(qv), _p(G), _p(lg), _p(ug), _p(lb), _p(ub), C.byref(st), _p(out["x"]),
                                      _p(out["y"]), _p(out["status"]), _p(out["iters"]), _p(out["res"]),
                                      C.c_int32(HOST_PTRS), None), "qpc_solve_qp_batch")
    return out


def solve_qp_batch_device(B, n, mg, nbox, P, qv, G, lg, ug, lb, ub, settings: Optional[OSQPSettings], x, y, status, iters,
                          residuals, device: int = 0, stream: int = 0):
    """Raw batched dense QPs through qpc_solve_qp_batch with DEVICE buffers (objects with .data_ptr(), e.g. torch
    tensors): asynchronous on `stream`, no host copies."""
    lib = load()
    st = qpc_settings.from_py(settings or OSQPSettings())
    dp = lambda t: None if t is None else C.c_void_p(t.data_ptr())
    check(lib, lib.qpc_solve_qp_batch(C.c_int32(device), C.c_int64(B), C.c_int32(n), C.c_int32(mg), C.c_int32(nbox),
                                      dp(P), dp(qv), dp(G), dp(lg), dp(ug), dp(lb), dp(ub), C.byref(st), dp(x), dp(y),
                                      dp(status), dp(iters), dp(residuals), C.c_int32(DEVICE_PTRS),
                                      C.c_void_p(stream) if stream else None), "qpc_solve_qp_batch")
