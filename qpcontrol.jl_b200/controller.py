"""Host-side mirror of the reference's controllers; the numerics run in libqpcontrol_b200.so (CUDA, sm_100a).

Mirrors, with the same names, argument meaning and error behaviour:
  * `MomentumBasedController{N}(mechanism, optimizer; floatingjoint)` -- reference `src/lowlevel/momentum.jl:1-34`,
    `addtask!` (:99-117, hard / scalar weight / matrix weight), `regularize!` (:128-131), `addcontact!` (:133-148),
    the functor `(controller)(tau, t, x)` (:41-81) and `checkstatus` (:83-91),
  * `StandingController(lowlevel, feet, pelvis, nominalstate; ...)` -- reference `src/highlevel/standing.jl:18-56`
    and its functor (:58-89).

Differences forced by batching: the functors take `q [B, nq]` and `v [B, nv]` (B independent robot instances) and
return a `BatchResult`; `t` is dropped because neither reference functor reads it.  The
"optimizer" argument is the `OSQPSettings` the reference would have put into `OSQP.Optimizer()`.

There is no CPU fallback: constructing the device controller raises if the CUDA library is missing or no GPU is
visible.  Setup-time recording (this file up to `finalize`) needs neither.
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import Dict, List, Optional, Sequence, Union

import numpy as np

from . import program as P
from .mechanism import Mechanism, QUAT_FLOATING, REVOLUTE
from .program import (AbstractMotionTask, AngularAccelerationTask, ContactPoint, JointAccelerationTask,
                      LinearMomentumRateTask, OSQPSettings, Program, QPSolveFailure, SpatialAccelerationTask,
                      StandingSpec, TaskEntry, checkstatus)


@dataclass
class BatchResult:
    """What the reference leaves in `tau`, `controller.result.v̇` and `controller.contactwrenches` (momentum.jl:62-80),
    for every instance, plus the solver diagnostics OSQP reports."""
    tau: np.ndarray  # [B, nv]   joint torques, floating-joint entries exactly 0
    vdot: np.ndarray  # [B, nv]
    wrenches: np.ndarray  # [B, ncontacts, 6] world-frame wrench (angular; linear) of every contact point
    status: np.ndarray  # [B] int32, OSQP status codes
    iters: np.ndarray  # [B] int32
    residuals: np.ndarray  # [B, 2] primal, dual (unscaled, OSQP definition on the QP the device solves)
    factorizations: Optional[np.ndarray] = None  # [B] int32, KKT factorisations (1 + rho updates) of every solve

    def contactwrenches(self, program: Program) -> Dict[int, np.ndarray]:
        """`controller.contactwrenches`: per-body sums in world frame (momentum.jl:65-72) -> {body: [B, 6]}."""
        out: Dict[int, np.ndarray] = {}
        for i, c in enumerate(program.contacts):
            out[c.body] = out.get(c.body, 0) + self.wrenches[:, i, :]
        return out


class MomentumBasedController:
    def __init__(self, mechanism: Mechanism, optimizer: Optional[OSQPSettings] = None, N: int = 4,
                 floatingjoint: Optional[int] = None, device: int = 0):
        if floatingjoint is not None and mechanism.jtype[floatingjoint] != QUAT_FLOATING:
            raise ValueError("floatingjoint must be a QuaternionFloating joint")
        self.mechanism = mechanism
        self.program = Program(mechanism, int(N), -1 if floatingjoint is None else int(floatingjoint),
                               optimizer if optimizer is not None else OSQPSettings())
        self.device = device
        self._dev = None  # device controller, created by finalize() (the reference's lazy `initialize!`)

    # -- setup API -----------------------------------------------------------------------------------------------
    def _check_open(self):
        if self._dev is not None:
            raise RuntimeError("controller already initialized; tasks/contacts cannot be added after the first solve")

    def addtask(self, task: AbstractMotionTask, weight: Union[None, float, np.ndarray] = None):
        """`addtask!`: no weight -> hard constraint `task_error == 0`; number -> cost `w * e'e`; matrix -> `e' W e`."""
        self._check_open()
        if weight is None:
            entry = TaskEntry(task, P.HARD)
        elif np.isscalar(weight):
            entry = TaskEntry(task, P.SCALAR_WEIGHT, weight=float(weight))
        else:
            W = np.ascontiguousarray(weight, dtype=np.float64)
            if W.shape != (task.dimension, task.dimension):
                raise ValueError("weight matrix must be dimension x dimension")
            entry = TaskEntry(task, P.MATRIX_WEIGHT, W=W)
        self.program.tasks.append(entry)
        self.program.events.append(("task", len(self.program.tasks) - 1))
        return len(self.program.tasks) - 1

    def regularize(self, joint: int, weight: float):
        """`regularize!`: adds `weight * vd_joint . vd_joint` to the objective."""
        self._check_open()
        for k in self.mechanism.velocity_range(joint):
            self.program.reg[k] += float(weight)

    def addcontact(self, body: int, position, normal, mu: float) -> ContactPoint:
        """`addcontact!(controller, body, position, normal, mu)`; position and normal in the body frame."""
        self._check_open()
        point = ContactPoint(body, position, normal, mu, self.program.N)
        self.program.contacts.append(point)
        self.program.events.append(("contact", len(self.program.contacts) - 1))
        return point

    def bind_se3pd(self, task: int, controller) -> int:
        """Lets the device evaluate `setdesired!(task, controller(t, state))` every tick: `controller` is an
        `SE3PDController` (se3pdcontroller.jl:1-18) whose reference is an `SE3Trajectory` of Interpolated / Constant /
        Piecewise components, `task` the index `addtask` returned for the SpatialAccelerationTask it drives (expressed
        in the controller's body frame).  The tick's `time` argument is the functor's t; trajectory and gains may be
        replaced between ticks (they are Refs in the reference).  Returns the binding's index."""
        self._check_open()
        from .program import SE3PDSpec, SpatialAccelerationTask
        if not (0 <= task < len(self.program.tasks)) or not isinstance(self.program.tasks[task].task, SpatialAccelerationTask):
            raise ValueError("bind_se3pd: `task` must be the index of a SpatialAccelerationTask")
        self.program.se3pd.append(SE3PDSpec(int(task), controller))
        return len(self.program.se3pd) - 1

    # -- device side ---------------------------------------------------------------------------------------------
    def finalize(self):
        """`initialize!` (momentum.jl:150-156): freezes the program and builds the device controller."""
        if self._dev is None:
            from ._lib import DeviceController  # raises loudly if the CUDA library cannot be loaded
            self._dev = DeviceController(self.program, self.device)
        return self._dev

    def __call__(self, q, v, desired=None, contact_weight=None, contact_maxnormalforce=None,
                 check: bool = True, task_weight=None, contact_geometry=None, task_weight_matrix=None,
                 time=None) -> BatchResult:
        """The control tick for B instances.  `desired` [B, ndes] (task order) overrides the tasks' `setdesired!`
        values; contact arrays [B, ncontacts] override the ContactPoint fields per instance; `task_weight`
        [B, ntasks] and `contact_geometry` [B, ncontacts, 7] = (position, normal, mu) are the reference's
        Parameter-valued task weights and contact frames (momentum.jl:107-110, contacts.jl:39,53-61);
        `task_weight_matrix` [B, sum dim^2] are Parameter-valued MATRIX weights (momentum.jl:113-117): the dim x dim
        matrices (row-major) of the matrix-weighted tasks, concatenated in addtask! order.  `time` (scalar or [B]) is
        the functor's t, read by the SE3PDControllers bound with `bind_se3pd`."""
        dev = self.finalize()
        res = dev.solve_host(q, v, desired, contact_weight, contact_maxnormalforce, task_weight, contact_geometry,
                             task_weight_matrix, time)
        if check:
            checkstatus(res.status)
        return res


    def set_warm_start(self, on: bool = True):
        """Sequential ticks start from the previous tick's iterates and rho of the same batch slot, as the reference's
        single OSQP workspace does between `solve!` calls (momentum.jl:58)."""
        self.finalize().set_warm_start(on)

    def reset_warm_start(self):
        self.finalize().reset_warm_start()

    def simulate_plant(self, q, v, dt: float, nticks: int, ground_z: float, substeps: int = 8, check: bool = True, **plant):
        """`simulate(state, T, PeriodicController(tau, dt, controller))` of notebooks/Standing controller.ipynb:202-214 for a
        batch, WITH a plant: after every control tick the forward dynamics under a soft ground contact at z = ground_z
        advances the robots by `substeps` steps of dt / substeps (qpc_simulate_batch).  Returns (q, v, last result)."""
        q, v, res = self.finalize().simulate_host(q, v, dt, nticks, ground_z, substeps=substeps, **plant)
        if check:
            checkstatus(res.status)
        return q, v, res

    def simulate(self, q, v, dt: float, nsteps: int, desired=None, contact_weight=None, contact_maxnormalforce=None,
                 check: bool = True, time=None):
        """Closed loop of `nsteps` control ticks at period `dt` for B instances, on the device
        (notebooks/Standing controller.ipynb:202-214 batched; see qpc_step_batch).  Returns (q, v, last BatchResult)."""
        dev = self.finalize()
        q, v, res = dev.step_host(q, v, dt, nsteps, desired, contact_weight, contact_maxnormalforce, time=time)
        if check:
            checkstatus(res.status)
        return q, v, res


class StandingController:
    """reference `src/highlevel/standing.jl`.  `feet` / `pelvis` are body indices, `nominal_q` the nominal
    configuration; keyword defaults are the reference's (:23-29)."""

    def __init__(self, lowlevel: MomentumBasedController, feet: Sequence[int], pelvis: int, nominal_q: np.ndarray,
                 joint_regularization: float = 0.05, linear_momentum_weight: float = 1.0,
                 comgains=(10.0, 2 * np.sqrt(10.0)), pelvisgains=(20.0, 2 * np.sqrt(20.0)),
                 jointgains=(100.0, 20.0), comref: Optional[np.ndarray] = None,
                 nominal_com: Optional[np.ndarray] = None):
        mech = lowlevel.mechanism
        self.lowlevel = lowlevel
        self.robotmass = mech.total_mass
        world = -1
        for j in range(mech.nb):  # regularize!.(Ref(lowlevel), tree_joints(mechanism), joint_regularization)
            lowlevel.regularize(j, joint_regularization)
        self.foottasks = {}
        for foot in feet:
            task = SpatialAccelerationTask(mech, world, foot)
            self.foottasks[foot] = (task, lowlevel.addtask(task))
        self.linmomtask = LinearMomentumRateTask(mech)
        linmom_idx = lowlevel.addtask(self.linmomtask, float(linear_momentum_weight))
        self.pelvistask = AngularAccelerationTask(mech, world, pelvis)
        pelvis_idx = lowlevel.addtask(self.pelvistask)
        onpaths = set()
        for foot in feet:
            onpaths.update(b for b, _ in mech.path(world, foot))
        self.jointtasks = {}
        joints, jtasks = [], []
        for j in range(mech.nb):
            if mech.jtype[j] == REVOLUTE and j not in onpaths:
                task = JointAccelerationTask(mech, j)
                idx = lowlevel.addtask(task)
                self.jointtasks[j] = task
                joints.append(j)
                jtasks.append(idx)
        if comref is None:
            if nominal_com is None:
                nominal_com = center_of_mass_host(mech, nominal_q)
            comref = np.asarray(nominal_com, dtype=np.float64) - np.array([0.0, 0.0, 0.05])
        kp = np.full(len(joints), float(jointgains[0]))
        kd = np.full(len(joints), float(jointgains[1]))
        ref = np.array([nominal_q[mech.qoff[j]] for j in joints], dtype=np.float64)
        lowlevel.program.standing = StandingSpec(linmom_idx, pelvis_idx, int(pelvis), jtasks, joints, kp, kd, ref,
                                                 float(comgains[0]), float(comgains[1]), float(pelvisgains[0]),
                                                 float(pelvisgains[1]), np.asarray(comref, dtype=np.float64))

    def __call__(self, q, v, contact_weight=None, contact_maxnormalforce=None, check: bool = True) -> BatchResult:
        """PD laws for CoM / pelvis / joints (standing.jl:60-85) are evaluated on the device, then the low-level
        controller runs (standing.jl:87)."""
        return self.lowlevel(q, v, None, contact_weight, contact_maxnormalforce, check=check)

    def simulate(self, q, v, dt: float, nsteps: int, contact_weight=None, contact_maxnormalforce=None,
                 check: bool = True):
        return self.lowlevel.simulate(q, v, dt, nsteps, None, contact_weight, contact_maxnormalforce, check=check)


def center_of_mass_host(mech: Mechanism, q: np.ndarray) -> np.ndarray:
    """Setup-time helper (the reference evaluates `center_of_mass(nominalstate)` once in the constructor,
    standing.jl:28): plain forward kinematics in numpy."""
    R = [None] * mech.nb
    p = [None] * mech.nb
    com = np.zeros(3)
    for i in range(mech.nb):
        o = mech.qoff[i]
        Rj, pj = np.eye(3), np.zeros(3)
        if mech.jtype[i] == REVOLUTE:
            a = mech.axis[i]
            K = np.array([[0, -a[2], a[1]], [a[2], 0, -a[0]], [-a[1], a[0], 0]])
            Rj = np.eye(3) + np.sin(q[o]) * K + (1 - np.cos(q[o])) * (K @ K)
        elif mech.jtype[i] == QUAT_FLOATING:
            w, x, y, z = q[o:o + 4] / np.linalg.norm(q[o:o + 4])
            Rj = np.array([[1 - 2 * (y * y + z * z), 2 * (x * y - w * z), 2 * (x * z + w * y)],
                           [2 * (x * y + w * z), 1 - 2 * (x * x + z * z), 2 * (y * z - w * x)],
                           [2 * (x * z - w * y), 2 * (y * z + w * x), 1 - 2 * (x * x + y * y)]])
            pj = q[o + 4:o + 7]
        elif mech.jtype[i] == 1:  # prismatic
            pj = q[o] * mech.axis[i]
        Rl = mech.X_R[i] @ Rj
        pl = mech.X_R[i] @ pj + mech.X_p[i]
        if mech.parent[i] < 0:
            R[i], p[i] = Rl, pl
        else:
            R[i] = R[mech.parent[i]] @ Rl
            p[i] = R[mech.parent[i]] @ pl + p[mech.parent[i]]
        com += mech.mass[i] * (R[i] @ mech.com[i] + p[i])
    return com / mech.total_mass
