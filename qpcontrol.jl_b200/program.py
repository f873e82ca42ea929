"""Host-side mirror of the reference's problem vocabulary (setup-time API).

Mirrors, with the same names and argument meaning:
  * the task types of reference `src/tasks.jl` (`SpatialAccelerationTask` :3-44, `AngularAccelerationTask` :47-84,
    `LinearAccelerationTask` :86-123, `PointAccelerationTask` :125-171, `JointAccelerationTask` :173-189,
    `MomentumRateTask` :192-236, `LinearMomentumRateTask` :239-262) with `dimension` / `setdesired!`,
  * `ContactPoint{N}` of reference `src/contacts.jl:27-79` (`weight`, `maxnormalforce`, `disable!`, `isenabled`),
  * the OSQP settings the reference passes (`test/runtests.jl:35-43`, `notebooks/Standing controller.ipynb:66-71`),
  * `QPSolveFailure` of reference `src/exceptions.jl:1-11`.

Julia's `f!(x, ...)` becomes a method / function without the bang.  These objects only *describe* the problem; all
per-tick numerics run in the CUDA library behind the C ABI (see controller.py).
"""
from __future__ import annotations

from dataclasses import dataclass, field
from typing import List, Optional, Tuple, Union

import numpy as np

from .mechanism import Mechanism, REVOLUTE

# task kinds / modes -- values are part of the C ABI (include/qpcontrol_b200.h)
SPATIAL, ANGULAR, LINEAR, POINT, JOINT, MOMENTUM_RATE, LINEAR_MOMENTUM_RATE = range(7)
HARD, SCALAR_WEIGHT, MATRIX_WEIGHT = range(3)

# per-instance solver status -- OSQP's own codes
SOLVED, SOLVED_INACCURATE = 1, 2
PRIMAL_INFEASIBLE_INACCURATE, DUAL_INFEASIBLE_INACCURATE = 3, 4
MAX_ITER_REACHED, PRIMAL_INFEASIBLE, DUAL_INFEASIBLE, NON_FINITE, UNSOLVED = -2, -3, -4, -8, -10

_STATUS_NAMES = {1: "OPTIMAL", 2: "ALMOST_OPTIMAL", 3: "ALMOST_INFEASIBLE", 4: "ALMOST_DUAL_INFEASIBLE",
                 -2: "ITERATION_LIMIT", -3: "INFEASIBLE", -4: "DUAL_INFEASIBLE", -8: "NUMERICAL_ERROR",
                 -10: "OPTIMIZE_NOT_CALLED"}


class QPSolveFailure(Exception):
    """reference `src/exceptions.jl:1-11`; raised by `checkstatus` (reference `src/lowlevel/momentum.jl:83-91`) for
    the first instance whose status is neither SOLVED nor SOLVED_INACCURATE."""

    def __init__(self, instance: int, status: int):
        self.instance = instance
        self.status = status
        self.terminationstatus = _STATUS_NAMES.get(status, str(status))
        super().__init__(f"QP solve unsuccessful.\n    Instance: {instance}\n    Termination status: "
                         f"{self.terminationstatus}")


def checkstatus(status: np.ndarray) -> None:
    """(OPTIMAL, FEASIBLE_POINT) or (ALMOST_OPTIMAL, UNKNOWN_RESULT_STATUS) are accepted, anything else throws
    (reference `src/lowlevel/momentum.jl:83-91`)."""
    bad = np.flatnonzero((status != SOLVED) & (status != SOLVED_INACCURATE))
    if bad.size:
        raise QPSolveFailure(int(bad[0]), int(status[bad[0]]))


@dataclass
class OSQPSettings:
    """OSQP 0.5.x defaults with the fields the reference overrides."""
    rho: float = 0.1
    sigma: float = 1e-6
    alpha: float = 1.6
    eps_abs: float = 1e-3
    eps_rel: float = 1e-3
    eps_prim_inf: float = 1e-4
    eps_dual_inf: float = 1e-4
    max_iter: int = 4000
    scaling: int = 10
    adaptive_rho: int = 1
    adaptive_rho_interval: int = 25
    adaptive_rho_tolerance: float = 5.0
    check_termination: int = 25
    warm_start: int = 0

    @staticmethod
    def test_suite() -> "OSQPSettings":
        """`defaultoptimizer()` of reference `test/runtests.jl:35-43`."""
        return OSQPSettings(eps_abs=1e-8, eps_rel=1e-16, max_iter=20000, adaptive_rho_interval=25)

    @staticmethod
    def standing_notebook() -> "OSQPSettings":
        """reference `notebooks/Standing controller.ipynb:66-71`."""
        return OSQPSettings(eps_abs=1e-5, eps_rel=1e-5, max_iter=5000, adaptive_rho_interval=25)

    @staticmethod
    def acrobot_notebook() -> "OSQPSettings":
        """reference `notebooks/PointAccelerationTask Demo.ipynb:77-82`."""
        return OSQPSettings(eps_abs=1e-8, eps_rel=1e-8, max_iter=10000, adaptive_rho_interval=25)


class AbstractMotionTask:
    kind: int = -1
    source: int = -1
    target: int = -1
    frame: int = -1
    joint: int = -1
    point: Tuple[float, float, float] = (0.0, 0.0, 0.0)

    def __init__(self, dimension: int):
        self.dimension = dimension
        self.desired = np.zeros(dimension)

    def setdesired(self, desired) -> None:
        """`setdesired!`: the value used for every instance that is not given a per-instance desired."""
        d = np.asarray(desired, dtype=np.float64).reshape(-1)
        if d.shape != (self.dimension,):
            raise ValueError(f"desired has dimension {d.shape}, task has {self.dimension}")
        self.desired = d.copy()


class _PathTask(AbstractMotionTask):
    def __init__(self, mechanism: Mechanism, source: int, target: int, frame: Optional[int], dimension: int):
        super().__init__(dimension)
        self.source, self.target = int(source), int(target)
        # default frame: the target body's frame (tasks.jl:11,55,94)
        self.frame = int(target) if frame is None else int(frame)
        for b in (self.source, self.target, self.frame):
            if not -1 <= b < mechanism.nb:
                raise ValueError(f"body index {b} out of range")


class SpatialAccelerationTask(_PathTask):
    """Desired spatial acceleration (angular; linear) of `target` w.r.t. `source`, expressed in `frame`."""
    kind = SPATIAL

    def __init__(self, mechanism, source, target, frame=None):
        super().__init__(mechanism, source, target, frame, 6)


class AngularAccelerationTask(_PathTask):
    kind = ANGULAR

    def __init__(self, mechanism, source, target, frame=None):
        super().__init__(mechanism, source, target, frame, 3)


class LinearAccelerationTask(_PathTask):
    kind = LINEAR

    def __init__(self, mechanism, source, target, frame=None):
        super().__init__(mechanism, source, target, frame, 3)


class PointAccelerationTask(_PathTask):
    """Acceleration of a point fixed in `target` (given in the target body frame), expressed in the `source` (base)
    body frame (tasks.jl:131-142)."""
    kind = POINT

    def __init__(self, mechanism, source, target, point):
        super().__init__(mechanism, source, target, source, 3)
        self.point = tuple(float(x) for x in point)


class JointAccelerationTask(AbstractMotionTask):
    kind = JOINT

    def __init__(self, mechanism: Mechanism, joint: int):
        super().__init__(int(mechanism.nvj[joint]))
        self.joint = int(joint)


class MomentumRateTask(AbstractMotionTask):
    """Desired rate of centroidal momentum (angular; linear) in the centroidal frame (world axes at the CoM)."""
    kind = MOMENTUM_RATE

    def __init__(self, mechanism: Mechanism):
        super().__init__(6)


class LinearMomentumRateTask(AbstractMotionTask):
    kind = LINEAR_MOMENTUM_RATE

    def __init__(self, mechanism: Mechanism):
        super().__init__(3)


class ContactPoint:
    """`ContactPoint{N}` (contacts.jl:27-70): created disabled (`weight = maxnormalforce = 0`, contacts.jl:50)."""

    def __init__(self, body: int, position, normal, mu: float, N: int):
        self.body = int(body)
        self.position = np.asarray(position, dtype=np.float64).reshape(3).copy()
        self.normal = np.asarray(normal, dtype=np.float64).reshape(3).copy()
        self.mu = float(mu)
        self.N = N
        self.weight = 0.0
        self.maxnormalforce = 0.0

    def disable(self) -> None:  # contacts.jl:72
        self.maxnormalforce = 0.0

    def isenabled(self) -> bool:  # contacts.jl:73
        return self.maxnormalforce > 0


@dataclass
class TaskEntry:
    task: AbstractMotionTask
    mode: int
    weight: float = 0.0
    W: Optional[np.ndarray] = None


@dataclass
class StandingSpec:
    """Constants of the reference's `StandingController` (standing.jl:1-16)."""
    linmom_task: int
    pelvis_task: int
    pelvis_body: int
    joint_tasks: List[int]
    joints: List[int]
    joint_kp: np.ndarray
    joint_kd: np.ndarray
    joint_ref: np.ndarray
    com_kp: float
    com_kd: float
    pelvis_kp: float
    pelvis_kd: float
    comref: np.ndarray


@dataclass
class SE3PDSpec:
    """An `SE3PDController` (se3pdcontroller.jl:1-11) bound to the SpatialAccelerationTask whose desired it produces."""
    task: int
    controller: object  # qpcontrol.jl_b200.se3pd.SE3PDController


@dataclass
class Program:
    """Everything `addtask!` / `addcontact!` / `regularize!` recorded, in call order (the order defines the lifted
    QP's variable and row order in the reference, momentum.jl:28,123 and contacts.jl:46-48)."""
    mechanism: Mechanism
    N: int
    floating_body: int  # successor body of the floating joint, -1 if fixed base
    settings: OSQPSettings
    events: List[Tuple[str, int]] = field(default_factory=list)
    tasks: List[TaskEntry] = field(default_factory=list)
    contacts: List[ContactPoint] = field(default_factory=list)
    reg: Optional[np.ndarray] = None
    standing: Optional[StandingSpec] = None
    se3pd: List[SE3PDSpec] = field(default_factory=list)

    def __post_init__(self):
        if self.reg is None:
            self.reg = np.zeros(self.mechanism.nv)

    @property
    def ndes(self) -> int:
        return sum(t.task.dimension for t in self.tasks)

    def des_offsets(self) -> List[int]:
        off, out = 0, []
        for t in self.tasks:
            out.append(off)
            off += t.task.dimension
        return out

    def default_desired(self) -> np.ndarray:
        if not self.tasks:
            return np.zeros(0)
        return np.concatenate([t.task.desired for t in self.tasks])
