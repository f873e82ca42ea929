"""qpcontrol.jl_b200 -- B200-native batched implementation of QPControl.jl's per-timestep control loop.

The directory name contains a dot, so it is loaded through `qpc_loader.load()` (repo root) under the module name
`qpcontrol_jl_b200`.  Host-side description objects live in `mechanism.py` / `program.py` / `controller.py`; the
numerics are hand-written sm_100a CUDA kernels in `csrc/` behind the C ABI declared in `include/qpcontrol_b200.h`.
"""
from .mechanism import (FIXED, PRISMATIC, QUAT_FLOATING, REVOLUTE, Mechanism, acrobot, atlas_like,
                        atlas_nominal_configuration, rand_floating_humanoid, rand_tree)
from .program import (ANGULAR, HARD, JOINT, LINEAR, LINEAR_MOMENTUM_RATE, MATRIX_WEIGHT, MOMENTUM_RATE, POINT,
                      SCALAR_WEIGHT, SPATIAL, AbstractMotionTask, AngularAccelerationTask, ContactPoint,
                      JointAccelerationTask, LinearAccelerationTask, LinearMomentumRateTask, MomentumRateTask,
                      OSQPSettings, PointAccelerationTask, Program, QPSolveFailure, SpatialAccelerationTask,
                      checkstatus)
from .controller import BatchResult, MomentumBasedController, StandingController, center_of_mass_host
from .urdf import parse_urdf
from .se3pd import PDGains, SE3PDController, SE3PDGains
from . import scenarios, se3pd, sharding, trajectories
