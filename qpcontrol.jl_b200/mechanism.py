"""Flat (POD) description of a kinematic tree + the model generators the hot path is exercised on.

The reference never stores a mechanism itself: it takes a `RigidBodyDynamics.Mechanism`
(reference `src/lowlevel/momentum.jl:15-17`) built from a URDF (`notebooks/Standing controller.ipynb:39-40`,
`notebooks/PointAccelerationTask Demo.ipynb:43`) or from `rand_tree_mechanism` (`test/tasks.jl:3`,
`test/controller.jl:70`).  None of those packages/URDFs exist offline, so this module provides

* `Mechanism`        -- struct-of-arrays tree, topologically sorted (parent[i] < i, -1 = world),
* `acrobot()`        -- the Acrobot of the PointAccelerationTask demo,
* `atlas_like()`     -- a 36-DoF humanoid with the Atlas v5 topology and joint names used by the
                        standing-controller notebook (link data is a plausible
                        stand-in, real Atlas is a data swap),
* `rand_tree()`      -- analogue of `RigidBodyDynamics.rand_tree_mechanism` for the invariant tests,
* `rand_floating_humanoid()` -- stand-in for the Valkyrie cases of `test/controller.jl:128-285`.

Conventions (RigidBodyDynamics 2.2.0): one frame per body (the frame after its joint); `X_tree[i]` maps the
frame before joint i to the parent body's frame; spatial vectors are (angular; linear); the floating joint has
q = (w, x, y, z, px, py, pz) and v = (omega; v) in the body frame.
"""
from __future__ import annotations

from dataclasses import dataclass, field
from typing import Dict, List, Optional, Sequence

import numpy as np

REVOLUTE, PRISMATIC, QUAT_FLOATING, FIXED = 0, 1, 2, 3
_NQ = {REVOLUTE: 1, PRISMATIC: 1, QUAT_FLOATING: 7, FIXED: 0}
_NV = {REVOLUTE: 1, PRISMATIC: 1, QUAT_FLOATING: 6, FIXED: 0}


def _rot_axis_angle(axis, angle):
    axis = np.asarray(axis, dtype=np.float64)
    axis = axis / np.linalg.norm(axis)
    K = np.array([[0, -axis[2], axis[1]], [axis[2], 0, -axis[0]], [-axis[1], axis[0], 0]])
    return np.eye(3) + np.sin(angle) * K + (1 - np.cos(angle)) * (K @ K)


@dataclass
class Mechanism:
    """Topologically sorted tree of `nb` bodies (world excluded). All arrays are float64 / int32, C-contiguous."""

    names: List[str]
    joint_names: List[str]
    parent: np.ndarray  # [nb] int32, -1 = world
    jtype: np.ndarray  # [nb] int32
    axis: np.ndarray  # [nb,3] joint axis in the body frame (unused for floating/fixed)
    X_R: np.ndarray  # [nb,3,3] rotation of joint_to_parent
    X_p: np.ndarray  # [nb,3]   translation of joint_to_parent
    mass: np.ndarray  # [nb]
    com: np.ndarray  # [nb,3] centre of mass in the body frame
    inertia_com: np.ndarray  # [nb,3,3] rotational inertia about the centre of mass, body axes
    gravity: np.ndarray = field(default_factory=lambda: np.array([0.0, 0.0, -9.81]))
    contact_points: Dict[int, np.ndarray] = field(default_factory=dict)  # body -> [k,3] (body frame)
    contact_mu: float = 0.8

    def __post_init__(self):
        self.parent = np.ascontiguousarray(self.parent, dtype=np.int32)
        self.jtype = np.ascontiguousarray(self.jtype, dtype=np.int32)
        for name in ("axis", "X_R", "X_p", "mass", "com", "inertia_com", "gravity"):
            setattr(self, name, np.ascontiguousarray(getattr(self, name), dtype=np.float64))
        nb = self.nb
        assert self.parent.shape == (nb,) and np.all(self.parent < np.arange(nb)) and np.all(self.parent >= -1)
        self.nqj = np.array([_NQ[int(t)] for t in self.jtype], dtype=np.int32)
        self.nvj = np.array([_NV[int(t)] for t in self.jtype], dtype=np.int32)
        self.qoff = np.concatenate([[0], np.cumsum(self.nqj)[:-1]]).astype(np.int32)
        self.voff = np.concatenate([[0], np.cumsum(self.nvj)[:-1]]).astype(np.int32)

    # -- sizes ------------------------------------------------------------------------------------------------
    @property
    def nb(self) -> int:
        return len(self.names)

    @property
    def nq(self) -> int:
        return int(self.nqj.sum())

    @property
    def nv(self) -> int:
        return int(self.nvj.sum())

    @property
    def total_mass(self) -> float:
        return float(self.mass.sum())

    # -- lookups (mirror findbody / findjoint of RigidBodyDynamics) ----------------------------------------------
    def findbody(self, name: str) -> int:
        if name == "world":
            return -1
        return self.names.index(name)

    def findjoint(self, name: str) -> int:
        """Joints are identified with their successor body index."""
        return self.joint_names.index(name)

    def velocity_range(self, joint: int) -> range:
        return range(int(self.voff[joint]), int(self.voff[joint] + self.nvj[joint]))

    def configuration_range(self, joint: int) -> range:
        return range(int(self.qoff[joint]), int(self.qoff[joint] + self.nqj[joint]))

    def depth(self) -> np.ndarray:
        d = np.zeros(self.nb, dtype=np.int32)
        for i in range(self.nb):
            d[i] = 0 if self.parent[i] < 0 else d[self.parent[i]] + 1
        return d

    def ancestors(self, body: int) -> List[int]:
        """body, parent(body), ... up to (excluding) world."""
        out = []
        while body >= 0:
            out.append(body)
            body = int(self.parent[body])
        return out

    def path(self, source: int, target: int):
        """Joints between two bodies as (joint, sign) pairs: -1 while climbing from `source` to the lowest
        common ancestor, +1 while descending to `target` (RigidBodyDynamics `path(mechanism, source, target)`)."""
        up = self.ancestors(source)
        down = self.ancestors(target)
        common = set(up) & set(down)
        out = [(b, -1) for b in up if b not in common]
        out += [(b, +1) for b in reversed(down) if b not in common]
        return out

    def inertia_origin(self) -> np.ndarray:
        """Rotational inertia about the body-frame origin (the `moment` of an RBD SpatialInertia)."""
        out = np.empty_like(self.inertia_com)
        for i in range(self.nb):
            c = self.com[i]
            out[i] = self.inertia_com[i] + self.mass[i] * (np.dot(c, c) * np.eye(3) - np.outer(c, c))
        return out

    def zero_configuration(self) -> np.ndarray:
        q = np.zeros(self.nq)
        for i in range(self.nb):
            if self.jtype[i] == QUAT_FLOATING:
                q[self.qoff[i]] = 1.0
        return q

    def rand_configuration(self, rng: np.random.Generator) -> np.ndarray:
        q = np.zeros(self.nq)
        for i in range(self.nb):
            o = self.qoff[i]
            if self.jtype[i] == QUAT_FLOATING:
                quat = rng.standard_normal(4)
                q[o:o + 4] = quat / np.linalg.norm(quat)
                q[o + 4:o + 7] = rng.uniform(-1, 1, 3)
            elif self.jtype[i] == REVOLUTE:
                q[o] = rng.uniform(-np.pi, np.pi)
            elif self.jtype[i] == PRISMATIC:
                q[o] = rng.uniform(-1, 1)
        return q


# ---------------------------------------------------------------------------------------------------------------
# builders
# ---------------------------------------------------------------------------------------------------------------
class _Builder:
    def __init__(self):
        self.names, self.joint_names, self.parent, self.jtype = [], [], [], []
        self.axis, self.X_R, self.X_p, self.mass, self.com, self.inertia = [], [], [], [], [], []

    def add(self, name, joint_name, parent, jtype, axis=(0, 0, 1), origin=(0, 0, 0), rot=None,
            mass=1.0, com=(0, 0, 0), inertia=None):
        pidx = -1 if parent is None else self.names.index(parent)
        self.names.append(name)
        self.joint_names.append(joint_name)
        self.parent.append(pidx)
        self.jtype.append(jtype)
        a = np.asarray(axis, dtype=np.float64)
        self.axis.append(a / max(np.linalg.norm(a), 1e-300))
        self.X_R.append(np.eye(3) if rot is None else np.asarray(rot, dtype=np.float64))
        self.X_p.append(np.asarray(origin, dtype=np.float64))
        self.mass.append(float(mass))
        self.com.append(np.asarray(com, dtype=np.float64))
        if inertia is None:
            inertia = np.eye(3) * 0.01 * mass
        inertia = np.asarray(inertia, dtype=np.float64)
        if inertia.ndim == 1:
            inertia = np.diag(inertia)
        self.inertia.append(inertia)
        return len(self.names) - 1

    def build(self, **kw) -> Mechanism:
        return Mechanism(self.names, self.joint_names, np.array(self.parent), np.array(self.jtype),
                         np.array(self.axis), np.array(self.X_R), np.array(self.X_p), np.array(self.mass),
                         np.array(self.com), np.array(self.inertia), **kw)


def acrobot() -> Mechanism:
    """RigidBodyDynamics `test/urdf/Acrobot.urdf` as used by `notebooks/PointAccelerationTask Demo.ipynb:43`:
    two revolute joints about +y, base link welded to the world."""
    b = _Builder()
    b.add("upper_link", "shoulder", None, REVOLUTE, axis=(0, 1, 0), origin=(0, 0.15, 0), mass=1.0,
          com=(0, 0, -0.5), inertia=(0.083, 0.083, 0.001))
    b.add("lower_link", "elbow", "upper_link", REVOLUTE, axis=(0, 1, 0), origin=(0, 0.1, -1.0), mass=1.0,
          com=(0, 0, -1.0), inertia=(0.33, 0.33, 0.001))
    return b.build()


def _box_inertia(mass, sx, sy, sz):
    return mass / 12.0 * np.array([sy * sy + sz * sz, sx * sx + sz * sz, sx * sx + sy * sy])


def atlas_like() -> Mechanism:
    """Atlas-v5 topology: floating pelvis + 30 revolute joints, nq = 37, nv = 36, with the joint/body names the
    standing-controller notebook looks up (`pelvis_to_world`, `{l,r}_leg_{kny,hpy,aky}`, `pelvis`, `{l,r}_foot`;
    `notebooks/Standing controller.ipynb:85,111-113,131-132`) and 4 contact points per foot."""
    b = _Builder()
    b.add("pelvis", "pelvis_to_world", None, QUAT_FLOATING, mass=9.5, com=(0.011, 0, 0.027),
          inertia=_box_inertia(9.5, 0.25, 0.3, 0.2))
    b.add("ltorso", "back_bkz", "pelvis", REVOLUTE, axis=(0, 0, 1), origin=(-0.0125, 0, 0), mass=2.27,
          com=(-0.011, 0, 0.075), inertia=_box_inertia(2.27, 0.1, 0.15, 0.12))
    b.add("mtorso", "back_bky", "ltorso", REVOLUTE, axis=(0, 1, 0), origin=(0, 0, 0.162), mass=0.8,
          com=(-0.008, 0, 0.02), inertia=_box_inertia(0.8, 0.08, 0.12, 0.06))
    b.add("utorso", "back_bkx", "mtorso", REVOLUTE, axis=(1, 0, 0), origin=(0, 0, 0.05), mass=63.7,
          com=(-0.035, 0.0, 0.27), inertia=_box_inertia(63.7, 0.35, 0.45, 0.6))
    b.add("head", "neck_ry", "utorso", REVOLUTE, axis=(0, 1, 0), origin=(0.2546, 0, 0.6215), mass=1.42,
          com=(-0.075, 0, 0.034), inertia=_box_inertia(1.42, 0.15, 0.15, 0.15))
    for side, s in (("l", 1.0), ("r", -1.0)):
        b.add(f"{side}_clav", f"{side}_arm_shz", "utorso", REVOLUTE, axis=(0, 0, 1),
              origin=(0.1406, s * 0.2256, 0.4776), mass=4.47, com=(0, s * -0.048, -0.084),
              inertia=_box_inertia(4.47, 0.1, 0.15, 0.2))
        b.add(f"{side}_scap", f"{side}_arm_shx", f"{side}_clav", REVOLUTE, axis=(1, 0, 0),
              origin=(0, s * 0.11, -0.245), mass=3.9, com=(0, s * 0.075, 0.036),
              inertia=_box_inertia(3.9, 0.1, 0.2, 0.1))
        b.add(f"{side}_uarm", f"{side}_arm_ely", f"{side}_scap", REVOLUTE, axis=(0, 1, 0),
              origin=(0, s * 0.187, 0.016), mass=4.42, com=(0, s * 0.065, 0),
              inertia=_box_inertia(4.42, 0.1, 0.2, 0.1))
        b.add(f"{side}_larm", f"{side}_arm_elx", f"{side}_uarm", REVOLUTE, axis=(1, 0, 0),
              origin=(0, s * 0.119, 0.0092), mass=3.39, com=(0, s * 0.065, 0),
              inertia=_box_inertia(3.39, 0.1, 0.2, 0.1))
        b.add(f"{side}_ufarm", f"{side}_arm_wry", f"{side}_larm", REVOLUTE, axis=(0, 1, 0),
              origin=(0, s * 0.187, -0.0092), mass=2.51, com=(0, s * 0.04, 0),
              inertia=_box_inertia(2.51, 0.08, 0.15, 0.08))
        b.add(f"{side}_lfarm", f"{side}_arm_wrx", f"{side}_ufarm", REVOLUTE, axis=(1, 0, 0),
              origin=(0, s * 0.119, 0.0092), mass=0.85, com=(0, s * 0.03, 0),
              inertia=_box_inertia(0.85, 0.06, 0.1, 0.06))
        b.add(f"{side}_hand", f"{side}_arm_wry2", f"{side}_lfarm", REVOLUTE, axis=(0, 1, 0),
              origin=(0, s * 0.05, 0), mass=2.26, com=(0, s * 0.09, 0),
              inertia=_box_inertia(2.26, 0.1, 0.2, 0.08))
    for side, s in (("l", 1.0), ("r", -1.0)):
        b.add(f"{side}_uglut", f"{side}_leg_hpz", "pelvis", REVOLUTE, axis=(0, 0, 1),
              origin=(0, s * 0.089, 0), mass=1.96, com=(0.005, s * -0.003, 0.032),
              inertia=_box_inertia(1.96, 0.1, 0.1, 0.1))
        b.add(f"{side}_lglut", f"{side}_leg_hpx", f"{side}_uglut", REVOLUTE, axis=(1, 0, 0),
              origin=(0, 0, 0), mass=2.9, com=(0.013, s * 0.017, -0.031),
              inertia=_box_inertia(2.9, 0.12, 0.12, 0.12))
        b.add(f"{side}_uleg", f"{side}_leg_hpy", f"{side}_lglut", REVOLUTE, axis=(0, 1, 0),
              origin=(0.05, s * 0.0225, -0.066), mass=8.2, com=(0, 0, -0.21),
              inertia=_box_inertia(8.2, 0.15, 0.15, 0.4))
        b.add(f"{side}_lleg", f"{side}_leg_kny", f"{side}_uleg", REVOLUTE, axis=(0, 1, 0),
              origin=(-0.05, 0, -0.374), mass=4.52, com=(0.001, 0, -0.187),
              inertia=_box_inertia(4.52, 0.1, 0.1, 0.4))
        b.add(f"{side}_talus", f"{side}_leg_aky", f"{side}_lleg", REVOLUTE, axis=(0, 1, 0),
              origin=(0, 0, -0.422), mass=0.125, com=(0, 0, 0),
              inertia=(1.0e-5, 1.0e-5, 1.0e-5))
        b.add(f"{side}_foot", f"{side}_leg_akx", f"{side}_talus", REVOLUTE, axis=(1, 0, 0),
              origin=(0, 0, 0), mass=2.41, com=(0.027, 0, -0.067),
              inertia=_box_inertia(2.41, 0.26, 0.13, 0.05))
    mech = b.build()
    # heel pair at x = -0.0876, toe pair at x = 0.1728, sole at z = -0.07645, mu = 0.8
    for side in ("l", "r"):
        mech.contact_points[mech.findbody(f"{side}_foot")] = np.array([
            [-0.0876, 0.0626, -0.07645], [-0.0876, -0.0626, -0.07645],
            [0.1728, 0.0626, -0.07645], [0.1728, -0.0626, -0.07645]])
    mech.contact_mu = 0.8
    return mech


def atlas_nominal_configuration(mech: Mechanism) -> np.ndarray:
    """`initialize!` of `notebooks/Standing controller.ipynb:105-121`."""
    q = mech.zero_configuration()
    kneebend, hipbendextra = 1.1, 0.1
    for side in ("l", "r"):
        q[mech.qoff[mech.findjoint(f"{side}_leg_kny")]] = kneebend
        q[mech.qoff[mech.findjoint(f"{side}_leg_hpy")]] = -kneebend / 2 + hipbendextra
        q[mech.qoff[mech.findjoint(f"{side}_leg_aky")]] = -kneebend / 2 - hipbendextra
    fj = mech.findjoint("pelvis_to_world")
    q[mech.qoff[fj]:mech.qoff[fj] + 7] = [1, 0, 0, 0, 0, 0, 0.85]
    return q


def rand_tree(rng: np.random.Generator, jtypes: Sequence[int], floating: bool = False) -> Mechanism:
    """Analogue of `RigidBodyDynamics.rand_tree_mechanism(Float64, jointtypes...)` (`test/tasks.jl:3`,
    `test/controller.jl:70`): each new body is attached to a uniformly random existing body (or the world) through
    a joint with a random axis and a random joint-to-parent pose, and gets a random physically valid inertia."""
    b = _Builder()
    names = []
    if floating:
        b.add("base", "base_to_world", None, QUAT_FLOATING, mass=5.0 + rng.uniform(0, 5),
              com=rng.uniform(-0.1, 0.1, 3), inertia=_rand_inertia(rng, 5.0))
        names.append("base")
    for k, jt in enumerate(jtypes):
        cands: List[Optional[str]] = list(names) if floating else [None] + list(names)
        parent = cands[int(rng.integers(len(cands)))]
        mass = rng.uniform(0.5, 3.0)
        axis = rng.standard_normal(3)
        rot = _rot_axis_angle(rng.standard_normal(3), rng.uniform(-np.pi, np.pi))
        name = f"body{k + 1}"
        b.add(name, f"joint{k + 1}", parent, jt, axis=axis, origin=rng.uniform(-0.5, 0.5, 3), rot=rot, mass=mass,
              com=rng.uniform(-0.2, 0.2, 3), inertia=_rand_inertia(rng, mass))
        names.append(name)
    return b.build()


def _rand_inertia(rng, mass):
    A = rng.standard_normal((3, 3))
    Q, _ = np.linalg.qr(A)
    # principal moments of a box satisfy the triangle inequality
    d = _box_inertia(mass, *rng.uniform(0.05, 0.4, 3))
    return Q @ np.diag(d) @ Q.T


def rand_floating_humanoid(rng: np.random.Generator) -> Mechanism:
    """Stand-in for `ValkyrieRobot.Valkyrie()` in `test/controller.jl:128-285`: the Atlas-like tree with randomly
    perturbed link data, so the free-fall / achievable-momentum-rate / spatial-acceleration invariants are exercised
    on a second floating-base humanoid."""
    mech = atlas_like()
    mech.mass *= rng.uniform(0.7, 1.3, mech.nb)
    mech.com += rng.uniform(-0.02, 0.02, mech.com.shape)
    mech.X_p += rng.uniform(-0.01, 0.01, mech.X_p.shape)
    return mech
