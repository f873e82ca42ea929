"""URDF -> flat `Mechanism` (the on-disk model format on the caller's side of the hot path).

The reference loads its robots with `RigidBodyDynamics.parse_urdf` followed by `remove_fixed_tree_joints!`
(`notebooks/Standing controller.ipynb:39-40`: AtlasRobot.mechanism(); `PointAccelerationTask Demo.ipynb:43`:
parse_urdf(Float64, "Acrobot.urdf")).  This module does the same two steps on the host, in Python, and produces the
struct-of-arrays tree the C ABI takes (`qpc_mechanism_create`):

* links: `<inertial>` origin (xyz, rpy), mass, inertia about the centre of mass in the inertial frame's axes;
* joints: `revolute` / `continuous` -> REVOLUTE, `prismatic` -> PRISMATIC, `floating` -> QUAT_FLOATING, `fixed` -> merged
  into the parent (composite inertia, children re-attached), anything else rejected;
* the root link is welded to the world (`floating=False`, parse_urdf's default) or attached with a quaternion
  floating joint (`floating=True`, what AtlasRobot / ValkyrieRobot do);
* bodies are renumbered breadth-first from the root so that parent[i] < i.

Conventions are URDF's, which coincide with the `Mechanism` ones: a body's frame is its link frame (the frame after its
joint), `<joint><origin>` is the pose of that frame in the parent link's frame, `<axis>` is expressed in the child frame.
"""
from __future__ import annotations

import xml.etree.ElementTree as ET
from typing import Dict, List, Optional

import numpy as np

from .mechanism import FIXED, PRISMATIC, QUAT_FLOATING, REVOLUTE, Mechanism

_JOINT_TYPES = {"revolute": REVOLUTE, "continuous": REVOLUTE, "prismatic": PRISMATIC, "floating": QUAT_FLOATING,
                "fixed": FIXED}


def _floats(text: Optional[str], n: int, default: float = 0.0) -> np.ndarray:
    if text is None:
        return np.full(n, default)
    v = np.array([float(t) for t in text.split()], dtype=np.float64)
    if v.shape != (n,):
        raise ValueError(f"expected {n} numbers, got '{text}'")
    return v


def rpy_to_rot(rpy) -> np.ndarray:
    """URDF fixed-axis roll-pitch-yaw: R = Rz(yaw) Ry(pitch) Rx(roll)."""
    r, p, y = rpy
    cr, sr, cp, sp, cy, sy = np.cos(r), np.sin(r), np.cos(p), np.sin(p), np.cos(y), np.sin(y)
    Rx = np.array([[1, 0, 0], [0, cr, -sr], [0, sr, cr]])
    Ry = np.array([[cp, 0, sp], [0, 1, 0], [-sp, 0, cp]])
    Rz = np.array([[cy, -sy, 0], [sy, cy, 0], [0, 0, 1]])
    return Rz @ Ry @ Rx


def _origin(elem) -> (np.ndarray, np.ndarray):
    o = None if elem is None else elem.find("origin")
    if o is None:
        return np.eye(3), np.zeros(3)
    return rpy_to_rot(_floats(o.get("rpy"), 3)), _floats(o.get("xyz"), 3)


def _parallel_axis(m: float, d: np.ndarray) -> np.ndarray:
    return m * (np.dot(d, d) * np.eye(3) - np.outer(d, d))


class _Link:
    def __init__(self, name):
        self.name = name
        self.mass = 0.0
        self.com = np.zeros(3)
        self.inertia = np.zeros((3, 3))  # about the centre of mass, link axes
        self.children: List["_Joint"] = []
        self.parent_joint: Optional["_Joint"] = None

    def absorb(self, m2: float, c2: np.ndarray, I2: np.ndarray):
        """Composite of this link's inertia with (m2, c2, I2) given in this link's frame."""
        m1, c1, I1 = self.mass, self.com, self.inertia
        m = m1 + m2
        if m <= 0.0:
            return
        c = (m1 * c1 + m2 * c2) / m
        self.inertia = I1 + _parallel_axis(m1, c1 - c) + I2 + _parallel_axis(m2, c2 - c)
        self.mass, self.com = m, c


class _Joint:
    def __init__(self, name, jtype, parent, child, R, p, axis):
        self.name, self.jtype, self.parent, self.child = name, jtype, parent, child
        self.R, self.p, self.axis = R, p, axis


def parse_urdf(source: str, floating: bool = False, remove_fixed_joints: bool = True,
               gravity=(0.0, 0.0, -9.81)) -> Mechanism:
    """`source` is a file name or an XML string.  Returns the flat tree; raises ValueError on unsupported content."""
    root = ET.fromstring(source) if source.lstrip().startswith("<") else ET.parse(source).getroot()
    if root.tag != "robot":
        raise ValueError("not a URDF: root element is not <robot>")
    links: Dict[str, _Link] = {}
    for le in root.findall("link"):
        link = _Link(le.get("name"))
        ine = le.find("inertial")
        if ine is not None:
            R, p = _origin(ine)
            me = ine.find("mass")
            link.mass = float(me.get("value")) if me is not None else 0.0
            ie = ine.find("inertia")
            if ie is not None:
                g = lambda k: float(ie.get(k, "0"))  # noqa: E731
                I = np.array([[g("ixx"), g("ixy"), g("ixz")], [g("ixy"), g("iyy"), g("iyz")],
                              [g("ixz"), g("iyz"), g("izz")]])
                link.inertia = R @ I @ R.T
            link.com = p
        if link.name in links:
            raise ValueError(f"duplicate link '{link.name}'")
        links[link.name] = link
    for je in root.findall("joint"):
        t = je.get("type")
        if t not in _JOINT_TYPES:
            raise ValueError(f"joint '{je.get('name')}': unsupported type '{t}'")
        parent, child = je.find("parent").get("link"), je.find("child").get("link")
        if parent not in links or child not in links:
            raise ValueError(f"joint '{je.get('name')}' refers to an unknown link")
        R, p = _origin(je)
        ae = je.find("axis")
        axis = _floats(None if ae is None else ae.get("xyz"), 3)
        if ae is None:
            axis = np.array([1.0, 0.0, 0.0])  # URDF default
        j = _Joint(je.get("name"), _JOINT_TYPES[t], links[parent], links[child], R, p, axis)
        if j.child.parent_joint is not None:
            raise ValueError(f"link '{child}' has two parents: kinematic loops are not supported")
        j.child.parent_joint = j
        j.parent.children.append(j)
    roots = [l for l in links.values() if l.parent_joint is None]
    if len(roots) != 1:
        raise ValueError(f"expected exactly one root link, found {[l.name for l in roots]}")
    base = roots[0]

    # ---- remove_fixed_tree_joints!: merge fixed children into their parents, deepest first -------------------------
    if remove_fixed_joints:
        def merge(link: _Link):
            for j in list(link.children):
                merge(j.child)
            for j in list(link.children):
                if j.jtype != FIXED:
                    continue
                c = j.child
                link.absorb(c.mass, j.R @ c.com + j.p, j.R @ c.inertia @ j.R.T)
                link.children.remove(j)
                for jc in c.children:  # grandchildren hang off this link now
                    jc.R, jc.p = j.R @ jc.R, j.R @ jc.p + j.p
                    jc.parent = link
                    link.children.append(jc)
        merge(base)

    # ---- flatten breadth-first; the root link is welded to the world or floats ---------------------------------------
    names, jnames, parent, jtype, axis, XR, Xp, mass, com, inertia = [], [], [], [], [], [], [], [], [], []

    def emit(link: _Link, jname: str, pidx: int, jt: int, ax, R, p):
        names.append(link.name)
        jnames.append(jname)
        parent.append(pidx)
        jtype.append(jt)
        a = np.asarray(ax, dtype=np.float64)
        n = np.linalg.norm(a)
        axis.append(a / n if n > 0 else np.array([0.0, 0.0, 1.0]))
        XR.append(R)
        Xp.append(p)
        mass.append(link.mass)
        com.append(link.com)
        inertia.append(link.inertia)
        return len(names) - 1

    queue = []
    if floating:
        idx = emit(base, f"{base.name}_to_world", -1, QUAT_FLOATING, (0, 0, 1), np.eye(3), np.zeros(3))
        queue.append((base, idx, np.eye(3), np.zeros(3)))
    else:
        # welded root: its inertia does not move; its children attach to the world through the root's (identity) pose
        queue.append((base, -1, np.eye(3), np.zeros(3)))
    while queue:
        link, idx, Rw, pw = queue.pop(0)
        for j in link.children:
            if j.jtype == FIXED:
                if idx == -1:  # fixed chain hanging off a welded root that was not merged (remove_fixed_joints=False)
                    queue.append((j.child, -1, Rw @ j.R, Rw @ j.p + pw))
                    continue
                raise ValueError("fixed joints below a moving body need remove_fixed_joints=True")
            R, p = (Rw @ j.R, Rw @ j.p + pw) if idx == -1 else (j.R, j.p)
            cidx = emit(j.child, j.name, idx, j.jtype, j.axis, R, p)
            queue.append((j.child, cidx, np.eye(3), np.zeros(3)))
    if not names:
        raise ValueError("URDF has no moving bodies")
    return Mechanism(names, jnames, np.array(parent), np.array(jtype), np.array(axis), np.array(XR), np.array(Xp),
                     np.array(mass), np.array(com), np.array(inertia), gravity=np.asarray(gravity, dtype=np.float64))
