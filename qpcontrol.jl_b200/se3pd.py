"""`SE3PDController` (reference src/lowlevel/se3pdcontroller.jl:1-18) restated for the host side, batched over instances.

SURVEY.md 8(f) rank 3, second half.  The reference's functor is four lines (se3pdcontroller.jl:13-18):

    Href, Tref, Tdref = trajectory(t, Val(2))                      # pose, twist, spatial acceleration of the reference
    H = relative_transform(state, Href.from, Href.to)              # actual pose of `body` in `base`
    T = transform(relative_twist(state, body, base), inv(transform_to_root(state, Tref.frame)))   # twist in body frame
    return Tdref + pd(gains, H, Href, T, Tref)

`pd(::SE3PDGains, ...)` is RigidBodyDynamics.PDControl's (RigidBodyDynamics 2.2.0, absent from the container -- restated
from its published algorithm, [dep-memory], parity unpinned like the rest of the third-party semantics, SURVEY.md
appendix B): the double-geodesic law of Bullo & Murray, "Proportional derivative (PD) control on the Euclidean group"
(1995), theorem 12, with gains expressed in the actual body frame:

    e  = inv(x_des) * x                  (pose of the body frame in the desired body frame: R_e, p_e)
    ed = v - inv(e) * v_des              (desired twist re-expressed in the actual body frame, subtracted)
    angular = -K_ang rotvec(R_e) - D_ang ed_angular
    linear  = -K_lin R_e' p_e      - D_lin ed_linear

The result (angular; linear), expressed in the body frame, is what `setdesired!(::SpatialAccelerationTask, ...)` takes
(tasks.jl:23-29) and goes into `qpc_batch_in.desired` for that task.  Everything here is setup-rate host arithmetic
(numpy, leading batch dimension); the per-tick device path is unchanged.
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import Tuple, Union

import numpy as np

from .mechanism import PRISMATIC, QUAT_FLOATING, REVOLUTE, Mechanism
from .trajectories import Constant, Interpolated, Piecewise, PointTrajectory, Polynomial, quat_to_rot

__all__ = ["PDGains", "SE3PDGains", "pd", "pd_se3", "rotation_vector", "body_pose_and_twist", "SE3PDController",
           "compile_trajectory", "gains_matrix"]

MAX_PIECES = 6  # QPC_MAX_PIECES of include/qpcontrol_b200.h

Gain = Union[float, np.ndarray]


def _apply(gain: Gain, x: np.ndarray) -> np.ndarray:
    gain = np.asarray(gain, dtype=np.float64)
    if gain.ndim == 0:
        return gain * x
    if gain.ndim == 1:  # diagonal gain
        return gain * x
    return np.einsum("...ij,...j->...i", gain, x)


@dataclass
class PDGains:
    """PDControl.PDGains: proportional gain `k` and derivative gain `d` (scalar, per-axis vector or 3 x 3 matrix)."""

    k: Gain
    d: Gain


def pd(gains: PDGains, e, ed):
    """PDControl.pd(gains, e, ė) = -k e - d ė (the form standing.jl:66,75,83 uses for its CoM / pelvis / joint laws)."""
    return -_apply(gains.k, np.asarray(e, dtype=np.float64)) - _apply(gains.d, np.asarray(ed, dtype=np.float64))


@dataclass
class SE3PDGains:
    """PDControl.SE3PDGains: angular and linear gains, both expressed in the (actual) body frame."""

    angular: PDGains
    linear: PDGains


def rotation_vector(R: np.ndarray) -> np.ndarray:
    """Rotation matrix -> rotation vector (Rotations.RotationVector; the same map standing.jl:73-75 applies to the
    pelvis orientation), via the unit quaternion so that small angles and angles near pi are both well conditioned."""
    R = np.asarray(R, dtype=np.float64)
    m00, m11, m22 = R[..., 0, 0], R[..., 1, 1], R[..., 2, 2]
    # quaternion with the largest component computed from the diagonal (Shepperd's method)
    cand = np.stack([1 + m00 + m11 + m22, 1 + m00 - m11 - m22, 1 - m00 + m11 - m22, 1 - m00 - m11 + m22], axis=-1)
    which = np.argmax(cand, axis=-1)
    a, b, c = R[..., 2, 1] - R[..., 1, 2], R[..., 0, 2] - R[..., 2, 0], R[..., 1, 0] - R[..., 0, 1]
    s01, s02, s12 = R[..., 0, 1] + R[..., 1, 0], R[..., 0, 2] + R[..., 2, 0], R[..., 1, 2] + R[..., 2, 1]
    t = np.take_along_axis(cand, which[..., None], axis=-1)[..., 0]
    opts = np.stack([np.stack([t, a, b, c], -1), np.stack([a, t, s01, s02], -1), np.stack([b, s01, t, s12], -1),
                     np.stack([c, s02, s12, t], -1)], axis=-2)  # [..., 4 options, 4 components]
    quat = np.take_along_axis(opts, which[..., None, None], axis=-2)[..., 0, :]
    quat = quat / np.linalg.norm(quat, axis=-1, keepdims=True)
    quat = np.where(quat[..., :1] < 0, -quat, quat)
    w, xyz = quat[..., 0], quat[..., 1:]
    s = np.linalg.norm(xyz, axis=-1)
    angle = 2 * np.arctan2(s, w)
    small = s < 1e-8
    scale = np.where(small, 2.0, angle / np.where(small, 1.0, s))
    return xyz * scale[..., None]


def pd_se3(gains: SE3PDGains, R, p, R_des, p_des, twist, twist_des) -> np.ndarray:
    """`pd(gains::SE3PDGains, x, xdes, v, vdes)` with the double-geodesic method (module docstring).

    x = (R, p), x_des = (R_des, p_des): poses of the body / desired body frame in the base frame; `twist` = (omega; nu)
    of the body w.r.t. base expressed in the BODY frame, `twist_des` the reference twist expressed in the DESIRED body
    frame (what `SE3Trajectory` returns).  Returns the (angular; linear) PD spatial acceleration in the body frame."""
    R, p, R_des, p_des = (np.asarray(a, dtype=np.float64) for a in (R, p, R_des, p_des))
    twist, twist_des = np.asarray(twist, dtype=np.float64), np.asarray(twist_des, dtype=np.float64)
    Rdt = np.swapaxes(R_des, -1, -2)
    R_e = Rdt @ R  # body -> desired body
    p_e = np.einsum("...ij,...j->...i", Rdt, p - p_des)
    R_et = np.swapaxes(R_e, -1, -2)
    rot = lambda M, x: np.einsum("...ij,...j->...i", M, x)  # noqa: E731
    # inv(e) maps the desired body frame to the body frame: rotation R_e', translation -R_e' p_e
    w_des = rot(R_et, twist_des[..., :3])
    nu_des = rot(R_et, twist_des[..., 3:]) + np.cross(-rot(R_et, p_e), w_des)
    ed_ang = twist[..., :3] - w_des
    ed_lin = twist[..., 3:] - nu_des
    ang = pd(gains.angular, rotation_vector(R_e), ed_ang)
    lin = pd(gains.linear, rot(R_et, p_e), ed_lin)
    return np.concatenate([ang, lin], axis=-1)


def _joint_transform(mech: Mechanism, i: int, q: np.ndarray) -> Tuple[np.ndarray, np.ndarray]:
    B = q.shape[0]
    o = int(mech.qoff[i])
    Rj = np.broadcast_to(np.eye(3), (B, 3, 3)).copy()
    pj = np.zeros((B, 3))
    t = int(mech.jtype[i])
    if t == REVOLUTE:
        a = mech.axis[i]
        K = np.array([[0, -a[2], a[1]], [a[2], 0, -a[0]], [-a[1], a[0], 0]])
        Rj = np.eye(3) + np.sin(q[:, o])[:, None, None] * K + (1 - np.cos(q[:, o]))[:, None, None] * (K @ K)
    elif t == PRISMATIC:
        pj = q[:, o:o + 1] * mech.axis[i]
    elif t == QUAT_FLOATING:
        Rj = quat_to_rot(q[:, o:o + 4] / np.linalg.norm(q[:, o:o + 4], axis=-1, keepdims=True))
        pj = q[:, o + 4:o + 7]
    return Rj, pj


def body_pose_and_twist(mech: Mechanism, q, v, body: int, base: int = -1):
    """relative_transform(state, body, base) and relative_twist(state, body, base) expressed in the body frame
    (se3pdcontroller.jl:15-16), for a batch of states q [B, nq], v [B, nv].  `base` = -1 is the world.

    Returns (R [B,3,3], p [B,3]) and the twist [B,6] = (omega; nu)."""
    q = np.atleast_2d(np.asarray(q, dtype=np.float64))
    v = np.atleast_2d(np.asarray(v, dtype=np.float64))

    def to_root(b):
        """pose in the world and the twist w.r.t. the world expressed in the body's own frame"""
        B = q.shape[0]
        if b < 0:
            return np.broadcast_to(np.eye(3), (B, 3, 3)), np.zeros((B, 3)), np.zeros((B, 6))
        chain = mech.ancestors(b)[::-1]  # root-most body first
        R = np.broadcast_to(np.eye(3), (B, 3, 3))
        p = np.zeros((B, 3))
        tw = np.zeros((B, 6))
        for i in chain:
            Rj, pj = _joint_transform(mech, i, q)
            Rl = mech.X_R[i] @ Rj  # body i -> parent
            pl = np.einsum("ij,bj->bi", mech.X_R[i], pj) + mech.X_p[i]
            # parent's twist re-expressed in body i: omega' = Rl' omega, nu' = Rl' (nu + omega x pl)
            Rlt = np.swapaxes(Rl, -1, -2)
            w_par, nu_par = tw[:, :3], tw[:, 3:]
            w = np.einsum("bij,bj->bi", Rlt, w_par)
            nu = np.einsum("bij,bj->bi", Rlt, nu_par + np.cross(w_par, pl))
            o, t = int(mech.voff[i]), int(mech.jtype[i])
            if t == REVOLUTE:
                w = w + v[:, o:o + 1] * mech.axis[i]
            elif t == PRISMATIC:
                nu = nu + v[:, o:o + 1] * mech.axis[i]
            elif t == QUAT_FLOATING:
                w = w + v[:, o:o + 3]
                nu = nu + v[:, o + 3:o + 6]
            tw = np.concatenate([w, nu], axis=-1)
            p = np.einsum("bij,bj->bi", R, pl) + p
            R = R @ Rl
        return R, p, tw

    Rb, pb, twb = to_root(body)
    Ra, pa, twa = to_root(base)
    Rat = np.swapaxes(Ra, -1, -2)
    R = Rat @ Rb
    p = np.einsum("bij,bj->bi", Rat, pb - pa)
    # the base's twist re-expressed in the body frame (body -> base: R, p)
    Rt = np.swapaxes(R, -1, -2)
    w_a = np.einsum("bij,bj->bi", Rt, twa[:, :3])
    nu_a = np.einsum("bij,bj->bi", Rt, twa[:, 3:] + np.cross(twa[:, :3], p))
    return (R, p), twb - np.concatenate([w_a, nu_a], axis=-1)


class SE3PDController:
    """se3pdcontroller.jl:1-18.  `trajectory` is an `SE3Trajectory` (or any callable with its return convention) of
    `body` relative to `base`; `weight` is carried for the caller's `addtask!` like in the reference (it is not used by
    the functor); `gains` may be swapped between ticks (the reference holds it in a Ref)."""

    def __init__(self, base: int, body: int, trajectory, weight, gains: SE3PDGains):
        self.base, self.body, self.trajectory, self.weight, self.gains = base, body, trajectory, weight, gains

    def __call__(self, t, mech: Mechanism, q, v) -> np.ndarray:
        """Desired spatial acceleration [B, 6] (angular; linear) of `body` w.r.t. `base` in the body frame at time(s) t
        for the batch of states (q, v)."""
        (R_des, p_des), (w_des, nu_des), (wd_des, nud_des) = self.trajectory(t, 2)
        (R, p), twist = body_pose_and_twist(mech, q, v, self.body, self.base)
        feed_forward = np.concatenate([np.broadcast_to(wd_des, twist[:, :3].shape),
                                       np.broadcast_to(nud_des, twist[:, 3:].shape)], axis=-1)
        twist_des = np.concatenate([np.broadcast_to(w_des, twist[:, :3].shape),
                                    np.broadcast_to(nu_des, twist[:, 3:].shape)], axis=-1)
        return feed_forward + pd_se3(self.gains, R, p, R_des, p_des, twist, twist_des)


# ---- device form: what qpc_add_se3pd / qpc_se3pd_update take (include/qpcontrol_b200.h: qpc_interp_piece) -------------------
def _gain3(g: Gain) -> np.ndarray:
    g = np.asarray(g, dtype=np.float64)
    if g.ndim == 0:
        return float(g) * np.eye(3)
    if g.ndim == 1:
        return np.diag(g)
    return g


def gains_matrix(gains: SE3PDGains) -> np.ndarray:
    """[4, 3, 3]: K_angular, D_angular, K_linear, D_linear (scalar and per-axis gains become diagonal matrices)."""
    return np.ascontiguousarray(np.stack([_gain3(gains.angular.k), _gain3(gains.angular.d), _gain3(gains.linear.k),
                                          _gain3(gains.linear.d)]))


def _piece(f, rotation: bool, break_start: float) -> dict:
    if isinstance(f, Constant):
        y0 = np.asarray(f.value, dtype=np.float64).ravel()
        return dict(break_start=break_start, x0=0.0, xf=1.0, y0=y0, dy=np.array([1.0, 0, 0]) if rotation else np.zeros(3),
                    angle=0.0, coeffs=[])
    if not isinstance(f, Interpolated):
        raise TypeError("the device evaluates Interpolated / Constant trajectories and Piecewise lists of them; got "
                        + type(f).__name__)
    if not f.clamp:
        raise ValueError("device trajectories are clamped to their range (Interpolated(..., clamp=True))")
    if bool(f.rotation) != rotation:
        raise ValueError("angular components interpolate rotations, linear components 3-vectors")
    if f.interpolator is None:
        coeffs = []
    elif isinstance(f.interpolator, Polynomial):
        coeffs = list(f.interpolator.coeffs)
        if len(coeffs) > 6:
            raise ValueError("interpolator polynomials have at most 6 coefficients (quintic)")
    else:
        raise TypeError("the interpolator must be None (identity) or a Polynomial")
    if rotation:
        return dict(break_start=break_start, x0=f.x0, xf=f.xf, y0=f.y0, dy=f.axis, angle=float(f.angle), coeffs=coeffs)
    return dict(break_start=break_start, x0=f.x0, xf=f.xf, y0=f.y0, dy=f.yf - f.y0, angle=0.0, coeffs=coeffs)


def compile_trajectory(traj, rotation: bool):
    """(pieces, piecewise, break_end) of one component of an SE3Trajectory for the device tables."""
    while isinstance(traj, PointTrajectory):  # Point / FreeVector wrappers only carry the frame
        traj = traj.trajectory
    if isinstance(traj, Piecewise):
        if not traj.clamp:
            raise ValueError("device trajectories are clamped to their range (Piecewise(..., clamp=True))")
        if len(traj.subfunctions) > MAX_PIECES:
            raise ValueError(f"at most {MAX_PIECES} pieces per trajectory")
        return ([_piece(f, rotation, float(b)) for f, b in zip(traj.subfunctions, traj.breaks[:-1])], True,
                float(traj.breaks[-1]))
    return [_piece(traj, rotation, 0.0)], False, 0.0
