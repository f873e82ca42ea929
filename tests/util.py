"""Test helpers independent of both the oracle and the CUDA path (plain numpy)."""
import numpy as np

import qpc_loader

qpc = qpc_loader.load()
M = qpc.mechanism if hasattr(qpc, "mechanism") else None
from qpcontrol_jl_b200.mechanism import PRISMATIC, QUAT_FLOATING, REVOLUTE  # noqa: E402


def hat(a):
    return np.array([[0, -a[2], a[1]], [a[2], 0, -a[0]], [-a[1], a[0], 0]])


def expm_so3(w):
    th = np.linalg.norm(w)
    if th < 1e-14:
        return np.eye(3) + hat(w)
    K = hat(w / th)
    return np.eye(3) + np.sin(th) * K + (1 - np.cos(th)) * (K @ K)


def quat_to_rot(q):
    w, x, y, z = q / np.linalg.norm(q)
    return np.array([[1 - 2 * (y * y + z * z), 2 * (x * y - w * z), 2 * (x * z + w * y)],
                     [2 * (x * y + w * z), 1 - 2 * (x * x + z * z), 2 * (y * z - w * x)],
                     [2 * (x * z - w * y), 2 * (y * z + w * x), 1 - 2 * (x * x + y * y)]])


def quat_mul(a, b):
    w1, x1, y1, z1 = a
    w2, x2, y2, z2 = b
    return np.array([w1 * w2 - x1 * x2 - y1 * y2 - z1 * z2, w1 * x2 + x1 * w2 + y1 * z2 - z1 * y2,
                     w1 * y2 - x1 * z2 + y1 * w2 + z1 * x2, w1 * z2 + x1 * y2 - y1 * x2 + z1 * w2])


def integrate_configuration(mech, q, v, dt):
    """q(t + dt) for constant generalized velocity v (floating joint: body-frame twist, exact exponential for the
    rotation, midpoint rule is unnecessary because tests use central differences)."""
    q2 = q.copy()
    for i in range(mech.nb):
        qo, vo = mech.qoff[i], mech.voff[i]
        if mech.jtype[i] in (REVOLUTE, PRISMATIC):
            q2[qo] = q[qo] + dt * v[vo]
        elif mech.jtype[i] == QUAT_FLOATING:
            w = v[vo:vo + 3] * dt
            th = np.linalg.norm(w)
            dq = np.array([1.0, 0, 0, 0]) if th < 1e-300 else np.concatenate([[np.cos(th / 2)], np.sin(th / 2) * w / th])
            q2[qo:qo + 4] = quat_mul(q[qo:qo + 4], dq)
            # position: p' = R v_lin (body-frame linear velocity); integrate with the mid-step rotation
            Rm = quat_to_rot(quat_mul(q[qo:qo + 4], np.concatenate([[np.cos(th / 4)], np.sin(th / 4) * w / max(th, 1e-300)])))
            q2[qo + 4:qo + 7] = q[qo + 4:qo + 7] + dt * (Rm @ v[vo + 3:vo + 6])
    return q2


def forward_kinematics(mech, q):
    """Independent numpy forward kinematics: list of (R, p) body -> world."""
    out = []
    for i in range(mech.nb):
        o = mech.qoff[i]
        Rj, pj = np.eye(3), np.zeros(3)
        if mech.jtype[i] == REVOLUTE:
            Rj = expm_so3(mech.axis[i] * q[o])
        elif mech.jtype[i] == PRISMATIC:
            pj = mech.axis[i] * q[o]
        elif mech.jtype[i] == QUAT_FLOATING:
            Rj = quat_to_rot(q[o:o + 4])
            pj = q[o + 4:o + 7]
        Rl, pl = mech.X_R[i] @ Rj, mech.X_R[i] @ pj + mech.X_p[i]
        if mech.parent[i] >= 0:
            Rp, pp = out[mech.parent[i]]
            Rl, pl = Rp @ Rl, Rp @ pl + pp
        out.append((Rl, pl))
    return out


def body_pose(fk, body):
    return (np.eye(3), np.zeros(3)) if body < 0 else fk[body]


def relative_twist_fd(mech, q, v, source, target, frame, dt=1e-6):
    """Twist of `target` w.r.t. `source` expressed in `frame`, by central differences of relative poses."""
    def rel(qq):
        fk = forward_kinematics(mech, qq)
        Rs, ps = body_pose(fk, source)
        Rt, pt = body_pose(fk, target)
        return Rs.T @ Rt, Rs.T @ (pt - ps), fk
    Rp, pp, _ = rel(integrate_configuration(mech, q, v, dt))
    Rm, pm, _ = rel(integrate_configuration(mech, q, v, -dt))
    R0, p0, fk0 = rel(q)
    Rdot = (Rp - Rm) / (2 * dt)
    pdot = (pp - pm) / (2 * dt)
    W = Rdot @ R0.T  # angular velocity of target wrt source, in source frame
    w_s = np.array([W[2, 1], W[0, 2], W[1, 0]])
    v_s = pdot - np.cross(w_s, p0)  # spatial (origin-referenced) linear part in the source frame
    # source frame -> `frame`
    Rs, ps = body_pose(fk0, source)
    Rf, pf = body_pose(fk0, frame)
    R = Rf.T @ Rs
    p = Rf.T @ (ps - pf)
    w = R @ w_s
    return np.concatenate([w, R @ v_s + np.cross(p, w)])


def random_state(mech, rng, vscale=1.0):
    return mech.rand_configuration(rng), vscale * rng.standard_normal(mech.nv)
