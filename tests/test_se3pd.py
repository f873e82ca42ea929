"""SE3PDController (reference src/lowlevel/se3pdcontroller.jl) restated on the host: qpcontrol.jl_b200/se3pd.py.

The reference has no test file for it (its PD law is RigidBodyDynamics.PDControl's, tested there); these tests pin the
restatement to the law's defining properties and the host kinematics to the oracle's."""
import numpy as np
import pytest

import util
from qpcontrol_jl_b200.mechanism import PRISMATIC, REVOLUTE, atlas_like, rand_tree
from qpcontrol_jl_b200.se3pd import (PDGains, SE3PDController, SE3PDGains, body_pose_and_twist, pd, pd_se3,
                                     rotation_vector)
from qpcontrol_jl_b200.trajectories import Constant, Interpolated, SE3Trajectory, fit_quintic, quat_to_rot


def _rot(axis, angle):
    axis = np.asarray(axis, dtype=np.float64) / np.linalg.norm(axis)
    K = np.array([[0, -axis[2], axis[1]], [axis[2], 0, -axis[0]], [-axis[1], axis[0], 0]])
    return np.eye(3) + np.sin(angle) * K + (1 - np.cos(angle)) * (K @ K)


def test_rotation_vector_inverts_the_exponential_map():
    rng = np.random.default_rng(0)
    axes = rng.normal(size=(64, 3))
    axes /= np.linalg.norm(axes, axis=-1, keepdims=True)
    angles = np.concatenate([[0.0, 1e-12, 1e-7, np.pi - 1e-9, np.pi - 1e-4], rng.uniform(0, np.pi, 59)])
    R = np.stack([_rot(a, t) for a, t in zip(axes, angles)])
    got = rotation_vector(R)
    np.testing.assert_allclose(got, axes * angles[:, None], atol=2e-8)
    assert rotation_vector(np.eye(3)).shape == (3,) and np.all(rotation_vector(np.eye(3)) == 0)


@pytest.mark.parametrize("name", ["floating_tree", "fixed_tree", "atlas"])
def test_host_pose_and_twist_match_the_oracle(orc, name):
    rng = np.random.default_rng(3)
    mech = {"floating_tree": lambda: rand_tree(rng, [REVOLUTE, PRISMATIC, REVOLUTE, REVOLUTE], floating=True),
            "fixed_tree": lambda: rand_tree(rng, [REVOLUTE] * 6), "atlas": atlas_like}[name]()
    st = orc.OracleState(orc.OracleMechanism(mech))
    qs, vs = zip(*[util.random_state(mech, rng) for _ in range(3)])
    q, v = np.stack(qs), np.stack(vs)
    body, base = mech.nb - 1, mech.nb // 2
    (R, p), tw = body_pose_and_twist(mech, q, v, body, -1)
    (Rr, pr), twr = body_pose_and_twist(mech, q, v, body, base)
    for i in range(3):
        st.set(q[i], v[i])
        Ro, po = st.transform_to_root(body)
        np.testing.assert_allclose(R[i], Ro, atol=1e-12)
        np.testing.assert_allclose(p[i], po, atol=1e-12)
        two = st.twist_wrt_world(body)  # (omega; v) expressed in the world frame
        w_world = R[i] @ tw[i, :3]
        np.testing.assert_allclose(w_world, two[:3], atol=1e-11)
        np.testing.assert_allclose(R[i] @ tw[i, 3:] + np.cross(p[i], w_world), two[3:], atol=1e-11)
        # relative to another body: relative_transform / relative_twist (se3pdcontroller.jl:15-16)
        Rb, pb = st.transform_to_root(base)
        np.testing.assert_allclose(Rr[i], Rb.T @ Ro, atol=1e-12)
        np.testing.assert_allclose(pr[i], Rb.T @ (po - pb), atol=1e-12)
        rel = two - st.twist_wrt_world(base)  # world frame
        w_b = Ro.T @ rel[:3]
        np.testing.assert_allclose(twr[i, :3], w_b, atol=1e-11)
        np.testing.assert_allclose(twr[i, 3:], Ro.T @ (rel[3:] + np.cross(rel[:3], po)), atol=1e-11)


def test_pd_law_properties():
    gains = SE3PDGains(PDGains(100.0, 20.0), PDGains(np.array([50.0, 60.0, 70.0]), np.diag([5.0, 6.0, 7.0])))
    R = _rot([1, 2, 3], 0.7)
    p = np.array([0.3, -0.2, 0.9])
    # no error: nothing to correct
    tw = np.array([0.1, 0.2, 0.3, -0.4, 0.5, 0.6])
    np.testing.assert_allclose(pd_se3(gains, R, p, R, p, tw, tw), 0, atol=1e-13)
    # position error only, at rest: -K R_e' p_e with R_e = I, p_e = R_des' (p - p_des)
    dp = np.array([0.01, 0.02, -0.03])
    out = pd_se3(gains, R, p + dp, R, p, np.zeros(6), np.zeros(6))
    np.testing.assert_allclose(out[:3], 0, atol=1e-13)
    np.testing.assert_allclose(out[3:], -np.array([50.0, 60.0, 70.0]) * (R.T @ dp), atol=1e-13)
    # orientation error only: the body is rotated by `angle` about a body axis away from the reference
    ax = np.array([0.0, 0.6, 0.8])
    out = pd_se3(gains, R @ _rot(ax, 0.2), p, R, p, np.zeros(6), np.zeros(6))
    np.testing.assert_allclose(out[:3], -100.0 * 0.2 * ax, atol=1e-12)
    np.testing.assert_allclose(out[3:], 0, atol=1e-13)
    # velocity error only: plain damping in the body frame
    out = pd_se3(gains, R, p, R, p, tw, np.zeros(6))
    np.testing.assert_allclose(out[:3], -20.0 * tw[:3], atol=1e-13)
    np.testing.assert_allclose(out[3:], -np.array([5.0, 6.0, 7.0]) * tw[3:], atol=1e-13)
    # batched call == per-instance calls
    rng = np.random.default_rng(1)
    Rs = np.stack([_rot(rng.normal(size=3), rng.uniform(0, 2)) for _ in range(5)])
    ps, tws = rng.normal(size=(5, 3)), rng.normal(size=(5, 6))
    batched = pd_se3(gains, Rs, ps, R, p, tws, tw)
    for i in range(5):
        np.testing.assert_array_equal(batched[i], pd_se3(gains, Rs[i], ps[i], R, p, tws[i], tw))
    np.testing.assert_array_equal(pd(PDGains(2.0, 3.0), [1.0, 0.0], [0.0, 1.0]), [-2.0, -3.0])


def test_controller_returns_feed_forward_on_the_reference_and_tracks_it_in_closed_loop():
    """A free rigid body (a one-body floating mechanism's kinematics) driven by the controller's spatial acceleration
    follows a quintic SE(3) reference and settles on its end pose."""
    rng = np.random.default_rng(2)
    mech = rand_tree(rng, [], floating=True)
    assert mech.nb == 1
    interp = fit_quintic(x0=0.0, xf=1.0, y0=0.0, yd0=0.0, ydd0=0.0, yf=1.0, ydf=0.0, yddf=0.0)
    q_end = np.array([np.cos(0.6), *(np.sin(0.6) * np.array([0.0, 0.6, 0.8]))])
    traj = SE3Trajectory(body=0, base=-1,
                         angular=Interpolated(0.0, 1.0, [1.0, 0, 0, 0], q_end, interp, rotation=True),
                         linear=Interpolated(0.0, 1.0, [0.0, 0.0, 1.0], [0.4, -0.3, 1.2], interp))
    ctrl = SE3PDController(-1, 0, traj, 1.0, SE3PDGains(PDGains(100.0, 20.0), PDGains(100.0, 20.0)))

    # on the reference itself the PD term vanishes (se3pdcontroller.jl:17 reduces to Tdref)
    (R_ref, p_ref), (w_ref, nu_ref), _ = traj(0.4, 2)
    quat = Interpolated(0.0, 1.0, [1.0, 0, 0, 0], q_end, interp, rotation=True)(0.4)
    q = np.concatenate([quat, p_ref])[None]
    v = np.concatenate([w_ref, nu_ref])[None]
    np.testing.assert_allclose(ctrl(0.4, mech, q, v)[0], traj.desired_spatial_acceleration(0.4), atol=1e-10)

    # closed loop from a perturbed start, two instances at once
    B, dt = 2, 1e-3
    quat0 = np.array([[np.cos(0.1), np.sin(0.1), 0, 0], [np.cos(0.15), 0, 0, -np.sin(0.15)]])
    q = np.concatenate([quat0, np.array([[0.05, 0.0, 0.9], [-0.1, 0.1, 1.1]])], axis=1)
    v = np.zeros((B, 6))
    hold = SE3Trajectory(0, -1, Constant(q_end, rotation=True), Constant(np.array([0.4, -0.3, 1.2])))
    for k in range(4000):
        t = k * dt
        ctrl.trajectory = traj if t <= 1.0 else hold
        a = ctrl(min(t, 1.0), mech, q, v)
        v = v + dt * a  # spatial acceleration in the body frame = rate of the body-frame twist components
        R = quat_to_rot(q[:, :4])
        dq = np.concatenate([np.ones((B, 1)), 0.5 * dt * v[:, :3]], axis=1)
        quat = np.stack([_qmul(q[i, :4], dq[i]) for i in range(B)])
        q = np.concatenate([quat / np.linalg.norm(quat, axis=1, keepdims=True),
                            q[:, 4:] + dt * np.einsum("bij,bj->bi", R, v[:, 3:])], axis=1)
    (R, p), tw = body_pose_and_twist(mech, q, v, 0, -1)
    np.testing.assert_allclose(p, np.broadcast_to([0.4, -0.3, 1.2], (B, 3)), atol=1e-4)
    np.testing.assert_allclose(R, np.broadcast_to(quat_to_rot(q_end), (B, 3, 3)), atol=1e-4)
    np.testing.assert_allclose(tw, 0, atol=1e-4)


def _qmul(a, b):
    w1, x1, y1, z1 = a
    w2, x2, y2, z2 = b
    return np.array([w1 * w2 - x1 * x2 - y1 * y2 - z1 * z2, w1 * x2 + x1 * w2 + y1 * z2 - z1 * y2,
                     w1 * y2 - x1 * z2 + y1 * w2 + z1 * x2, w1 * z2 + x1 * y2 - y1 * x2 + z1 * w2])
