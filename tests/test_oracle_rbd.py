"""Pins the oracle's restatement of RigidBodyDynamics (oracle/rbd.hpp) to physics, independently of the oracle:
finite differences of an independent numpy forward kinematics, kinetic-energy / momentum / power-balance identities.
These restate the assertions of reference test/tasks.jl (Jacobian and bias-term identities, :49-63 and parallels)
for which the reference itself uses RigidBodyDynamics as ground truth."""
import numpy as np
import pytest

import util
from qpcontrol_jl_b200.mechanism import PRISMATIC, QUAT_FLOATING, REVOLUTE, acrobot, atlas_like, rand_tree


def mechs():
    rng = np.random.default_rng(11)
    return [
        ("acrobot", acrobot()),
        ("tree10", rand_tree(rng, [REVOLUTE] * 10)),
        ("prr", rand_tree(rng, [PRISMATIC, REVOLUTE, REVOLUTE])),
        ("floating_tree", rand_tree(rng, [REVOLUTE, PRISMATIC, REVOLUTE, REVOLUTE, REVOLUTE], floating=True)),
        ("atlas", atlas_like()),
    ]


MECHS = mechs()


@pytest.fixture(params=MECHS, ids=[m[0] for m in MECHS])
def setup(request, orc):
    mech = request.param[1]
    om = orc.OracleMechanism(mech)
    return mech, orc.OracleState(om), np.random.default_rng(5)


def test_transforms_and_com_match_independent_fk(setup):
    mech, st, rng = setup
    q, v = util.random_state(mech, rng)
    st.set(q, v)
    fk = util.forward_kinematics(mech, q)
    com = np.zeros(3)
    for b in range(mech.nb):
        R, p = st.transform_to_root(b)
        np.testing.assert_allclose(R, fk[b][0], atol=1e-13)
        np.testing.assert_allclose(p, fk[b][1], atol=1e-13)
        com += mech.mass[b] * (fk[b][0] @ mech.com[b] + fk[b][1])
    np.testing.assert_allclose(st.center_of_mass(), com / mech.total_mass, atol=1e-13)


def test_geometric_jacobian_times_v_is_relative_twist(setup):
    """J v == twist of target w.r.t. source in `frame` (finite differences of poses), for random base/body/frame
    triples -- the `J * vd` identity of test/tasks.jl:63,118,164."""
    mech, st, rng = setup
    for _ in range(6):
        q, v = util.random_state(mech, rng)
        st.set(q, v)
        source, target, frame = (int(rng.integers(-1, mech.nb)) for _ in range(3))
        J = st.geometric_jacobian(source, target, frame)
        fd = util.relative_twist_fd(mech, q, v, source, target, frame)
        np.testing.assert_allclose(J @ v, fd, atol=2e-7)
        # columns of joints off the path are structurally zero
        onpath = {k for b, _ in mech.path(source, target) for k in mech.velocity_range(b)}
        for k in range(mech.nv):
            if k not in onpath:
                assert np.all(J[:, k] == 0.0)


def test_twist_wrt_world(setup):
    mech, st, rng = setup
    q, v = util.random_state(mech, rng)
    st.set(q, v)
    for b in range(mech.nb):
        np.testing.assert_allclose(st.twist_wrt_world(b), util.relative_twist_fd(mech, q, v, -1, b, -1), atol=2e-7)


def test_bias_term_is_time_derivative_of_twist_at_zero_vdot(setup):
    """`Jdot v`: with vd = 0 the time derivative of the frame-expressed relative twist equals
    transform(state, -bias(source) + bias(target), frame)  (test/tasks.jl:57,110,158)."""
    mech, st, rng = setup
    dt = 1e-5
    for _ in range(6):
        q, v = util.random_state(mech, rng)
        source, target, frame = (int(rng.integers(-1, mech.nb)) for _ in range(3))
        tw = []
        for s in (+1, -1):
            st.set(util.integrate_configuration(mech, q, v, s * dt), v)
            tw.append(st.geometric_jacobian(source, target, frame) @ v)
        st.set(q, v)
        np.testing.assert_allclose(st.bias_in_frame(source, target, frame), (tw[0] - tw[1]) / (2 * dt), atol=5e-8)


def test_momentum_matrix_and_rate_bias(setup):
    mech, st, rng = setup
    q, v = util.random_state(mech, rng)
    st.set(q, v)
    A = st.momentum_matrix()
    np.testing.assert_allclose(A @ v, st.momentum(), atol=1e-11)
    # momentum from first principles: sum over bodies of (angular about world origin; linear)
    fk = util.forward_kinematics(mech, q)
    h = np.zeros(6)
    for b in range(mech.nb):
        T = st.twist_wrt_world(b)
        R, p = fk[b]
        c = R @ mech.com[b] + p
        vc = T[3:] + np.cross(T[:3], c)
        Iw = R @ mech.inertia_com[b] @ R.T
        h[3:] += mech.mass[b] * vc
        h[:3] += Iw @ T[:3] + np.cross(c, mech.mass[b] * vc)
    np.testing.assert_allclose(st.momentum(), h, atol=1e-11)
    # centroidal version is the force-transform to the centre of mass
    Ac = st.momentum_matrix(centroidal=True)
    com = st.center_of_mass()
    np.testing.assert_allclose(Ac[3:], A[3:], atol=1e-12)
    np.testing.assert_allclose(Ac[:3], A[:3] - util.hat(com) @ A[3:], atol=1e-11)
    # Adot v = d/dt (A v) at vd = 0
    dt = 1e-5
    hs = []
    for s in (+1, -1):
        st.set(util.integrate_configuration(mech, q, v, s * dt), v)
        hs.append(st.momentum())
    st.set(q, v)
    np.testing.assert_allclose(st.momentum_rate_bias(), (hs[0] - hs[1]) / (2 * dt), atol=2e-6 * max(1, np.abs(h).max()))


def test_mass_matrix_energy_and_inverse_dynamics(setup):
    mech, st, rng = setup
    q, v = util.random_state(mech, rng)
    st.set(q, v)
    M = st.mass_matrix()
    np.testing.assert_allclose(M, M.T, atol=1e-12)
    assert np.linalg.eigvalsh(M).min() > 0
    ke = 0.0
    fk = util.forward_kinematics(mech, q)
    for b in range(mech.nb):
        T = st.twist_wrt_world(b)
        R, p = fk[b]
        c = R @ mech.com[b] + p
        vc = T[3:] + np.cross(T[:3], c)
        ke += 0.5 * mech.mass[b] * vc @ vc + 0.5 * T[:3] @ (R @ mech.inertia_com[b] @ R.T) @ T[:3]
    np.testing.assert_allclose(0.5 * v @ M @ v, ke, rtol=1e-12)
    # tau is affine in vd with slope M
    vd = rng.standard_normal(mech.nv)
    tau0 = st.inverse_dynamics(np.zeros(mech.nv))
    np.testing.assert_allclose(st.inverse_dynamics(vd) - tau0, M @ vd, atol=1e-10 * max(1, np.abs(M).max()))
    # external wrenches enter through the geometric Jacobian transpose
    ext = np.zeros((mech.nb, 6))
    b = int(rng.integers(mech.nb))
    ext[b] = rng.standard_normal(6)
    Jb = st.geometric_jacobian(-1, b, -1)
    np.testing.assert_allclose(st.inverse_dynamics(vd, ext), st.inverse_dynamics(vd) - Jb.T @ ext[b], atol=1e-10)


def test_power_balance(setup):
    """d/dt (kinetic + potential energy) == tau . v along the motion generated by (v, vd): pins the velocity-product
    and gravity terms of inverse dynamics."""
    mech, st, rng = setup
    q, v = util.random_state(mech, rng)
    vd = rng.standard_normal(mech.nv)
    st.set(q, v)
    tau = st.inverse_dynamics(vd)

    def energy(qq, vv):
        st.set(qq, vv)
        Mq = st.mass_matrix()
        return 0.5 * vv @ Mq @ vv - mech.total_mass * mech.gravity @ st.center_of_mass()

    dt = 1e-6
    # second-order accurate configuration update with the mid-step velocity
    e_p = energy(util.integrate_configuration(mech, q, v + 0.5 * dt * vd, dt), v + dt * vd)
    e_m = energy(util.integrate_configuration(mech, q, v - 0.5 * dt * vd, -dt), v - dt * vd)
    np.testing.assert_allclose((e_p - e_m) / (2 * dt), tau @ v, rtol=2e-5, atol=2e-5 * mech.total_mass)


def test_floating_rows_are_newton_euler(orc):
    """For a floating base the first six equations of motion are the momentum balance: A vd + Adot v = W_gravity
    when the floating-joint torques vanish (what add_wrench_balance_constraint! encodes, momentum.jl:162-193)."""
    mech = atlas_like()
    st = orc.OracleState(orc.OracleMechanism(mech))
    rng = np.random.default_rng(3)
    q, v = util.random_state(mech, rng)
    st.set(q, v)
    M = st.mass_matrix()
    c = st.inverse_dynamics(np.zeros(mech.nv))
    tau = np.concatenate([np.zeros(6), rng.standard_normal(mech.nv - 6) * 10])
    vd = np.linalg.solve(M, tau - c)
    fg = mech.total_mass * mech.gravity
    Wg = np.concatenate([np.cross(st.center_of_mass(), fg), fg])
    np.testing.assert_allclose(st.momentum_matrix() @ vd + st.momentum_rate_bias(), Wg, atol=1e-8)
