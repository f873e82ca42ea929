"""Restates the reference's own controller tests (reference test/tasks.jl, test/controller.jl) against the oracle,
plus a solver-independent KKT check of the lifted QP in numpy.  No GPU needed."""
import numpy as np
import pytest

import util
from qpcontrol_jl_b200 import (AngularAccelerationTask, JointAccelerationTask, LinearAccelerationTask,
                               LinearMomentumRateTask, MomentumBasedController, MomentumRateTask, OSQPSettings,
                               PointAccelerationTask, SpatialAccelerationTask, scenarios)
from qpcontrol_jl_b200.mechanism import PRISMATIC, REVOLUTE, atlas_like, rand_floating_humanoid, rand_tree


# ---- test/tasks.jl ------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("cls,seed", [(SpatialAccelerationTask, 3), (AngularAccelerationTask, 2),
                                      (LinearAccelerationTask, 4)])
def test_path_task_error_identities(orc, cls, seed):
    rng = np.random.default_rng(seed)
    mech = rand_tree(rng, [REVOLUTE] * 10)
    for _ in range(10):
        base = int(rng.integers(-1, mech.nb))
        body = int(rng.choice([b for b in range(-1, mech.nb) if b != base]))
        ctrl = MomentumBasedController(mech, OSQPSettings.test_suite())
        task = cls(mech, base, body)
        ctrl.addtask(task)
        oc = orc.OracleController(ctrl.program)
        st = orc.OracleState(oc.om)
        q = mech.rand_configuration(rng)
        rows = {SpatialAccelerationTask: slice(0, 6), AngularAccelerationTask: slice(0, 3),
                LinearAccelerationTask: slice(3, 6)}[cls]
        # zero velocity, zero vd: error == -desired exactly  (tasks.jl:49,95,195)
        st.set(q, np.zeros(mech.nv))
        J, b = oc.task_rows(st, 0)
        assert np.all(b == 0.0)
        # random velocity: the bias term  (tasks.jl:57,110,203)
        v = rng.standard_normal(mech.nv)
        st.set(q, v)
        J, b = oc.task_rows(st, 0)
        assert np.array_equal(b, st.bias_in_frame(base, body, body)[rows])
        # J vd against the geometric Jacobian in the body frame  (tasks.jl:63,118,209)
        vd = rng.random(mech.nv)
        np.testing.assert_allclose(J @ vd, (st.geometric_jacobian(base, body, body) @ vd)[rows], atol=1e-12)


def test_point_task_is_point_acceleration(orc):
    """task_error(PointAccelerationTask) + desired == acceleration of the point expressed in the base frame
    (tasks.jl:153-171), checked by second differences of an independent forward kinematics."""
    rng = np.random.default_rng(6)
    mech = rand_tree(rng, [REVOLUTE] * 8)
    for _ in range(6):
        base = int(rng.integers(-1, mech.nb))
        body = int(rng.choice([b for b in range(mech.nb) if b != base]))
        point = rng.standard_normal(3)
        ctrl = MomentumBasedController(mech, OSQPSettings.test_suite())
        ctrl.addtask(PointAccelerationTask(mech, base, body, point))
        oc = orc.OracleController(ctrl.program)
        st = orc.OracleState(oc.om)
        q, v = util.random_state(mech, rng)
        vd = rng.standard_normal(mech.nv)
        st.set(q, v)
        J, b = oc.task_rows(st, 0)

        def pos(qq):
            fk = util.forward_kinematics(mech, qq)
            Rs, ps = util.body_pose(fk, base)
            Rt, pt = util.body_pose(fk, body)
            return Rs.T @ (Rt @ point + pt - ps)

        dt = 1e-4
        qp = util.integrate_configuration(mech, q, v + 0.5 * dt * vd, dt)
        qm = util.integrate_configuration(mech, q, v - 0.5 * dt * vd, -dt)
        acc = (pos(qp) - 2 * pos(q) + pos(qm)) / dt ** 2
        np.testing.assert_allclose(J @ vd + b, acc, atol=5e-5)


def test_joint_and_momentum_task_rows(orc):
    rng = np.random.default_rng(1)
    mech = rand_tree(rng, [REVOLUTE] * 10, floating=True)
    ctrl = MomentumBasedController(mech, OSQPSettings.test_suite(), floatingjoint=0)
    ctrl.addtask(JointAccelerationTask(mech, 4))
    ctrl.addtask(MomentumRateTask(mech))
    ctrl.addtask(LinearMomentumRateTask(mech))
    oc = orc.OracleController(ctrl.program)
    st = orc.OracleState(oc.om)
    q, v = util.random_state(mech, rng)
    st.set(q, v)
    J, b = oc.task_rows(st, 0)
    vd = rng.random(mech.nv)
    assert np.array_equal(J @ vd, vd[list(mech.velocity_range(4))]) and np.all(b == 0)
    A = st.momentum_matrix(centroidal=True)
    com = st.center_of_mass()
    hb = st.momentum_rate_bias()
    hb_c = np.concatenate([hb[:3] - np.cross(com, hb[3:]), hb[3:]])
    J, b = oc.task_rows(st, 1)
    np.testing.assert_allclose(J, A, atol=1e-12)
    np.testing.assert_allclose(b, hb_c, atol=1e-12)
    J3, b3 = oc.task_rows(st, 2)
    assert np.array_equal(J3, J[3:]) and np.array_equal(b3, b[3:])


# ---- test/controller.jl -------------------------------------------------------------------------------------------
def forward_dynamics(st, mech, tau, ext=None):
    M = st.mass_matrix()
    c = st.inverse_dynamics(np.zeros(mech.nv), ext)
    return np.linalg.solve(M, tau - c)


@pytest.mark.parametrize("constrained", [True, False])
def test_fixed_base_joint_space_control(orc, constrained):
    """test/controller.jl:68-97: forward dynamics of the returned torques reproduces the desired accelerations."""
    rng = np.random.default_rng(42)
    mech = rand_tree(rng, [PRISMATIC, REVOLUTE, REVOLUTE])
    ctrl = MomentumBasedController(mech, OSQPSettings.test_suite())
    tasks = []
    for j in range(mech.nb):
        t = JointAccelerationTask(mech, j)
        ctrl.addtask(t) if constrained else ctrl.addtask(t, 1.0)
        t.setdesired(rng.random(1))
        tasks.append(t)
    oc = orc.OracleController(ctrl.program)
    q, v = util.random_state(mech, rng)
    out = oc.solve_batch(q, v)
    assert out["status"][0] == 1
    st = orc.OracleState(oc.om).set(q, v)
    vd = forward_dynamics(st, mech, out["tau"][0])
    np.testing.assert_allclose(vd, np.concatenate([t.desired for t in tasks]), atol=1e-7)


def add_all_contacts(ctrl, mech, rng=None):
    pts = []
    for body in range(mech.nb):
        for pos in mech.contact_points.get(body, ()):
            if rng is None:
                normal, mu = (0.0, 0.0, 1.0), mech.contact_mu
            else:  # parametric_contact_surface = true (test/controller.jl:110-118)
                normal = rng.standard_normal(3)
                normal /= np.linalg.norm(normal)
                mu = float(rng.random())
            pts.append(ctrl.addcontact(body, pos, normal, mu))
    return pts


def test_zero_velocity_free_fall(orc):
    """test/controller.jl:128-165: contacts present but disabled, regularisation 1.0 => joints do not accelerate and
    the base falls with gravity."""
    rng = np.random.default_rng(5354)
    mech = rand_floating_humanoid(rng)
    ctrl = MomentumBasedController(mech, OSQPSettings.test_suite(), floatingjoint=0)
    add_all_contacts(ctrl, mech)
    for j in range(1, mech.nb):
        ctrl.regularize(j, 1.0)
    oc = orc.OracleController(ctrl.program)
    q = mech.rand_configuration(rng)
    out = oc.solve_batch(q, np.zeros(mech.nv))
    assert out["status"][0] == 1
    vd = out["vd"][0]
    np.testing.assert_allclose(vd[6:], 0, atol=1e-4)
    np.testing.assert_allclose(vd[:3], 0, atol=1e-4)
    R = util.quat_to_rot(q[:4])
    np.testing.assert_allclose(R @ vd[3:6], mech.gravity, atol=1e-4)
    np.testing.assert_allclose(out["wrenches"][0], 0, atol=1e-6)


def test_achievable_momentum_rate(orc):
    """test/controller.jl:169-230: random active contact sets, random in-cone forces; the hard MomentumRateTask is met."""
    rng = np.random.default_rng(533454)
    mech = rand_floating_humanoid(rng)
    ctrl = MomentumBasedController(mech, OSQPSettings.test_suite(), floatingjoint=0)
    pts = add_all_contacts(ctrl, mech, rng)
    task = MomentumRateTask(mech)
    ctrl.addtask(task)
    for j in range(mech.nb):
        ctrl.regularize(j, 1e-6)
    oc = orc.OracleController(ctrl.program)
    st = orc.OracleState(oc.om)
    for p in np.linspace(0, 1, 5):
        q, v = util.random_state(mech, rng)
        st.set(q, v)
        com = st.center_of_mass()
        fk = util.forward_kinematics(mech, q)
        hd = np.concatenate([np.zeros(3), mech.total_mass * mech.gravity])
        for c in pts:
            if rng.random() < p:
                c.weight, c.maxnormalforce = 1e-6, 1e9
                fn = 50.0 * rng.random()
                mur = np.sqrt(2) / 2 * c.mu
                d = rng.standard_normal(3)
                ft = mur * fn * rng.random() * np.cross(c.normal, d / np.linalg.norm(d))
                f = fn * c.normal + ft
                R, pb = fk[c.body]
                fw = R @ f
                pw = R @ c.position + pb
                hd += np.concatenate([np.cross(pw - com, fw), fw])
            else:
                c.disable()
        task.setdesired(hd)
        out = oc.solve_batch(q, v)
        assert out["status"][0] in (1, 2)
        vd = out["vd"][0]
        hdot = st.momentum_matrix(centroidal=True) @ vd
        hb = st.momentum_rate_bias()
        hdot += np.concatenate([hb[:3] - np.cross(com, hb[3:]), hb[3:]])
        np.testing.assert_allclose(hdot, hd, atol=1e-3)


@pytest.mark.parametrize("mode", ["constraint", "scalar", "matrix"])
def test_spatial_acceleration_modes(orc, mode):
    """test/controller.jl:232-285: left foot w.r.t. right palm, expressed in the palm frame; hard constraint, scalar
    weight and matrix weight all achieve the desired relative spatial acceleration.  (The two `Parameter` weight
    modes differ from these only in when the weight is read.)"""
    rng = np.random.default_rng(533)
    mech = rand_floating_humanoid(rng)
    ctrl = MomentumBasedController(mech, OSQPSettings.test_suite(), floatingjoint=0)
    body, base = mech.findbody("l_foot"), mech.findbody("r_hand")
    task = SpatialAccelerationTask(mech, base, body, frame=base)
    if mode == "constraint":
        ctrl.addtask(task)
        for j in range(mech.nb):
            ctrl.regularize(j, 1.0)
    elif mode == "scalar":
        ctrl.addtask(task, 1.0)
    else:
        ctrl.addtask(task, np.eye(6))
    des = rng.random(6)
    task.setdesired(des)
    oc = orc.OracleController(ctrl.program)
    q, v = util.random_state(mech, rng)
    out = oc.solve_batch(q, v)
    assert out["status"][0] == 1
    st = orc.OracleState(oc.om).set(q, v)
    acc = st.geometric_jacobian(base, body, base) @ out["vd"][0] + st.bias_in_frame(base, body, base)
    np.testing.assert_allclose(acc, des, atol=1e-6)


def test_parameterized_contact_frame(orc):
    """test/controller.jl:22-66: a body rotating about x with a contact whose normal is fixed in the world; the world
    force of a unit local normal force stays (0, 0, 1) whatever the joint angle."""
    from qpcontrol_jl_b200.mechanism import _Builder
    b = _Builder()
    b.add("body", "rx", None, REVOLUTE, axis=(1, 0, 0), mass=10.0, inertia=(1.0, 1.0, 1.0))
    mech = b.build()
    for th in np.linspace(-np.pi, np.pi, 10):
        R = util.expm_so3(np.array([th, 0, 0]))
        normal_body = R.T @ np.array([0.0, 0.0, 1.0])  # the Parameter of test/controller.jl:42-46
        ctrl = MomentumBasedController(mech, OSQPSettings.test_suite())
        c = ctrl.addcontact(0, (0.0, 0.0, 0.0), normal_body, 1.0)
        c.maxnormalforce, c.weight = 1e3, 1e-3
        oc = orc.OracleController(ctrl.program)
        P, qv, A, l, u = oc.lifted_qp(np.array([th]), np.zeros(1))
        # rows 11..13: lin(w) - R f == 0 ; column block of f_local is 5..7 (after vd and rho)
        Rtot = -A[11:14, 5:8]
        np.testing.assert_allclose(Rtot @ np.array([0, 0, 1.0]), [0, 0, 1.0], atol=1e-12)


# ---- Atlas standing (notebooks/Standing controller.ipynb) ---------------------------------------------------------------
def kkt_check(P, qv, A, l, u, x, y, tol):
    """Solver-independent optimality check of min 1/2 x'Px + q'x, l <= Ax <= u."""
    stat = P @ x + qv + A.T @ y
    z = A @ x
    scale = max(1.0, np.abs(x).max())
    assert np.abs(stat).max() <= tol * max(1.0, np.abs(A.T @ y).max())
    assert np.all(z >= l - tol * scale) and np.all(z <= u + tol * scale)
    # complementary slackness: y_i < 0 only on active lower bounds, y_i > 0 only on active upper bounds
    for i in range(len(y)):
        if y[i] > tol * max(1, np.abs(y).max()):
            assert abs(z[i] - u[i]) <= tol * scale
        if y[i] < -tol * max(1, np.abs(y).max()):
            assert abs(z[i] - l[i]) <= tol * scale


def test_atlas_standing_lifted_dims_and_kkt(orc):
    mech, low, ctrl, qnom = scenarios.atlas_standing(OSQPSettings.test_suite())
    oc = orc.OracleController(low.program)
    assert (oc.nvar, oc.nrows) == (143, 178)
    q, v = scenarios.atlas_random_states(mech, qnom, 4, seed=3)
    out = oc.solve_batch(q, v, return_lifted=True)
    assert np.all(out["status"] == 1)
    assert np.all(out["tau"][:, :6] == 0.0)
    import ctypes as C
    for i in range(4):
        P, qv, A, l, u = oc.lifted_qp(q[i], v[i])
        x = out["x_lifted"][i]
        # duals are not exported by the batch call; recover them by least squares on the stationarity condition of
        # the active rows and check primal feasibility + stationarity residual
        z = A @ x
        act = (np.abs(z - l) < 1e-6) | (np.abs(z - u) < 1e-6)
        y_act, *_ = np.linalg.lstsq(A[act].T, -(P @ x + qv), rcond=None)
        y = np.zeros(len(l))
        y[act] = y_act
        assert np.abs(P @ x + qv + A.T @ y).max() < 1e-5
        assert np.all(z >= l - 1e-6) and np.all(z <= u + 1e-6)
        # friction cones: every contact force inside its (inner-approximated) cone, pushing
        for c in range(8):
            f_local = x[36 + 13 * c + 4:36 + 13 * c + 7]
            assert f_local[2] >= -1e-6
            assert np.hypot(f_local[0], f_local[1]) <= mech.contact_mu * f_local[2] + 1e-6
    # Newton-Euler: A vd + Adot v == W_gravity + sum of contact wrenches (momentum.jl:190)
    st = orc.OracleState(oc.om)
    for i in range(4):
        st.set(q[i], v[i])
        fg = mech.total_mass * mech.gravity
        Wg = np.concatenate([np.cross(st.center_of_mass(), fg), fg])
        lhs = st.momentum_matrix() @ out["vd"][i] + st.momentum_rate_bias()
        np.testing.assert_allclose(lhs, Wg + out["wrenches"][i].sum(0), atol=1e-5)
        # and the torques are consistent with forward dynamics under those contact wrenches
        ext = np.zeros((mech.nb, 6))
        for c, cp in enumerate(low.program.contacts):
            ext[cp.body] += out["wrenches"][i, c]
        np.testing.assert_allclose(forward_dynamics(st, mech, out["tau"][i], ext), out["vd"][i], atol=1e-5)


def test_standing_desireds(orc):
    """standing.jl:60-85 restated in numpy."""
    mech, low, ctrl, qnom = scenarios.atlas_standing()
    oc = orc.OracleController(low.program)
    q, v = scenarios.atlas_random_states(mech, qnom, 1, seed=9)
    st = orc.OracleState(oc.om).set(q[0], v[0])
    des = oc.standing_desireds(st)
    sp = low.program.standing
    offs = low.program.des_offsets()
    m = mech.total_mass
    c = st.center_of_mass()
    h = st.momentum()
    np.testing.assert_allclose(des[offs[sp.linmom_task]:offs[sp.linmom_task] + 3],
                               m * (-sp.com_kp * (c - sp.comref) - sp.com_kd * h[3:] / m), atol=1e-10)
    R = util.quat_to_rot(q[0, :4])
    ang = np.arccos(np.clip((np.trace(R) - 1) / 2, -1, 1))
    axis = np.array([R[2, 1] - R[1, 2], R[0, 2] - R[2, 0], R[1, 0] - R[0, 1]]) / (2 * np.sin(ang))
    np.testing.assert_allclose(des[offs[sp.pelvis_task]:offs[sp.pelvis_task] + 3],
                               -sp.pelvis_kp * ang * axis - sp.pelvis_kd * v[0, :3], atol=1e-9)
    for k, j in enumerate(sp.joints):
        np.testing.assert_allclose(des[offs[sp.joint_tasks[k]]],
                                   -100.0 * (q[0, mech.qoff[j]] - qnom[mech.qoff[j]]) - 20.0 * v[0, mech.voff[j]])
    assert len(sp.joints) == 18


def test_dense_qp_solver_kkt(orc):
    P, qv, A, l, u = scenarios.synthetic_qps(3, 30, 30, seed=5)
    out = orc.solve_dense_qp_batch(P, qv, A, l, u, eps_abs=1e-9, eps_rel=1e-9)
    assert np.all(out["status"] == 1)
    for i in range(3):
        kkt_check(P[i], qv[i], A[i], l[i], u[i], out["x"][i], out["y"][i], 1e-6)
