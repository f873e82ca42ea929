"""GPU coverage matrix of SURVEY.md 8(a) through the C ABI (qpc_solve_batch) against the oracle:

  * every task kind of reference src/tasks.jl:3-262 (Spatial / Angular / Linear / Point / Joint acceleration, momentum
    rate, linear momentum rate) x every `addtask!` mode of src/lowlevel/momentum.jl:99-117 (hard constraint, scalar
    weight with slack variables, matrix weight);
  * the reference's integration invariants restated against the CUDA path: free fall (test/controller.jl:128-165),
    achievable momentum rate with random active contact sets p in {0, .25, .5, .75, 1} (:169-230), the five weight modes
    giving the same achieved spatial acceleration (:232-285);
  * dense QPs at the remaining sizes of BASELINE config 5 (n = 50, the lifted Atlas pair (143, 178), (200, 200))."""
import numpy as np
import pytest

import parity
import util
from qpcontrol_jl_b200 import (AngularAccelerationTask, JointAccelerationTask, LinearAccelerationTask,
                               LinearMomentumRateTask, MomentumBasedController, MomentumRateTask, OSQPSettings,
                               PointAccelerationTask, SpatialAccelerationTask, scenarios)
from qpcontrol_jl_b200.mechanism import rand_floating_humanoid

pytestmark = pytest.mark.gpu

KINDS = ["spatial", "angular", "linear", "point", "joint", "momentum_rate", "linear_momentum_rate"]
MODES = ["hard", "scalar", "matrix"]


def _make_task(kind, mech, rng):
    base, body = mech.findbody("pelvis"), mech.findbody("l_foot")
    if kind == "spatial":
        return SpatialAccelerationTask(mech, base, body, frame=base)
    if kind == "angular":
        return AngularAccelerationTask(mech, base, body, frame=base)
    if kind == "linear":
        return LinearAccelerationTask(mech, base, body, frame=base)
    if kind == "point":
        return PointAccelerationTask(mech, base, body, rng.uniform(-0.2, 0.2, 3))
    if kind == "joint":
        return JointAccelerationTask(mech, mech.nb - 1)
    if kind == "momentum_rate":
        return MomentumRateTask(mech)
    return LinearMomentumRateTask(mech)


def _controller(kind, mode, seed=77):
    rng = np.random.default_rng(seed)
    mech = rand_floating_humanoid(rng)
    ctrl = MomentumBasedController(mech, OSQPSettings.test_suite(), floatingjoint=0)
    for body in range(mech.nb):
        for pos in mech.contact_points.get(body, ()):
            c = ctrl.addcontact(body, pos, (0.0, 0.0, 1.0), mech.contact_mu)
            c.weight, c.maxnormalforce = 1e-3, 1e5
    task = _make_task(kind, mech, rng)
    d = task.dimension
    if mode == "hard":
        ctrl.addtask(task)
    elif mode == "scalar":
        ctrl.addtask(task, 3.0)
    else:
        W = rng.random((d, d))
        ctrl.addtask(task, W @ W.T + np.eye(d))
    for j in range(mech.nb):
        ctrl.regularize(j, 0.05)      # makes vd unique in every mode: the comparison is on torques, not on a task residual
    if kind in ("momentum_rate", "linear_momentum_rate"):
        # contacts can only push: ask for a momentum rate the feet can produce (they carry 70 % of the weight)
        des = (np.concatenate([np.zeros(3), 0.3 * mech.total_mass * mech.gravity]) +
               np.concatenate([rng.uniform(-0.2, 0.2, 3), rng.uniform(-1, 1, 3)]))[-d:]
    else:
        des = rng.uniform(-1, 1, d)
    task.setdesired(des)
    return mech, ctrl, task, rng


def _upright_states(mech, rng, B):
    """Floating base near the identity orientation, moderate joint angles: the feet's contact normals point roughly up, so
    that momentum-rate requests which need the feet to carry weight are feasible (contacts can only push)."""
    q = np.zeros((B, mech.nq))
    q[:, 0] = 1.0
    q[:, 1:4] = 0.05 * rng.standard_normal((B, 3))
    q[:, :4] /= np.linalg.norm(q[:, :4], axis=1, keepdims=True)
    q[:, 4:7] = rng.uniform(-0.5, 0.5, (B, 3))
    q[:, 7:] = 0.15 * rng.standard_normal((B, mech.nq - 7))
    return q, 0.2 * rng.standard_normal((B, mech.nv))


@pytest.mark.parametrize("mode", MODES)
@pytest.mark.parametrize("kind", KINDS)
def test_task_kind_and_weight_mode_match_oracle(orc, kind, mode):
    mech, ctrl, task, rng = _controller(kind, mode)
    B = 24
    q, v = _upright_states(mech, rng, B)
    res = ctrl(q, v, check=False)
    ref = orc.OracleController(ctrl.program).solve_batch(q, v)
    ok = (ref["status"] == 1) & (res.status == 1)
    # a hard full momentum-rate task fixes the centre of pressure: states whose CoM is not over the feet are infeasible --
    # for both solvers (status -3), which is what the accept / reject comparison below pins
    assert ok.mean() > (0.3 if (kind, mode) == ("momentum_rate", "hard") else 0.9), (kind, mode, res.status, ref["status"])
    assert np.array_equal((ref["status"] == 1) | (ref["status"] == 2), (res.status == 1) | (res.status == 2))
    assert parity.rel_err(res.tau[ok], ref["tau"][ok]).max() < parity.REL_TOL
    assert parity.rel_err(res.vdot[ok], ref["vd"][ok]).max() < parity.REL_TOL
    assert parity.rel_err(res.wrenches[ok], ref["wrenches"][ok]).max() < parity.REL_TOL
    assert np.all(res.tau[:, :6] == 0.0)


def _add_all_contacts(ctrl, mech, rng=None):
    pts = []
    for body in range(mech.nb):
        for pos in mech.contact_points.get(body, ()):
            if rng is None:
                normal, mu = (0.0, 0.0, 1.0), mech.contact_mu
            else:
                normal = rng.standard_normal(3)
                normal /= np.linalg.norm(normal)
                mu = float(rng.random())
            pts.append(ctrl.addcontact(body, pos, normal, mu))
    return pts


def test_free_fall_through_the_cuda_path():
    """test/controller.jl:128-165: contacts present but disabled (maxnormalforce = 0), regularisation 1.0, zero velocity
    => joints do not accelerate, the base falls with gravity, no contact wrench."""
    rng = np.random.default_rng(5354)
    mech = rand_floating_humanoid(rng)
    ctrl = MomentumBasedController(mech, OSQPSettings.test_suite(), floatingjoint=0)
    pts = _add_all_contacts(ctrl, mech)
    for c in pts:
        c.disable()
    for j in range(1, mech.nb):
        ctrl.regularize(j, 1.0)
    B = 32
    q = np.stack([mech.rand_configuration(rng) for _ in range(B)])
    res = ctrl(q, np.zeros((B, mech.nv)))
    assert np.all(res.status == 1)
    np.testing.assert_allclose(res.vdot[:, 6:], 0, atol=1e-4)
    np.testing.assert_allclose(res.vdot[:, :3], 0, atol=1e-4)
    for i in range(B):
        R = util.quat_to_rot(q[i, :4])
        np.testing.assert_allclose(R @ res.vdot[i, 3:6], mech.gravity, atol=1e-4)
    np.testing.assert_allclose(res.wrenches, 0, atol=1e-6)
    np.testing.assert_allclose(res.tau, 0, atol=1e-4)


def test_achievable_momentum_rate_through_the_cuda_path(orc):
    """test/controller.jl:169-230: parametric contact surfaces, random active contact sets (each point enabled with
    probability p in {0, .25, .5, .75, 1}), random in-cone forces define an achievable momentum rate; the hard
    MomentumRateTask is met: A vd + Adot v = hdot_desired (1e-3), contact wrenches inside their cones.  One batch per p,
    per-instance desireds and contact sets."""
    rng = np.random.default_rng(533454)
    mech = rand_floating_humanoid(rng)
    ctrl = MomentumBasedController(mech, OSQPSettings.test_suite(), floatingjoint=0)
    pts = _add_all_contacts(ctrl, mech, rng)
    task = MomentumRateTask(mech)
    ctrl.addtask(task)
    for j in range(mech.nb):
        ctrl.regularize(j, 1e-6)
    om = orc.OracleMechanism(mech)
    st = orc.OracleState(om)
    nc = len(pts)
    per_p = 8
    for p in np.linspace(0, 1, 5):
        q = np.zeros((per_p, mech.nq)); v = np.zeros((per_p, mech.nv))
        des = np.zeros((per_p, 6)); cm = np.zeros((per_p, nc)); cw = np.full((per_p, nc), 1e-6)
        hdot_bias = []
        for i in range(per_p):
            q[i], v[i] = util.random_state(mech, rng)
            st.set(q[i], v[i])
            com = st.center_of_mass()
            fk = util.forward_kinematics(mech, q[i])
            hd = np.concatenate([np.zeros(3), mech.total_mass * mech.gravity])
            for k, c in enumerate(pts):
                if rng.random() < p:
                    cm[i, k] = 1e9
                    fn = 50.0 * rng.random()
                    mur = np.sqrt(2) / 2 * c.mu
                    d = rng.standard_normal(3)
                    ft = mur * fn * rng.random() * np.cross(c.normal, d / np.linalg.norm(d))
                    f = fn * c.normal + ft
                    R, pb = fk[c.body]
                    fw = R @ f
                    pw = R @ c.position + pb
                    hd += np.concatenate([np.cross(pw - com, fw), fw])
            des[i] = hd
            hb = st.momentum_rate_bias()
            hdot_bias.append((st.momentum_matrix(centroidal=True).copy(),
                              np.concatenate([hb[:3] - np.cross(com, hb[3:]), hb[3:]])))
        res = ctrl(q, v, desired=des, contact_weight=cw, contact_maxnormalforce=cm, check=False)
        ref = orc.OracleController(ctrl.program).solve_batch(q, v, desired=des, cweight=cw, cmaxnf=cm)
        ok = (res.status == 1) | (res.status == 2)
        # the reference draws one state per p and OSQP solves it; a batch of random draws may contain one (momentum rate on the
        # boundary of what the cones can produce) on which neither OSQP form reaches 1e-8 within 20,000 iterations: such
        # instances (1 of the 40 drawn here; the oracle's OSQP returns MAX_ITER_REACHED on it as well) are excluded from the
        # accept / reject comparison; the physical check below (1e-3) runs on every accepted solve
        limit = (res.iters >= 20000) | (ref["iters"] >= 20000)
        assert ok.mean() >= 0.75 and limit.mean() <= 0.25, (p, res.status, ref["status"])
        assert np.array_equal(ok[~limit], ((ref["status"] == 1) | (ref["status"] == 2))[~limit]), (p, res.status, ref["status"])
        for i in np.where(ok)[0]:
            A, hb = hdot_bias[i]
            np.testing.assert_allclose(A @ res.vdot[i] + hb, des[i], atol=1e-3, err_msg=f"p={p} instance {i}")
            np.testing.assert_allclose(res.wrenches[i][cm[i] == 0], 0, atol=1e-6)


def test_five_weight_modes_give_the_same_spatial_acceleration(orc):
    """test/controller.jl:232-285: hard constraint, scalar weight, scalar Parameter weight, matrix weight, matrix
    Parameter weight (here: per-tick `task_weight` for the scalar one) all achieve the desired spatial acceleration (1e-8
    in the reference, which solves to 1e-8; 1e-6 here on the achieved J vd + Jdot v)."""
    rng = np.random.default_rng(533)
    mech = rand_floating_humanoid(rng)
    base, body = mech.findbody("r_hand"), mech.findbody("l_foot")
    B = 6
    q = np.stack([mech.rand_configuration(rng) for _ in range(B)])
    v = rng.standard_normal((B, mech.nv))
    des = rng.random(6)
    Wm = rng.random((6, 6))
    Wm = Wm @ Wm.T + np.eye(6)
    om = orc.OracleMechanism(mech)
    st = orc.OracleState(om)
    achieved = {}
    for mode in ["hard", "scalar", "scalar_parameter", "matrix"]:
        ctrl = MomentumBasedController(mech, OSQPSettings.test_suite(), floatingjoint=0)
        task = SpatialAccelerationTask(mech, base, body, frame=base)
        kw = {}
        if mode == "hard":
            ctrl.addtask(task)
            for j in range(mech.nb):
                ctrl.regularize(j, 1.0)
        elif mode == "scalar":
            ctrl.addtask(task, 1.0)
        elif mode == "scalar_parameter":
            ctrl.addtask(task, 0.5)                       # setup-time value, overridden per tick below
            kw = dict(task_weight=np.full((B, 1), 2.5))    # momentum.jl:107-110 with a Parameter weight
        else:
            ctrl.addtask(task, Wm)
        task.setdesired(des)
        res = ctrl(q, v, check=False, **kw) if kw else ctrl(q, v, check=False)
        assert np.all(res.status == 1), (mode, res.status)
        acc = []
        for i in range(B):
            st.set(q[i], v[i])
            J, b = st.geometric_jacobian(base, body, base), st.bias_in_frame(base, body, base)
            acc.append(J @ res.vdot[i] + b)
        achieved[mode] = np.array(acc)
        np.testing.assert_allclose(achieved[mode], np.tile(des, (B, 1)), atol=1e-6, err_msg=mode)


@pytest.mark.parametrize("n,m", [(50, 50), (143, 178), (200, 200)])
def test_dense_qp_remaining_config5_sizes(orc, n, m):
    from qpcontrol_jl_b200 import _lib
    B = 8 if n > 100 else 16
    P, qv, A, l, u = scenarios.synthetic_qps(B, n, m, seed=5)
    st = OSQPSettings(eps_abs=1e-8, eps_rel=1e-8, max_iter=20000)
    res = _lib.solve_qp_batch_host(P, qv, A, l, u, settings=st)
    ref = orc.solve_dense_qp_batch(P, qv, A, l, u, eps_abs=1e-8, eps_rel=1e-8)
    assert np.all(res["status"] == 1) and np.all(ref["status"] == 1)
    assert parity.rel_err(res["x"], ref["x"]).max() < 1e-5
