"""Warm start and closed loop.

CPU part: the generic ADMM body (csrc/admm.cuh, compiled by tests/emu) warm-started from the solution of a nearby QP
needs fewer iterations and lands on the cold-start solution.  GPU part: the register-resident kernel does the same
through the C ABI, and the device closed loop (qpc_step_batch) follows a host loop built from single ticks checked
against the oracle."""
import numpy as np
import pytest

import qpc_loader

qpc = qpc_loader.load()
from qpcontrol_jl_b200 import OSQPSettings, scenarios  # noqa: E402

import parity  # noqa: E402


def _integrate_host(mech, q, v, vd, dt):
    """Reference integrator of the closed-loop test: semi-implicit Euler, exponential map on the floating quaternion."""
    q, v = q.copy(), v.copy()
    v += dt * vd
    for b in range(mech.nb):
        jt, qo, vo = mech.jtype[b], mech.qoff[b], mech.voff[b]
        if jt in (0, 1):
            q[:, qo] += dt * v[:, vo]
        elif jt == 2:
            w, x, y, z = (q[:, qo + k] for k in range(4))
            R = np.stack([1 - 2 * (y * y + z * z), 2 * (x * y - w * z), 2 * (x * z + w * y),
                          2 * (x * y + w * z), 1 - 2 * (x * x + z * z), 2 * (y * z - w * x),
                          2 * (x * z - w * y), 2 * (y * z + w * x), 1 - 2 * (x * x + y * y)], 1).reshape(-1, 3, 3)
            om, vl = v[:, vo:vo + 3], np.einsum("bij,bj->bi", R, v[:, vo + 3:vo + 6])
            th = np.linalg.norm(om, axis=1) * dt
            hs = np.where(th > 1e-12, np.sin(0.5 * th) / np.where(th > 1e-12, th, 1.0) * dt, 0.5 * dt)
            e = np.concatenate([np.cos(0.5 * th)[:, None], hs[:, None] * om], 1)
            qn = np.stack([w * e[:, 0] - x * e[:, 1] - y * e[:, 2] - z * e[:, 3],
                           w * e[:, 1] + x * e[:, 0] + y * e[:, 3] - z * e[:, 2],
                           w * e[:, 2] - x * e[:, 3] + y * e[:, 0] + z * e[:, 1],
                           w * e[:, 3] + x * e[:, 2] - y * e[:, 1] + z * e[:, 0]], 1)
            q[:, qo:qo + 4] = qn / np.linalg.norm(qn, axis=1, keepdims=True)
            q[:, qo + 4:qo + 7] += dt * vl
    return q, v


def test_emu_warm_start_fewer_iterations_same_solution():
    from emu import emu
    P, qv, A, l, u = scenarios.synthetic_qps(24, 30, 30, seed=11)
    st = OSQPSettings(eps_abs=1e-8, eps_rel=1e-8, max_iter=20000)
    cold0 = emu.solve_qp_batch(P, qv, A, l, u, settings=st)
    assert np.all(cold0["status"] == 1) and np.all(cold0["rho"] > 0)
    qv2 = qv + 0.01 * np.random.default_rng(0).standard_normal(qv.shape)  # the next tick: a nearby QP
    cold = emu.solve_qp_batch(P, qv2, A, l, u, settings=st)
    warm = emu.solve_qp_batch(P, qv2, A, l, u, settings=st, warm=cold0)
    assert np.all(warm["status"] == 1)
    assert parity.rel_err(warm["x"], cold["x"]).max() < 1e-6
    assert warm["iters"].mean() < 0.7 * cold["iters"].mean()
    # restarting from the exact solution of the same QP terminates at the first check
    again = emu.solve_qp_batch(P, qv2, A, l, u, settings=st, warm=cold)
    assert np.all(again["iters"] <= 2 * st.check_termination)


def test_host_integrator_keeps_quaternion_unit_and_is_first_order():
    mech, low, ctrl, qnom = scenarios.atlas_standing(OSQPSettings.test_suite())
    q, v = scenarios.atlas_random_states(mech, qnom, 4, seed=1)
    vd = np.random.default_rng(2).standard_normal(v.shape)
    q1, v1 = _integrate_host(mech, q, v, vd, 1e-3)
    assert np.allclose(np.linalg.norm(q1[:, :4], axis=1), 1.0, atol=1e-14)
    assert np.allclose(v1, v + 1e-3 * vd)
    assert np.abs(q1 - q).max() < 5e-3


@pytest.mark.gpu
def test_gpu_warm_start_matches_cold_and_saves_iterations():
    mech, low, ctrl, qnom = scenarios.atlas_standing(OSQPSettings.test_suite())
    q, v = scenarios.atlas_random_states(mech, qnom, 256, seed=21)
    cold0 = ctrl(q, v)
    q2, v2 = _integrate_host(mech, q, v, cold0.vdot, 2e-3)  # the state one tick later
    cold = ctrl(q2, v2)
    low.set_warm_start(True)
    ctrl(q, v)  # first tick after enabling: cold, stores the iterates
    warm = ctrl(q2, v2)
    low.set_warm_start(False)
    assert np.all(warm.status == 1) and np.all(cold.status == 1)
    assert parity.rel_err(warm.tau, cold.tau).max() < 1e-5
    assert parity.rel_err(warm.wrenches.reshape(256, -1), cold.wrenches.reshape(256, -1)).max() < 1e-5
    assert warm.iters.mean() < cold.iters.mean()
    # restarting from the solution of the same QP terminates at the first residual check
    low.set_warm_start(True)
    same = ctrl(q2, v2)
    low.set_warm_start(False)
    assert np.all(same.iters <= 2 * OSQPSettings.test_suite().check_termination)
    assert parity.rel_err(same.tau, cold.tau).max() < 1e-5
    # after reset the next tick is cold again: identical iteration counts to the cold controller
    low.set_warm_start(True)
    low.reset_warm_start()
    again = ctrl(q2, v2)
    low.set_warm_start(False)
    assert np.array_equal(again.iters, cold.iters) and np.array_equal(again.tau, cold.tau)


@pytest.mark.gpu
def test_gpu_closed_loop_follows_host_loop_and_oracle():
    from oracle import oracle as orc
    st = OSQPSettings.test_suite()
    mech, low, ctrl, qnom = scenarios.atlas_standing(st)
    B, dt, nsteps = 16, 2e-3, 10
    q0, v0 = scenarios.atlas_random_states(mech, qnom, B, seed=31)
    oc = orc.OracleController(low.program)
    oc.set_settings(st, warm_start=0)
    qh, vh = q0.copy(), v0.copy()
    for _ in range(nsteps):  # host loop: oracle tick (cold, tight tolerance) + the reference integrator
        r = oc.solve_batch(qh, vh)
        assert np.all(r["status"] == 1)
        qh, vh = _integrate_host(mech, qh, vh, r["vd"], dt)
    for warm in (False, True):
        low.set_warm_start(warm)
        low.reset_warm_start()
        qd, vd_, res = ctrl.simulate(q0, v0, dt, nsteps)
        assert np.all(res.status == 1)
        assert np.abs(qd - qh).max() < 1e-6 and np.abs(vd_ - vh).max() < 1e-5
        assert np.allclose(np.linalg.norm(qd[:, :4], axis=1), 1.0, atol=1e-12)
    low.set_warm_start(False)


@pytest.mark.gpu
def test_gpu_standing_closed_loop_settles():
    """notebooks/Standing controller.ipynb:215-220 batched: perturbed robots come to rest at the reference posture."""
    st = OSQPSettings.standing_notebook()
    mech, low, ctrl, qnom = scenarios.atlas_standing(st)
    B, dt, nsteps = 64, 2e-3, 1500  # 3 s at 500 Hz
    sp = low.program.standing
    qj = np.array([mech.qoff[j] for j in sp.joints])
    vj = np.array([mech.voff[j] for j in sp.joints])
    # Start from the notebook's nominal stance with the position-controlled joints (arms, back, neck) displaced and
    # moving.  (Larger displacements, or displaced legs, move the CoM towards the edge of the support polygon and some
    # robots then tip over -- with this library and with the CPU oracle alike; these 64 settle in both.)
    rng = np.random.default_rng(41)
    q0 = np.tile(qnom, (B, 1))
    q0[:, qj] += 0.03 * rng.standard_normal((B, len(qj)))
    v0 = np.zeros((B, mech.nv))
    v0[:, vj] = 0.1 * rng.standard_normal((B, len(vj)))
    low.set_warm_start(True)
    low.reset_warm_start()
    q1, v1, res = ctrl.simulate(q0, v0, dt, nsteps, check=False)
    low.set_warm_start(False)
    assert np.all((res.status == 1) | (res.status == 2))
    assert np.abs(v1).max() < 5e-3
    err0 = np.abs(q0[:, qj] - np.asarray(sp.joint_ref)).max()
    err1 = np.abs(q1[:, qj] - np.asarray(sp.joint_ref)).max()
    assert err0 > 0.05 and err1 < 1e-4  # position-controlled joints are back at their references
    # the CoM reference sits 5 cm below the nominal CoM (standing.jl:28): the pelvis comes down by about that much
    assert np.all(np.abs(q1[:, 6] - (qnom[6] - 0.05)) < 0.02)
    assert np.allclose(np.linalg.norm(q1[:, :4], axis=1), 1.0, atol=1e-12)
    assert res.iters.mean() <= 100  # warm-started ticks near the fixed point stop at an early residual check
