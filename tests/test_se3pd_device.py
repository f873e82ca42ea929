"""SE3PDController + SE3Trajectory evaluated ON THE DEVICE (reference
src/lowlevel/se3pdcontroller.jl:13-18, src/trajectories/{interpolated,piecewise,constant,se3}.jl): the assembly prologue
kin_se3pd (csrc/kin.cuh) writes `controller(t, state)` over the desired of the SpatialAccelerationTask it is bound to.
Checked against the host restatement (qpcontrol.jl_b200/se3pd.py + trajectories.py, itself tested in test_se3pd.py /
test_trajectories.py): the CPU tests run the kernel body through the g++ emulation, the GPU tests through the C ABI."""
import numpy as np
import pytest

import qpc_loader

qpc = qpc_loader.load()
from qpcontrol_jl_b200 import (MomentumBasedController, OSQPSettings, PDGains, SE3PDController, SE3PDGains,  # noqa: E402
                               SpatialAccelerationTask, scenarios)
from qpcontrol_jl_b200 import trajectories as T  # noqa: E402


def quat(axis, angle):
    axis = np.asarray(axis, dtype=np.float64) / np.linalg.norm(axis)
    return np.concatenate([[np.cos(angle / 2)], np.sin(angle / 2) * axis])


def make(kind="interpolated", base_world=True, settings=None):
    """Atlas with a weighted SpatialAccelerationTask on one hand (in the hand frame), bound to an SE3PDController."""
    mech, low, ctrl, qnom = scenarios.atlas_standing(settings or OSQPSettings.test_suite())
    names = list(mech.names)
    hand = names.index("r_hand") if "r_hand" in names else mech.nb - 1
    pelvis = low.program.floating_body
    base = -1 if base_world else pelvis
    task = SpatialAccelerationTask(mech, base, hand, hand)
    ti = low.addtask(task, 5.0)
    quintic = T.fit_quintic(x0=0.0, xf=1.0, y0=0.0, yd0=0.0, ydd0=0.0, yf=1.0, ydf=0.0, yddf=0.0)
    q0, q1, q2 = quat([1, 0, 0], 0.3), quat([0.2, 1, -0.4], 1.1), quat([0, 0, 1], -0.7)
    p0, p1, p2 = np.array([0.3, -0.4, 0.2]), np.array([0.5, -0.2, 0.6]), np.array([0.1, -0.5, 0.4])
    if kind == "interpolated":
        ang = T.Interpolated(0.2, 1.7, q0, q1, quintic, rotation=True)
        lin = T.Interpolated(0.2, 1.7, p0, p1, quintic)
    elif kind == "linear_alpha":  # identity interpolator, as in reference test/controller.jl:8-12
        ang = T.Interpolated(0.0, 1.0, q0, q1, rotation=True)
        lin = T.Interpolated(0.0, 1.0, p0, p1)
    elif kind == "piecewise":
        ang = T.Piecewise([T.Interpolated(0.0, 0.5, q0, q1, quintic, rotation=True), T.Constant(q1, rotation=True),
                           T.Interpolated(0.0, 0.8, q1, q2, quintic, rotation=True)], [0.1, 0.6, 0.9, 1.7])
        lin = T.Piecewise([T.Interpolated(0.0, 0.7, p0, p1, quintic), T.Interpolated(0.0, 0.9, p1, p2)], [0.0, 0.7, 1.6])
    else:
        raise ValueError(kind)
    traj = T.SE3Trajectory(hand, base, ang, lin)
    gains = SE3PDGains(PDGains(100.0, np.array([20.0, 15.0, 25.0])),
                       PDGains(np.array([[1000.0, 30, 0], [30, 900, -20], [0, -20, 800]]), 200.0))
    se3 = SE3PDController(base, hand, traj, np.diag([0, 0, 0, 10.0, 10, 10]), gains)
    low.bind_se3pd(ti, se3)
    off = low.program.des_offsets()[ti]
    return mech, low, ctrl, qnom, se3, off


def states(mech, qnom, B, seed):
    q, v = scenarios.atlas_random_states(mech, qnom, B, seed=seed)
    return q, v


@pytest.mark.parametrize("kind,base_world", [("interpolated", True), ("interpolated", False), ("linear_alpha", False),
                                             ("piecewise", True)])
def test_device_se3pd_matches_the_host_functor_emulated(kind, base_world):
    from emu import emu
    mech, low, ctrl, qnom, se3, off = make(kind, base_world)
    B = 48
    q, v = states(mech, qnom, B, 5)
    t = np.linspace(-0.3, 2.1, B)  # before, inside and after every piece (trajectories clamp)
    ec = emu.EmuController(low.program)
    got = ec.assemble(q, v, time=t)["desired"][:, off:off + 6]
    want = np.stack([se3(float(t[i]), mech, q[i:i + 1], v[i:i + 1])[0] for i in range(B)])
    assert np.max(np.abs(got - want)) <= 1e-9 * max(1.0, np.max(np.abs(want)))
    # one time for the whole batch
    got1 = ec.assemble(q, v, time=0.9)["desired"][:, off:off + 6]
    want1 = se3(0.9, mech, q, v)
    assert np.max(np.abs(got1 - want1)) <= 1e-9 * max(1.0, np.max(np.abs(want1)))
    # the other desireds are untouched
    other = np.delete(ec.assemble(q, v, time=t)["desired"], np.s_[off:off + 6], axis=1)
    ref = np.delete(ec.assemble(q, v, time=t + 1.0)["desired"], np.s_[off:off + 6], axis=1)
    assert np.array_equal(other, ref)


def test_trajectory_and_gains_are_refs_emulated():
    """controller.trajectory[] / controller.gains[] may be replaced between ticks (se3pdcontroller.jl:4-6)"""
    from emu import emu
    mech, low, ctrl, qnom, se3, off = make("interpolated", True)
    q, v = states(mech, qnom, 8, 6)
    ec = emu.EmuController(low.program)
    a = ec.assemble(q, v, time=0.8)["desired"][:, off:off + 6]
    se3.gains = SE3PDGains(PDGains(50.0, 5.0), PDGains(300.0, 30.0))
    se3.trajectory = T.SE3Trajectory(se3.body, se3.base, T.Constant(quat([0, 1, 0], 0.4), rotation=True),
                                     T.Interpolated(0.0, 2.0, np.zeros(3), np.ones(3)))
    b = ec.assemble(q, v, time=0.8)["desired"][:, off:off + 6]
    want = se3(0.8, mech, q, v)
    assert np.max(np.abs(b - want)) <= 1e-9 * max(1.0, np.max(np.abs(want)))
    assert np.max(np.abs(a - b)) > 1e-3


def test_unsupported_references_are_rejected():
    mech, low, ctrl, qnom, se3, off = make("interpolated", True)
    se3.trajectory = T.SE3Trajectory(se3.body, se3.base, T.Interpolated(0, 1, quat([1, 0, 0], 0), quat([1, 0, 0], 1), clamp=False,
                                                                       rotation=True), T.Constant(np.zeros(3)))
    from qpcontrol_jl_b200 import _lib
    with pytest.raises(ValueError):
        _lib.Handles._se3pd_args(se3)
    other = next(i for i, e in enumerate(low.program.tasks) if not isinstance(e.task, SpatialAccelerationTask))
    with pytest.raises(ValueError):
        low.bind_se3pd(other, se3)


@pytest.mark.gpu
@pytest.mark.parametrize("kind", ["interpolated", "piecewise"])
def test_device_se3pd_through_the_c_abi(kind):
    """qpc_solve_batch with the device-side SE3PD == the same tick with the host functor's output passed as `desired`"""
    mech, low, ctrl, qnom, se3, off = make(kind, kind == "piecewise")
    B = 256
    q, v = states(mech, qnom, B, 7)
    t = np.linspace(-0.2, 2.0, B)
    dev = low.finalize()
    res = low(q, v, time=t, check=False)
    a = dev.assemble_host(q, v, time=t)
    want = np.stack([se3(float(t[i]), mech, q[i:i + 1], v[i:i + 1])[0] for i in range(B)])
    assert np.max(np.abs(a["desired"][:, off:off + 6] - want)) <= 1e-9 * max(1.0, np.max(np.abs(want)))
    # a second controller without the binding, fed the host-evaluated desireds
    mech2, low2, ctrl2, _, se3b, off2 = make(kind, kind == "piecewise")
    low2.program.se3pd.clear()
    des = np.tile(low2.program.default_desired(), (B, 1))
    des[:, off2:off2 + 6] = want
    res2 = low2(q, v, desired=des, check=False)
    acc, acc2 = (res.status == 1) | (res.status == 2), (res2.status == 1) | (res2.status == 2)
    ok = acc & acc2
    assert ok.mean() > 0.95
    # the two desireds differ in rounding (1e-12): accept / reject must agree; OPTIMAL vs ALMOST_OPTIMAL at the iteration
    # limit may flip for an instance that converges slowly
    assert np.array_equal(acc, acc2)
    assert np.max(np.abs(res.tau[ok] - res2.tau[ok])) <= 1e-6 * max(1.0, np.max(np.abs(res2.tau[ok])))


@pytest.mark.gpu
def test_closed_loop_advances_the_controller_time():
    """qpc_step_batch evaluates the SE3PD at time + k dt at tick k: two single-tick calls == one two-tick call, bit for bit"""
    mech, low, ctrl, qnom, se3, off = make("interpolated", True)
    q, v = states(mech, qnom, 32, 8)
    dt, t0 = 0.05, 0.4
    qa, va, ra = low.simulate(q, v, dt, 2, time=t0, check=False)
    q1, v1, _ = low.simulate(q, v, dt, 1, time=t0, check=False)
    qb, vb, rb = low.simulate(q1, v1, dt, 1, time=t0 + dt, check=False)
    assert np.array_equal(qa, qb) and np.array_equal(va, vb) and np.array_equal(ra.tau, rb.tau)
    qc, vc, rc = low.simulate(q1, v1, dt, 1, time=t0, check=False)  # wrong time: the result must differ
    assert not np.array_equal(rc.tau, rb.tau)


@pytest.mark.gpu
def test_trajectory_and_gains_are_refs_on_the_device():
    """qpc_se3pd_update through the C ABI: replacing controller.trajectory / controller.gains between ticks re-uploads the
    program tables; the desired the assembly kernel writes follows the host functor before and after the swap."""
    mech, low, ctrl, qnom, se3, off = make("interpolated", True)
    q, v = states(mech, qnom, 64, 6)
    dev = low.finalize()
    a = dev.assemble_host(q, v, time=0.8)["desired"][:, off:off + 6]
    want_a = se3(0.8, mech, q, v)
    assert np.max(np.abs(a - want_a)) <= 1e-9 * max(1.0, np.max(np.abs(want_a)))
    se3.gains = SE3PDGains(PDGains(50.0, 5.0), PDGains(300.0, 30.0))
    se3.trajectory = T.SE3Trajectory(se3.body, se3.base, T.Constant(quat([0, 1, 0], 0.4), rotation=True),
                                     T.Interpolated(0.0, 2.0, np.zeros(3), np.ones(3)))
    b = dev.assemble_host(q, v, time=0.8)["desired"][:, off:off + 6]
    want_b = se3(0.8, mech, q, v)
    assert np.max(np.abs(b - want_b)) <= 1e-9 * max(1.0, np.max(np.abs(want_b)))
    assert np.max(np.abs(a - b)) > 1e-3
