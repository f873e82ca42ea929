"""Restatement of the reference's test/trajectories.jl for qpcontrol.jl_b200/trajectories.py (ForwardDiff replaced by
central differences)."""
import numpy as np
import pytest

import qpc_loader

qpc = qpc_loader.load()
from qpcontrol_jl_b200.trajectories import (Constant, DomainError, FreeVectorTrajectory, Interpolated, Piecewise,  # noqa: E402
                                            PointTrajectory, SE3Trajectory, fit_cubic, fit_quintic, quat_to_rot)


def test_fit_cubic():  # test/trajectories.jl:17-39
    rng = np.random.default_rng(15)
    for _ in range(10):
        x0, xf, y0, yd0, yf, ydf = rng.random(6)
        if abs(xf - x0) < 0.2:
            xf = x0 + 0.5
        p = fit_cubic(x0=x0, xf=xf, y0=y0, yd0=yd0, yf=yf, ydf=ydf)
        pd = p.derivative()
        assert abs(p(x0) - y0) < 1e-6 and abs(pd(x0) - yd0) < 1e-6
        assert abs(p(xf) - yf) < 1e-6 and abs(pd(xf) - ydf) < 1e-6
        assert len(p.coeffs) == 4


def test_fit_quintic():  # test/trajectories.jl:41-69
    rng = np.random.default_rng(15)
    for _ in range(10):
        x0, xf, y0, yd0, ydd0, yf, ydf, yddf = rng.random(8)
        if abs(xf - x0) < 0.2:  # nearly coincident knots make the 6 x 6 system ill conditioned (the reference's draws avoid it)
            xf = x0 + 0.5
        p = fit_quintic(x0=x0, xf=xf, y0=y0, yd0=yd0, ydd0=ydd0, yf=yf, ydf=ydf, yddf=yddf)
        pd, pdd = p.derivative(), p.derivative().derivative()
        for got, want in ((p(x0), y0), (pd(x0), yd0), (pdd(x0), ydd0), (p(xf), yf), (pd(xf), ydf), (pdd(xf), yddf)):
            assert abs(got - want) < 1e-6
        assert len(p.coeffs) == 6


def test_interpolated_identity():  # test/trajectories.jl:71-107
    traj = Interpolated(0.0, 1.0, 1, 2)
    assert traj(0.5) == 1.5
    y, yd, ydd = traj(0.5, 2)
    assert (y, yd, ydd) == (1.5, 1.0, 0.0)
    traj = Interpolated(0.0, 1.0, [1, 2, 3], [0, 2, 4])
    y, yd, ydd = traj(0.5, 2)
    np.testing.assert_allclose(y, [0.5, 2.0, 3.5], atol=1e-15)
    np.testing.assert_allclose(yd, [-1.0, 0.0, 1.0], atol=1e-15)
    np.testing.assert_allclose(ydd, [0, 0, 0], atol=1e-15)
    angle, axis = np.pi / 2, np.array([1.0, 0.0, 0.0])
    y0 = np.array([1.0, 0, 0, 0])
    yf = np.concatenate([[np.cos(angle / 2)], np.sin(angle / 2) * axis])
    traj = Interpolated(0.0, 1.0, y0, yf, rotation=True)
    np.testing.assert_allclose(traj(0.0), y0, atol=1e-15)
    np.testing.assert_allclose(traj(1.0), yf, atol=1e-15)
    y, yd, ydd = traj(0.5, 2)
    np.testing.assert_allclose(y, np.concatenate([[np.cos(angle / 4)], np.sin(angle / 4) * axis]), atol=1e-15)
    np.testing.assert_allclose(yd / np.linalg.norm(yd), axis, atol=1e-15)
    assert abs(np.linalg.norm(yd) - abs(angle)) < 1e-15
    np.testing.assert_allclose(ydd, 0, atol=1e-15)


def test_interpolated_polynomial():  # test/trajectories.jl:109-132
    interp = fit_quintic(x0=0.0, xf=1.0, y0=0.0, yd0=0.0, ydd0=0.0, yf=1.0, ydf=0.0, yddf=0.0)
    x0, xf, y0, yf = -1.0, 2.0, 2.0, 3.0
    traj = Interpolated(x0, xf, y0, yf, interp)
    for xe, ye in ((x0, y0), (xf, yf)):
        y, yd, ydd = traj(xe, 2)
        assert abs(y - ye) < 1e-10 and abs(yd) < 1e-10 and abs(ydd) < 1e-10
    h = 1e-4
    for x in np.linspace(x0 + 0.05, xf - 0.05, 10):
        _, yd, ydd = traj(x, 2)
        assert abs(yd - (traj(x + h) - traj(x - h)) / (2 * h)) < 1e-4
        assert abs(ydd - (traj(x + h) - 2 * traj(x) + traj(x - h)) / h ** 2) < 1e-4
    # batched evaluation: one time per robot instance
    xs = np.linspace(x0, xf, 7)
    yb, ydb, yddb = traj(xs, 2)
    for k, x in enumerate(xs):
        y, yd, ydd = traj(float(x), 2)
        assert (yb[k], ydb[k], yddb[k]) == (y, yd, ydd)


def _check_piecewise(traj):  # test/trajectories.jl:134-157
    n = len(traj.subfunctions)
    for i, t in enumerate(traj.breaks):
        j = min(i, n - 1)
        assert traj(t) == traj.subfunctions[j](t - traj.breaks[j])
        if i < n:
            tmid = (t + traj.breaks[i + 1]) / 2
            assert traj(tmid) == traj.subfunctions[i](tmid - t)
            assert traj(tmid, 2) == traj.subfunctions[i](tmid - t, 2)
    t0, tf = traj.breaks[0], traj.breaks[-1]
    if traj.clamp:
        assert traj(t0 - 1) == traj(t0) and traj(tf + 1) == traj(tf)
    else:
        with pytest.raises(DomainError):
            traj(t0 - 1)
        with pytest.raises(DomainError):
            traj(tf + 1)


def test_piecewise_constant_and_interpolated():  # test/trajectories.jl:159-176
    n = 5
    breaks = [i ** 2 - 1 for i in range(1, n + 2)]
    subs = [Constant(i) for i in range(1, n + 1)]
    _check_piecewise(Piecewise(subs, breaks, clamp=True))
    _check_piecewise(Piecewise(subs, breaks, clamp=False))
    rng = np.random.default_rng(1)
    subs = [Interpolated(0.0, float(dt), rng.random(), rng.random()) for dt in np.diff(breaks)]
    _check_piecewise(Piecewise(subs, breaks, clamp=True))
    _check_piecewise(Piecewise(subs, breaks, clamp=False))


def test_point_and_free_vector_trajectories():  # test/trajectories.jl:178-211
    rng = np.random.default_rng(2)
    interp = fit_cubic(x0=0.0, xf=1.0, y0=rng.random(), yd0=rng.random(), yf=rng.random(), ydf=rng.random())
    inner = Interpolated(-1.0, 2.0, rng.random(3), rng.random(3), interp)
    for cls in (PointTrajectory, FreeVectorTrajectory):
        traj = cls(3, inner)
        assert traj.frame == 3
        np.testing.assert_array_equal(traj(1.0), inner(1.0))
        for a, b in zip(traj(1.0, 2), inner(1.0, 2)):
            np.testing.assert_array_equal(a, b)


def test_se3_trajectory_is_consistent_with_its_own_derivatives():  # test/trajectories.jl:213-233 + kinematic identities
    angle, axis = np.pi / 2, np.array([1.0, 0.0, 0.0])
    interp = fit_quintic(x0=0.0, xf=1.0, y0=0.0, yd0=0.0, ydd0=0.0, yf=1.0, ydf=0.0, yddf=0.0)
    angular = Interpolated(0.0, 1.0, [1.0, 0, 0, 0], np.concatenate([[np.cos(angle / 2)], np.sin(angle / 2) * axis]),
                           interp, rotation=True)
    linear = Interpolated(0.0, 1.0, [0.0, 1.0, 2.0], [2.0, 3.0, 4.0], interp)
    traj = SE3Trajectory(body=5, base=-1, angular=angular, linear=linear)
    (R, p), (w, nu), (wd, nud) = traj(0.4, 2)
    assert traj.body == 5 and traj.base == -1
    h = 1e-5
    (Rp, pp), (wp, nup), _ = traj(0.4 + h, 2)
    (Rm, pm), (wm, num), _ = traj(0.4 - h, 2)
    # body-frame twist: R' Rdot = hat(w), R' pdot = nu
    Rdot = (Rp - Rm) / (2 * h)
    W = R.T @ Rdot
    np.testing.assert_allclose([W[2, 1], W[0, 2], W[1, 0]], w, atol=1e-8)
    np.testing.assert_allclose(R.T @ (pp - pm) / (2 * h), nu, atol=1e-8)
    # spatial acceleration = time derivative of the body-frame twist components
    np.testing.assert_allclose((wp - wm) / (2 * h), wd, atol=1e-6)
    # se3.jl:19: nud = R' pdd + w x nu, and d/dt (R' pd) = R' pdd - w x nu
    np.testing.assert_allclose((nup - num) / (2 * h), nud - 2 * np.cross(w, nu), atol=1e-6)
    assert traj.desired_spatial_acceleration(np.array([0.1, 0.4])).shape == (2, 6)
    np.testing.assert_allclose(quat_to_rot(angular(1.0)) @ np.array([0, 1.0, 0]), [0, 0, 1.0], atol=1e-14)
