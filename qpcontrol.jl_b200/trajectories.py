"""Reference-signal generators of `QPControl.Trajectories` (reference src/trajectories/*.jl), restated for the host side.

SURVEY.md 8(f) rank 3, first half: the trajectories that feed `setdesired!` / the `desired` batch input.  They are
evaluated on the host (numpy; `x` may be a scalar or an array of evaluation times, one per robot instance) and their
outputs go into `qpc_batch_in.desired`.  The second half of rank 3, `SE3PDController` (src/lowlevel/se3pdcontroller.jl),
which adds `pd(::SE3PDGains, ...)` to `SE3Trajectory`'s feed-forward term, is in `se3pd.py`.

Call convention mirrors the reference: `traj(x)` returns the value, `traj(x, n)` returns `(value, d/dx, ..., d^n/dx^n)`
(`Val(n)` in Julia).  Rotations are unit quaternions (w, x, y, z); their derivatives are angular velocity /
acceleration vectors (the Lie derivative of interpolated.jl:75-82).
"""
from __future__ import annotations

from typing import Callable, Sequence, Tuple, Union

import numpy as np

__all__ = ["Polynomial", "fit_cubic", "fit_quintic", "Constant", "Interpolated", "Piecewise", "PointTrajectory",
           "FreeVectorTrajectory", "SE3Trajectory"]


class Polynomial:
    """StaticUnivariatePolynomials.Polynomial: coefficients in ascending order."""

    def __init__(self, coeffs: Sequence[float]):
        self.coeffs = np.asarray(coeffs, dtype=np.float64)

    def __call__(self, x, num_derivs: int = None):
        if num_derivs is None:
            return self._eval(self.coeffs, x)
        out, c = [], self.coeffs
        for _ in range(num_derivs + 1):
            out.append(self._eval(c, x))
            c = c[1:] * np.arange(1, len(c)) if len(c) > 1 else np.zeros(1)
        return tuple(out)

    @staticmethod
    def _eval(c, x):
        x = np.asarray(x, dtype=np.float64)
        y = np.zeros_like(x)
        for a in c[::-1]:  # Horner
            y = y * x + a
        return y if y.ndim else float(y)

    def derivative(self) -> "Polynomial":
        c = self.coeffs
        return Polynomial(c[1:] * np.arange(1, len(c)) if len(c) > 1 else [0.0])


def _coefficient_gradient(x: float, n: int, order: int) -> np.ndarray:
    """d^order/dx^order of (1, x, ..., x^(n-1)) -- SUP.coefficient_gradient."""
    g = np.zeros(n)
    for k in range(order, n):
        g[k] = np.prod(np.arange(k - order + 1, k + 1)) * x ** (k - order)
    return g


def fit_cubic(*, x0, xf, y0, yd0, yf, ydf) -> Polynomial:
    """fit_polynomial.jl:1-11."""
    A = np.stack([_coefficient_gradient(x0, 4, 0), _coefficient_gradient(x0, 4, 1), _coefficient_gradient(xf, 4, 0),
                  _coefficient_gradient(xf, 4, 1)])
    return Polynomial(np.linalg.solve(A, np.array([y0, yd0, yf, ydf], dtype=np.float64)))


def fit_quintic(*, x0, xf, y0, yd0, ydd0, yf, ydf, yddf) -> Polynomial:
    """fit_polynomial.jl:13-25."""
    A = np.stack([_coefficient_gradient(x0, 6, k) for k in range(3)] + [_coefficient_gradient(xf, 6, k) for k in range(3)])
    return Polynomial(np.linalg.solve(A, np.array([y0, yd0, ydd0, yf, ydf, yddf], dtype=np.float64)))


class DomainError(ValueError):
    pass


def _quat_mul(a, b):
    w1, x1, y1, z1 = np.moveaxis(np.asarray(a, dtype=np.float64), -1, 0)
    w2, x2, y2, z2 = np.moveaxis(np.asarray(b, dtype=np.float64), -1, 0)
    return np.stack([w1 * w2 - x1 * x2 - y1 * y2 - z1 * z2, w1 * x2 + x1 * w2 + y1 * z2 - z1 * y2,
                     w1 * y2 - x1 * z2 + y1 * w2 + z1 * x2, w1 * z2 + x1 * y2 - y1 * x2 + z1 * w2], axis=-1)


def quat_to_rot(q) -> np.ndarray:
    w, x, y, z = np.moveaxis(np.asarray(q, dtype=np.float64), -1, 0)
    return np.stack([np.stack([1 - 2 * (y * y + z * z), 2 * (x * y - w * z), 2 * (x * z + w * y)], -1),
                     np.stack([2 * (x * y + w * z), 1 - 2 * (x * x + z * z), 2 * (y * z - w * x)], -1),
                     np.stack([2 * (x * z - w * y), 2 * (y * z + w * x), 1 - 2 * (x * x + y * y)], -1)], -2)


class Constant:
    """constant.jl: value and zero derivatives (a rotation's derivatives are zero 3-vectors)."""

    def __init__(self, value, rotation: bool = False):
        self.value = value if np.isscalar(value) else np.asarray(value, dtype=np.float64)
        self.rotation = rotation

    def __call__(self, x, num_derivs: int = None):
        if num_derivs is None:
            return self.value
        zero = np.zeros(3) if self.rotation else (0.0 if np.isscalar(self.value) else np.zeros_like(self.value))
        return (self.value,) + tuple(zero for _ in range(num_derivs))


class Interpolated:
    """interpolated.jl: y(x) = y0 + alpha(theta(x)) (yf - y0), theta = clamp((x - x0) / (xf - x0), 0, 1), alpha the
    interpolator (identity or a Polynomial); geodesic interpolation when y0, yf are rotations (quaternions)."""

    def __init__(self, x0: float, xf: float, y0, yf, interpolator: Union[None, Polynomial, Callable] = None,
                 clamp: bool = True, rotation: bool = False):
        self.x0, self.xf = float(x0), float(xf)
        self.y0 = np.asarray(y0, dtype=np.float64)
        self.yf = np.asarray(yf, dtype=np.float64)
        self.interpolator = interpolator
        self.clamp = clamp
        self.rotation = rotation
        if rotation:  # AngleAxis(y0 \\ yf): axis and angle of the relative rotation (interpolated.jl:75-78)
            qc = self.y0 * np.array([1.0, -1.0, -1.0, -1.0])
            d = _quat_mul(qc, self.yf)
            if d[0] < 0:
                d = -d
            s = np.linalg.norm(d[1:])
            self.angle = 2.0 * np.arctan2(s, d[0])
            self.axis = d[1:] / s if s > 0 else np.array([1.0, 0.0, 0.0])

    def __call__(self, x, num_derivs: int = None):
        n = 0 if num_derivs is None else num_derivs
        xa = np.asarray(x, dtype=np.float64)
        if not self.clamp and (np.any(xa < self.x0) or np.any(xa > self.xf)):
            raise DomainError(f"Trajectory evaluated outside of range [{self.x0}, {self.xf}]")
        dx = self.xf - self.x0
        th = (xa - self.x0) / dx
        inside = (th > 0) & (th < 1)
        th = np.clip(th, 0.0, 1.0)
        dth = np.where(inside, 1.0 / dx, 0.0)
        if self.interpolator is None:  # identity: alpha = theta, alpha' = 1, higher derivatives 0
            al = (th,) + tuple(np.ones_like(th) if k == 1 else np.zeros_like(th) for k in range(1, n + 1))
        else:
            al = tuple(np.asarray(a, dtype=np.float64) for a in self.interpolator(th, n))
        al_x = [al[i] * dth ** i for i in range(1, n + 1)]  # chain rule, d^k theta / dx^k = 0 for k > 1
        if self.rotation:
            half = 0.5 * al[0] * self.angle
            dq = np.concatenate([np.cos(half)[..., None], np.sin(half)[..., None] * self.axis], axis=-1)
            y = _quat_mul(np.broadcast_to(self.y0, dq.shape), dq)
            dy = self.axis * self.angle  # Lie derivative
        else:
            dy = self.yf - self.y0
            y = self.y0 + al[0][..., None] * dy if dy.ndim else self.y0 + al[0] * dy
        if num_derivs is None:
            return y if np.ndim(y) else float(y)
        derivs = tuple((a[..., None] * dy if np.ndim(dy) else a * dy) for a in al_x)
        return (y,) + derivs


class Piecewise:
    """piecewise.jl: subfunction i is active on [breaks[i], breaks[i+1]) and is evaluated at x - breaks[i]."""

    def __init__(self, subfunctions: Sequence, breaks: Sequence[float], clamp: bool = True):
        b = np.asarray(breaks, dtype=np.float64)
        assert np.all(np.diff(b) >= 0) and len(b) == len(subfunctions) + 1
        self.subfunctions, self.breaks, self.clamp = list(subfunctions), b, clamp

    def __call__(self, x, *args):
        x0, xf = self.breaks[0], self.breaks[-1]
        if self.clamp:
            xc = min(max(x, x0), xf)
        else:
            if x < x0 or x > xf:
                raise DomainError(f"Trajectory evaluated outside of range [{x0}, {xf}]")
            xc = x
        index = min(int(np.searchsorted(self.breaks, xc, side="right")) - 1, len(self.subfunctions) - 1)
        return self.subfunctions[index](xc - self.breaks[index], *args)


class PointTrajectory:
    """point_vector.jl:1-13: a trajectory of 3-vectors tagged with the frame (body index, -1 = world) they live in."""

    def __init__(self, frame: int, trajectory):
        self.frame, self.trajectory = frame, trajectory

    def __call__(self, x, num_derivs: int = None):
        return self.trajectory(x) if num_derivs is None else self.trajectory(x, num_derivs)


class FreeVectorTrajectory(PointTrajectory):
    """point_vector.jl:16-26."""


class SE3Trajectory:
    """se3.jl: `angular` yields the rotation body -> base (quaternion) with derivatives expressed in base, `linear` the
    translation with derivatives in base.  Returns the pose (R [.,3,3], p), the twist of body w.r.t. base expressed in
    body (omega, nu) and the spatial acceleration in body (omega_dot, nu_dot) -- the feed-forward term SE3PDController
    adds its PD term to (se3pdcontroller.jl:13-17)."""

    def __init__(self, body: int, base: int, angular, linear):
        self.body, self.base, self.angular, self.linear = body, base, angular, linear

    def __call__(self, x, num_derivs: int = 2):
        if num_derivs != 2:
            raise ValueError("SE3Trajectory is evaluated with two derivatives (se3.jl:9)")
        quat, w_base, wd_base = self.angular(x, 2)
        p, pd, pdd = self.linear(x, 2)
        R = quat_to_rot(quat)
        Rt = np.swapaxes(R, -1, -2)
        rot = lambda M, v: np.einsum("...ij,...j->...i", M, v)  # noqa: E731
        w_body, nu_body = rot(Rt, w_base), rot(Rt, pd)
        wd_body = rot(Rt, wd_base)
        nud_body = rot(Rt, pdd) + np.cross(w_body, nu_body)
        return (R, p), (w_body, nu_body), (wd_body, nud_body)

    def desired_spatial_acceleration(self, x) -> np.ndarray:
        """(angular; linear) 6-vector(s) ready for a SpatialAccelerationTask's `desired` expressed in the body frame."""
        _, _, (wd, nud) = self(x, 2)
        return np.concatenate([wd, nud], axis=-1)
