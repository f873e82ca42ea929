"""Workload builders for the BASELINE.json configs; all data is synthetic and seeded.

  config 1/3/4  `atlas_standing()`            -- notebooks/Standing controller.ipynb:39-135
                `atlas_random_states()`       -- nominal state + seeded perturbations (config 3)
                `contact_masks()`             -- per-instance active contact sets (config 4, test/controller.jl:188-215)
  config 2      `acrobot_point_task()`        -- notebooks/PointAccelerationTask Demo.ipynb:135-188
  config 5      `synthetic_qps()`             -- dense QP sweep
"""
from __future__ import annotations

from typing import Optional, Tuple

import numpy as np

from .controller import MomentumBasedController, StandingController
from .mechanism import Mechanism, acrobot, atlas_like, atlas_nominal_configuration
from .program import OSQPSettings, PointAccelerationTask


def atlas_standing(settings: Optional[OSQPSettings] = None, device: int = 0, mech: Optional[Mechanism] = None):
    """The Atlas standing controller exactly as the notebook builds it: `MomentumBasedController{4}` with the
    floating joint `pelvis_to_world`, every contact point of every body added with normal (0,0,1) in the body frame,
    `maxnormalforce = 1e6`, `weight = 1e-3` (cells 5), then `StandingController(lowlevel, feet, pelvis, nominal)`
    with default gains (cell 7)."""
    mech = mech if mech is not None else atlas_like()
    settings = settings if settings is not None else OSQPSettings.standing_notebook()
    lowlevel = MomentumBasedController(mech, settings, N=4, floatingjoint=mech.findjoint("pelvis_to_world"),
                                       device=device)
    for body in range(mech.nb):
        for pos in mech.contact_points.get(body, ()):
            contact = lowlevel.addcontact(body, pos, (0.0, 0.0, 1.0), mech.contact_mu)
            contact.maxnormalforce = 1e6
            contact.weight = 1e-3
    feet = [mech.findbody("l_foot"), mech.findbody("r_foot")]
    pelvis = mech.findbody("pelvis")
    qnom = atlas_nominal_configuration(mech)
    controller = StandingController(lowlevel, feet, pelvis, qnom)
    return mech, lowlevel, controller, qnom


def atlas_random_states(mech: Mechanism, qnom: np.ndarray, B: int, seed: int = 3) -> Tuple[np.ndarray, np.ndarray]:
    """Config 3: joint angles + N(0, 0.05^2) rad, pelvis position + N(0, 0.02^2) m, pelvis orientation = small random
    rotation vector N(0, 0.05^2) applied to the nominal quaternion, v ~ N(0, 0.1^2)."""
    rng = np.random.default_rng(seed)
    q = np.tile(qnom, (B, 1))
    fj = mech.findjoint("pelvis_to_world")
    o = int(mech.qoff[fj])
    mask = np.ones(mech.nq, dtype=bool)
    mask[o:o + 7] = False
    q[:, mask] += rng.normal(0.0, 0.05, (B, int(mask.sum())))
    q[:, o + 4:o + 7] += rng.normal(0.0, 0.02, (B, 3))
    rv = rng.normal(0.0, 0.05, (B, 3))
    ang = np.linalg.norm(rv, axis=1, keepdims=True)
    dq = np.concatenate([np.cos(ang / 2), np.sin(ang / 2) * rv / np.maximum(ang, 1e-300)], axis=1)
    q[:, o:o + 4] = _quat_mul(q[:, o:o + 4], dq)
    q[:, o:o + 4] /= np.linalg.norm(q[:, o:o + 4], axis=1, keepdims=True)
    v = rng.normal(0.0, 0.1, (B, mech.nv))
    return np.ascontiguousarray(q), np.ascontiguousarray(v)


def _quat_mul(a, b):
    w1, x1, y1, z1 = a.T
    w2, x2, y2, z2 = b.T
    return np.stack([w1 * w2 - x1 * x2 - y1 * y2 - z1 * z2, w1 * x2 + x1 * w2 + y1 * z2 - z1 * y2,
                     w1 * y2 - x1 * z2 + y1 * w2 + z1 * x2, w1 * z2 + x1 * y2 - y1 * x2 + z1 * w2], axis=1)


def contact_masks(B: int, ncontacts: int, p: float = 0.75, min_enabled: int = 3, seed: int = 4,
                  maxnormalforce: float = 1e6) -> np.ndarray:
    """Config 4: each contact enabled with probability p, redrawn while fewer than `min_enabled` are enabled;
    disabled <=> maxnormalforce = 0 (`disable!`, contacts.jl:72).  Returns maxnormalforce [B, ncontacts]."""
    rng = np.random.default_rng(seed)
    on = rng.random((B, ncontacts)) < p
    bad = on.sum(axis=1) < min_enabled
    while bad.any():
        on[bad] = rng.random((int(bad.sum()), ncontacts)) < p
        bad = on.sum(axis=1) < min_enabled
    return np.where(on, maxnormalforce, 0.0)


# ---- config 2 ---------------------------------------------------------------------------------------------------
ACROBOT_CENTER = np.array([0.0, 0.25, 2.2])
ACROBOT_RADIUS = 0.5
ACROBOT_SPEED = 1.0
ACROBOT_POINT = (0.0, 0.0, -2.05)


def acrobot_point_task(settings: Optional[OSQPSettings] = None, device: int = 0):
    """Low-level controller of the PointAccelerationTask demo: hard point-acceleration task on the tip of the lower
    arm, regularisation 1e-6 on both joints, no contacts, fixed base (notebook cell 9)."""
    mech = acrobot()
    settings = settings if settings is not None else OSQPSettings.acrobot_notebook()
    lowlevel = MomentumBasedController(mech, settings, N=4, device=device)
    body = mech.nb - 1
    task = PointAccelerationTask(mech, -1, body, ACROBOT_POINT)
    lowlevel.addtask(task)
    for j in range(mech.nb):
        lowlevel.regularize(j, 1e-6)
    return mech, lowlevel, task


def acrobot_random_inputs(mech: Mechanism, B: int, seed: int = 2):
    """q ~ U(-pi, pi)^2 rejecting |sin q2| < 0.05, v ~ N(0,1)^2, t ~ U(0, 2 pi); desired from the notebook's PD law
    `pd(PDGains(1, 1), p, pref, pdot, pdotref) + pddref` on the circular reference (cells 7, 9)."""
    rng = np.random.default_rng(seed)
    q = rng.uniform(-np.pi, np.pi, (B, 2))
    bad = np.abs(np.sin(q[:, 1])) < 0.05
    while bad.any():
        q[bad] = rng.uniform(-np.pi, np.pi, (int(bad.sum()), 2))
        bad = np.abs(np.sin(q[:, 1])) < 0.05
    v = rng.standard_normal((B, 2))
    t = rng.uniform(0, 2 * np.pi, B)
    # planar forward kinematics of the tip (both joints about +y)
    a1, a2 = q[:, 0], q[:, 0] + q[:, 1]
    w1, w2 = v[:, 0], v[:, 0] + v[:, 1]
    l1, l2 = 1.0, -ACROBOT_POINT[2]
    # a point (0, 0, -r) rotated by angle a about +y is (-r sin a, 0, -r cos a)
    px = -l1 * np.sin(a1) - l2 * np.sin(a2)
    pz = -l1 * np.cos(a1) - l2 * np.cos(a2)
    p = np.stack([px, np.full(B, 0.25), pz], axis=1)
    pd = np.stack([-l1 * np.cos(a1) * w1 - l2 * np.cos(a2) * w2, np.zeros(B),
                   l1 * np.sin(a1) * w1 + l2 * np.sin(a2) * w2], axis=1)
    c, s = np.cos(ACROBOT_SPEED * t), np.sin(ACROBOT_SPEED * t)
    pref = ACROBOT_CENTER + ACROBOT_RADIUS * np.stack([c, np.zeros(B), s], axis=1)
    pdref = ACROBOT_SPEED * ACROBOT_RADIUS * np.stack([-s, np.zeros(B), c], axis=1)
    pddref = ACROBOT_SPEED ** 2 * ACROBOT_RADIUS * np.stack([-c, np.zeros(B), -s], axis=1)
    desired = -1.0 * (p - pref) - 1.0 * (pd - pdref) + pddref
    return np.ascontiguousarray(q), np.ascontiguousarray(v), np.ascontiguousarray(desired)


# ---- config 5 ---------------------------------------------------------------------------------------------------
def synthetic_qps(B: int, n: int, m: int, seed: int = 5):
    """P = M'M/n + 1e-3 I, A ~ N(0,1), q ~ N(0,1), l = A x0 - U(0,1), u = A x0 + U(0,1); first m//4 rows equalities."""
    rng = np.random.default_rng(seed)
    M = rng.standard_normal((B, n, n))
    P = np.einsum("bki,bkj->bij", M, M) / n + 1e-3 * np.eye(n)
    A = rng.standard_normal((B, m, n))
    q = rng.standard_normal((B, n))
    x0 = rng.standard_normal((B, n))
    Ax0 = np.einsum("bij,bj->bi", A, x0)
    l = Ax0 - rng.uniform(0, 1, (B, m))
    u = Ax0 + rng.uniform(0, 1, (B, m))
    ne = m // 4
    l[:, :ne] = Ax0[:, :ne]
    u[:, :ne] = Ax0[:, :ne]
    return tuple(np.ascontiguousarray(a) for a in (P, q, A, l, u))
