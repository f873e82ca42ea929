"""CPU-only: the ONE-WARP ADMM kernel body (csrc/admm_warp.cuh) run on 32 CPU fibres by tests/emu/warp_emu.cpp, against
the oracle (different formulation: lifted 143 x 178 QP, OSQP with sparse LDL').  The kernel reduces the device QP to the
friction-cone multipliers by a Householder QR of the equality block and iterates a 32 x 32 operator; these tests pin its
reduction, its per-row-rho ADMM, its termination test on the full problem, its infeasibility certificate, its warm start
and its hand-back codes before GPU minutes are spent (tests/test_gpu_parity.py repeats the parity through the C ABI)."""
import numpy as np
import pytest

import parity
from emu import emu
from qpcontrol_jl_b200 import OSQPSettings, scenarios


def _assembled(settings, B, seed, masks=None):
    mech, low, ctrl, qnom = scenarios.atlas_standing(settings)
    q, v = scenarios.atlas_random_states(mech, qnom, B, seed=seed)
    cw = cm = None
    if masks is not None:
        cm = scenarios.contact_masks(B, 8, p=masks, seed=seed)
        cw = np.full((B, 8), 1e-3)
    return mech, low, q, v, cw, cm


def test_tick_through_the_warp_body_matches_oracle(orc):
    """reference test/controller.jl:87-96 restated for a batch: torques / accelerations / contact wrenches of the tick
    whose QP the warp body solved agree with the oracle to 1e-5 relative at the test-suite OSQP settings."""
    mech, low, q, v, cw, cm = _assembled(OSQPSettings.test_suite(), 24, 3)
    res = emu.EmuController(low.program).solve_warp(q, v)
    ref = orc.OracleController(low.program).solve_batch(q, v)
    assert np.all(res.fallback == 0) and np.all(res.status == 1)
    parity.assert_tick_parity(res, ref, low.program)
    assert np.all(res.tau[:, :6] == 0.0)
    assert np.all(res.residuals < 1e-8)      # OSQP's criterion at eps_abs = 1e-8, eps_rel = 1e-16
    assert res.iters.max() < 2000             # the KKT-system form needs up to 8000 on these states


def test_contact_masks_through_the_warp_body(orc):
    """config 4 (test/controller.jl:188-215): disabled contacts are l = u = 0 rows (rho_eq = 1e3 rho) of the reduced QP."""
    mech, low, q, v, cw, cm = _assembled(OSQPSettings.test_suite(), 24, 4, masks=0.75)
    res = emu.EmuController(low.program).solve_warp(q, v, None, cw, cm)
    ref = orc.OracleController(low.program).solve_batch(q, v, cweight=cw, cmaxnf=cm)
    assert np.all(res.fallback == 0)
    parity.assert_tick_parity(res, ref, low.program)
    ok = (res.status == 1) | (res.status == 2)
    assert np.abs(res.wrenches[ok][(cm == 0)[ok]]).max(initial=0) < 1e-6


def test_notebook_tolerance_meets_osqp_criteria_on_the_full_problem():
    """At eps = 1e-5 (notebooks/Standing controller.ipynb:66-71) the reported residuals satisfy OSQP's termination test
    evaluated independently, in numpy, on the 53-variable QP the assembly stage wrote."""
    st = OSQPSettings.standing_notebook()
    mech, low, q, v, cw, cm = _assembled(st, 16, 3)
    a = emu.EmuController(low.program).assemble(q, v)
    w = emu.warp_solve_qp_batch(a["P"], a["q"], a["G"], a["lg"], a["lb"], a["ub"], settings=st)
    assert np.all(w["status"] == 1) and np.all(w["fallback"] == 0)
    na = a["P"].shape[1] - a["lb"].shape[1]
    for i in range(16):
        x, yb = w["x"][i], w["y"][i][a["lg"].shape[1]:]
        P, qv, G, b = a["P"][i], a["q"][i], a["G"][i], a["lg"][i]
        xb = x[na:]
        z = np.clip(xb, a["lb"][i], a["ub"][i])
        # primal: equality rows hold to rounding, box rows within tolerance
        rp = max(np.abs(G @ x - b).max(), np.abs(xb - z).max())
        ps = max(np.abs(G @ x).max(), np.abs(xb).max(), np.abs(z).max())
        assert np.abs(G @ x - b).max() < 1e-7 * max(1.0, np.abs(b).max())
        assert rp < st.eps_abs + st.eps_rel * ps
        # dual: multipliers of the equality rows by least squares on the stationarity condition
        g = P @ x + qv
        g[na:] += yb
        nu = np.linalg.lstsq(G.T, -g, rcond=None)[0]
        rd = np.abs(g + G.T @ nu).max()
        ds = max(np.abs(P @ x).max(), np.abs(g - (P @ x + qv) + G.T @ nu).max(), np.abs(qv).max())
        assert rd < 1.05 * (st.eps_abs + st.eps_rel * ds)
        assert abs(rp - w["res"][i][0]) <= 1e-6 * max(1.0, rp) + 1e-9 or w["res"][i][0] >= rp  # reported >= clipped-z residual


def test_infeasible_contact_sets_are_certified_or_rejected_like_the_oracle(orc):
    """One enabled contact point (or none) cannot balance the robot (the moment rows A3 rho = b3 have no solution inside the
    cone): accept / reject must agree with the oracle, and rejected instances carry OSQP's primal-infeasible status
    (momentum.jl:83-91 then throws QPSolveFailure)."""
    mech, low, q, v, cw, cm = _assembled(OSQPSettings.test_suite(), 24, 11, masks=0.75)
    for i in range(4):      # one contact point of the first foot only
        cm[i] = 0.0
        cm[i, i] = 1e6
    cm[4:8] = 0.0           # no contact at all: nothing can carry the weight while the feet are pinned (standing.jl:37-38)
    for i in range(8, 16):  # two points: feasible again
        cm[i] = 0.0
        cm[i, [i - 8, (i - 5) % 8]] = 1e6
    res = emu.EmuController(low.program).solve_warp(q, v, None, cw, cm)
    ref = orc.OracleController(low.program).solve_batch(q, v, cweight=cw, cmaxnf=cm)
    ok_ref = (ref["status"] == 1) | (ref["status"] == 2)
    ok_res = (res.status == 1) | (res.status == 2)
    assert (~ok_ref).sum() >= 8, "workload should contain infeasible instances"
    assert np.array_equal(ok_ref, ok_res), (ref["status"], res.status)
    both = ok_ref & ok_res
    assert parity.rel_err(res.tau[both], ref["tau"][both]).max(initial=0) < 1e-5
    assert set(np.unique(res.status[~ok_res])) <= {-3, 3}, res.status
    assert res.iters[~ok_res].max() <= 500   # certified quickly, not by running into the iteration limit


def test_hand_back_codes():
    """What the reduction cannot handle is handed back with a reason code, never solved wrongly: infinite bounds (1),
    rank-deficient equality block (2)."""
    st = OSQPSettings.test_suite()
    mech, low, q, v, cw, cm = _assembled(st, 2, 3)
    a = emu.EmuController(low.program).assemble(q, v)
    ub = a["ub"].copy()
    ub[0, 3] = 1e30
    w = emu.warp_solve_qp_batch(a["P"], a["q"], a["G"], a["lg"], a["lb"], ub, settings=st)
    assert w["fallback"][0] == 1 and w["status"][0] == -99 and w["fallback"][1] == 0 and w["status"][1] == 1
    G = a["G"].copy()
    G[1, :, 5] = G[1, :, 4]  # two identical x_a columns: G_a loses rank
    w = emu.warp_solve_qp_batch(a["P"], a["q"], G, a["lg"], a["lb"], a["ub"], settings=st)
    assert w["fallback"][1] == 2 and w["fallback"][0] == 0


def test_warm_start_from_the_solution_stops_at_the_first_check():
    st = OSQPSettings.standing_notebook()
    mech, low, q, v, cw, cm = _assembled(st, 8, 3)
    a = emu.EmuController(low.program).assemble(q, v)
    cold = emu.warp_solve_qp_batch(a["P"], a["q"], a["G"], a["lg"], a["lb"], a["ub"], settings=st)
    warm = emu.warp_solve_qp_batch(a["P"], a["q"], a["G"], a["lg"], a["lb"], a["ub"], settings=st,
                                   warm=dict(x=cold["x"], y=cold["y"], rho=cold["rho"]))
    assert np.all(cold["status"] == 1) and np.all(warm["status"] == 1)
    assert np.all(cold["rho"] > 0)
    assert warm["iters"].max() <= 10 and warm["iters"].mean() < cold["iters"].mean() / 4
    assert np.abs(warm["x"] - cold["x"]).max() < 1e-2 * max(1.0, np.abs(cold["x"]).max())


def test_second_shape_standing_plus_weighted_hand_task(orc):
    """The (MG, NA) = (30, 27) instantiation: the standing program plus one weighted 6-row SpatialAccelerationTask (a hand,
    as an SE3PDController would drive it): six more slack variables and six more equality rows in the block the QR
    eliminates.  Tick through the warp body vs the oracle at the test-suite settings."""
    from qpcontrol_jl_b200 import SpatialAccelerationTask
    mech, low, ctrl, qnom = scenarios.atlas_standing(OSQPSettings.test_suite())
    hand = list(mech.names).index("r_hand")
    ti = low.addtask(SpatialAccelerationTask(mech, -1, hand, hand), 5.0)
    off = low.program.des_offsets()[ti]
    B = 16
    q, v = scenarios.atlas_random_states(mech, qnom, B, seed=9)
    des = np.tile(low.program.default_desired(), (B, 1))
    des[:, off:off + 6] = np.random.default_rng(9).normal(0.0, 0.5, (B, 6))
    ec = emu.EmuController(low.program)
    assert (ec.h.mg, ec.h.n - ec.h.nbox) == (30, 27)
    res = ec.solve_warp(q, v, des)
    ref = orc.OracleController(low.program).solve_batch(q, v, desired=des)
    assert np.all(res.fallback == 0) and np.all(res.status == 1)
    parity.assert_tick_parity(res, ref, low.program)
