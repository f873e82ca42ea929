"""The closed loop WITH A PLANT: notebooks/Standing controller.ipynb:202-220 runs
simulate(state, 10., PeriodicController(tau, dt, controller)) and checks that Atlas keeps standing.  qpc_simulate_batch does
that for a batch: control tick (warm-started), then forward dynamics vd = M^-1 (tau - c + J'f) under a soft ground contact,
semi-implicit Euler.  CPU: the forward-dynamics body (kin.cuh, emulation) against the oracle's mass matrix / RNEA; GPU: the
notebook's assertions for 32 perturbed robots."""
import numpy as np
import pytest

import util
from emu import emu
from qpcontrol_jl_b200 import OSQPSettings, center_of_mass_host, scenarios


def test_forward_dynamics_body_matches_oracle(orc):
    """test/controller.jl:92-96 uses RigidBodyDynamics.dynamics! as the check of the controller; here the device's own
    forward dynamics is checked first: M(q) vd + c(q, v) = tau against the oracle's CRBA / RNEA, no contact."""
    mech, low, ctrl, qnom = scenarios.atlas_standing(OSQPSettings.test_suite())
    q, v = scenarios.atlas_random_states(mech, qnom, 6, seed=3)
    rng = np.random.default_rng(0)
    tau = 20.0 * rng.standard_normal((6, mech.nv))
    tau[:, :6] = 0.0
    vd, fc = emu.EmuController(low.program).forward_dynamics(q, v, tau)
    assert np.all(fc == 0.0)
    st = orc.OracleState(orc.OracleMechanism(mech))
    for i in range(6):
        st.set(q[i], v[i])
        ref = np.linalg.solve(st.mass_matrix(), tau[i] - st.inverse_dynamics(np.zeros(mech.nv)))
        np.testing.assert_allclose(vd[i], ref, rtol=1e-9, atol=1e-9 * np.abs(ref).max())
        np.testing.assert_allclose(st.inverse_dynamics(vd[i]), tau[i], atol=1e-8)


def test_soft_ground_contact_forces(orc):
    """Points below the ground plane are pushed up by k * penetration (at rest), points above carry nothing, and the
    acceleration changes by M^-1 J'f."""
    mech, low, ctrl, qnom = scenarios.atlas_standing(OSQPSettings.test_suite())
    q = qnom[None, :].copy()
    v = np.zeros((1, mech.nv))
    tau = np.zeros((1, mech.nv))
    fk = util.forward_kinematics(mech, qnom)
    z = np.array([(fk[c.body][0] @ np.asarray(c.position) + fk[c.body][1])[2] for c in low.program.contacts])
    e = emu.EmuController(low.program)
    k = 4e4
    vd_free, f_free = e.forward_dynamics(q, v, tau, k=k, d=0.0, ground_z=z.min() - 1.0)
    vd_c, f_c = e.forward_dynamics(q, v, tau, k=k, d=0.0, ground_z=z.min() + 0.002)
    assert np.all(f_free == 0.0)
    np.testing.assert_allclose(f_c[0, :, 2], k * np.maximum(0.0, z.min() + 0.002 - z), rtol=1e-9, atol=1e-9)
    np.testing.assert_allclose(f_c[0, :, :2], 0.0, atol=1e-12)          # at rest: no tangential force
    # Newton's law for the whole robot: linear momentum rate = m g + sum of the contact forces (momentum.jl:162-193)
    st = orc.OracleState(orc.OracleMechanism(mech)).set(q[0], v[0])
    A, hb = st.momentum_matrix(), st.momentum_rate_bias()
    for vd, f in ((vd_free, f_free), (vd_c, f_c)):
        np.testing.assert_allclose(A[3:] @ vd[0] + hb[3:], mech.total_mass * mech.gravity + f[0].sum(0), atol=1e-7)


@pytest.mark.gpu
def test_atlas_keeps_standing_on_the_soft_ground():
    """Standing controller.ipynb:215-220 for 32 robots started off the nominal posture: every tick accepted, the robots
    come to rest (|v| -> 0) and stay upright with the centre of mass above 1 m (the controller lowers it by 5 cm from the
    nominal 1.10 m, standing.jl:28)."""
    st = OSQPSettings.standing_notebook()
    mech, low, ctrl, qnom = scenarios.atlas_standing(st)
    B = 32
    rng = np.random.default_rng(1)
    q = np.tile(qnom, (B, 1))
    v = np.zeros((B, mech.nv))
    o = int(mech.qoff[mech.findjoint("pelvis_to_world")])
    mask = np.ones(mech.nq, bool)
    mask[o:o + 7] = False
    q[:, mask] += rng.normal(0.0, 0.02, (B, int(mask.sum())))
    fk = util.forward_kinematics(mech, qnom)
    ground = min((fk[c.body][0] @ np.asarray(c.position) + fk[c.body][1])[2] for c in low.program.contacts)
    dev = low.finalize()
    dev.set_warm_start(True)
    dev.reset_warm_start()
    vmax = []
    for block in range(6):  # 6 x 0.5 s, control at 500 Hz, plant at 4 kHz
        q, v, res = dev.simulate_host(q, v, 2e-3, 250, ground_z=ground, substeps=8)
        assert np.all((res.status == 1) | (res.status == 2)), (block, res.status)
        vmax.append(float(np.abs(v).max()))
    dev.set_warm_start(False)
    com = np.array([center_of_mass_host(mech, q[i]) for i in range(B)])
    assert np.all(com[:, 2] > 1.0) and np.all(com[:, 2] < 1.10)
    assert vmax[-1] < 5e-3 and vmax[-1] < 0.05 * vmax[0]                # at rest
    assert np.all(np.abs(q[:, o]) > 0.999)                               # pelvis upright (quaternion w)
    assert res.iters.mean() < 15                                         # warm-started ticks at the fixed point
