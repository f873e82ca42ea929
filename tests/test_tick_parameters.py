"""Per-tick Parameters beyond desireds / contact weight / maxnormalforce: Parameter-valued task
weights (momentum.jl:107-110) and Parameter-valued contact position / normal / mu (contacts.jl:39,53-61; exercised by
the reference in test/controller.jl:42-47,110-118), one set per instance.  CPU: kernel bodies (tests/emu) vs oracle;
GPU: the CUDA path through the C ABI vs oracle."""
import numpy as np
import pytest

import qpc_loader

qpc = qpc_loader.load()
from qpcontrol_jl_b200 import OSQPSettings, scenarios  # noqa: E402

import parity  # noqa: E402
import util  # noqa: E402


def _inputs(B, seed):
    st = OSQPSettings.test_suite()
    mech, low, ctrl, qnom = scenarios.atlas_standing(st)
    q, v = scenarios.atlas_random_states(mech, qnom, B, seed=seed)
    rng = np.random.default_rng(seed)
    prog = low.program
    # task weights: only the weighted task (linear momentum rate) reacts; the others' entries are ignored
    tw = np.array([e.weight for e in prog.tasks])[None, :] * rng.uniform(0.5, 2.0, (B, len(prog.tasks)))
    # contact geometry: positions moved by up to 1 cm, normals tilted by up to ~6 degrees, mu in [0.6, 1.0]
    cg = np.zeros((B, len(prog.contacts), 7))
    for c, cp in enumerate(prog.contacts):
        cg[:, c, 0:3] = np.asarray(cp.position) + rng.uniform(-0.01, 0.01, (B, 3))
        n = np.asarray(cp.normal, dtype=np.float64) + rng.uniform(-0.1, 0.1, (B, 3))
        cg[:, c, 3:6] = n  # deliberately not normalised: rotation_between normalises (contacts.jl:11)
        cg[:, c, 6] = rng.uniform(0.6, 1.0, B)
    return mech, low, ctrl, q, v, tw, cg


def test_emu_matches_oracle_with_per_tick_parameters(orc):
    from emu import emu
    mech, low, ctrl, q, v, tw, cg = _inputs(24, seed=51)
    ref = orc.OracleController(low.program).solve_batch(q, v, task_weight=tw, contact_geometry=cg)
    res = emu.EmuController(low.program).solve(q, v, task_weight=tw, contact_geometry=cg)
    parity.assert_tick_parity(res, ref, low.program)
    # the parameters matter: the same states with the setup-time values give different torques
    base = orc.OracleController(low.program).solve_batch(q, v)
    assert parity.rel_err(base["tau"], ref["tau"]).max() > 1e-3


def test_setup_values_passed_as_parameters_change_nothing(orc):
    """Passing the setup-time weights / geometry explicitly reproduces the static program (device contact table ==
    host-compiled table)."""
    from emu import emu
    st = OSQPSettings.test_suite()
    mech, low, ctrl, qnom = scenarios.atlas_standing(st)
    q, v = scenarios.atlas_random_states(mech, qnom, 8, seed=52)
    prog = low.program
    tw = np.array([e.weight for e in prog.tasks])
    cg = np.array([list(cp.position) + list(cp.normal) + [cp.mu] for cp in prog.contacts])
    e = emu.EmuController(prog)
    a = e.solve(q, v)
    b = e.solve(q, v, task_weight=tw, contact_geometry=cg)
    assert parity.rel_err(a.tau, b.tau).max() < 1e-9
    assert np.array_equal(a.status, b.status)


def test_world_fixed_normal_parameter(orc):
    """test/controller.jl:42-47: a contact normal given as a per-tick Parameter that keeps it fixed in the WORLD while
    the body rotates: the contact force of a body hanging from a revolute joint stays along world z."""
    from qpcontrol_jl_b200 import MomentumBasedController
    from qpcontrol_jl_b200.mechanism import _Builder, REVOLUTE
    b = _Builder()
    b.add("body", "rx", None, REVOLUTE, axis=(1, 0, 0), mass=10.0, inertia=(1.0, 1.0, 1.0))
    mech = b.build()
    ctrl = MomentumBasedController(mech, OSQPSettings.test_suite())
    c = ctrl.addcontact(0, (0.0, 0.0, 0.0), (0.0, 0.0, 1.0), 1.0)
    c.maxnormalforce, c.weight = 1e3, 1e-3
    ctrl.regularize(0, 1.0)
    th = np.linspace(-1.2, 1.2, 7)
    q, v = th[:, None].copy(), np.zeros((7, 1))
    cg = np.zeros((7, 1, 7))
    for i, t in enumerate(th):
        R = util.expm_so3(np.array([t, 0, 0]))
        cg[i, 0, 3:6] = R.T @ np.array([0.0, 0.0, 1.0])  # world z expressed in the body frame
        cg[i, 0, 6] = 1.0
    oc = orc.OracleController(ctrl.program)
    P, qv, A, l, u = oc.lifted_qp(q[3], v[3])  # static program sanity
    ref = oc.solve_batch(q, v, contact_geometry=cg)
    assert np.all(ref["status"] == 1)
    f = ref["wrenches"][:, 0, 3:6]
    nz = np.linalg.norm(f, axis=1) > 1e-9
    # friction cone around world z with mu = 1: the force never points below the horizontal by more than 45 degrees
    assert np.all(f[nz][:, 2] >= np.linalg.norm(f[nz][:, :2], axis=1) - 1e-6)


@pytest.mark.gpu
def test_gpu_matches_oracle_with_per_tick_parameters(orc):
    mech, low, ctrl, q, v, tw, cg = _inputs(96, seed=53)
    ref = orc.OracleController(low.program).solve_batch(q, v, task_weight=tw, contact_geometry=cg)
    res = ctrl.lowlevel(q, v, task_weight=tw, contact_geometry=cg)
    parity.assert_tick_parity(res, ref, low.program)
    # broadcast rows (stride 0) == the same row repeated
    res_b = ctrl.lowlevel(q, v, task_weight=tw[0], contact_geometry=cg[0])
    res_r = ctrl.lowlevel(q, v, task_weight=np.tile(tw[0], (96, 1)), contact_geometry=np.tile(cg[0], (96, 1, 1)))
    assert np.array_equal(res_b.tau, res_r.tau)


# ---- Parameter-valued MATRIX weights (momentum.jl:113-117; test/controller.jl:232-285 modes 4 and 5) -----------------------
def _matrix_weight_setup(seed=31):
    from qpcontrol_jl_b200 import MomentumBasedController, SpatialAccelerationTask, LinearMomentumRateTask
    from qpcontrol_jl_b200.mechanism import rand_floating_humanoid
    rng = np.random.default_rng(seed)
    mech = rand_floating_humanoid(rng)
    B = 6
    q = np.stack([mech.rand_configuration(rng) for _ in range(B)])
    v = 0.3 * rng.standard_normal((B, mech.nv))
    W6 = rng.random((B, 6, 6))
    W6 = W6 @ W6.transpose(0, 2, 1) + np.eye(6)
    W3 = rng.random((B, 3, 3))            # deliberately unsymmetric: the Hessian block is W + W'
    W3 = W3 + 2.0 * np.eye(3)

    def build(Wa, Wb):
        ctrl = MomentumBasedController(mech, OSQPSettings.test_suite(), floatingjoint=0)
        t1 = SpatialAccelerationTask(mech, mech.findbody("r_hand"), mech.findbody("l_foot"), frame=mech.findbody("r_hand"))
        t2 = LinearMomentumRateTask(mech)
        ctrl.addtask(t1, Wa)
        ctrl.addtask(t2, Wb)
        for j in range(mech.nb):
            ctrl.regularize(j, 0.1)
        t1.setdesired(np.linspace(-0.3, 0.4, 6))
        t2.setdesired(mech.total_mass * mech.gravity * 0.0 + np.array([1.0, -2.0, 0.5]))
        return ctrl
    return mech, q, v, W6, W3, build


def _check_matrix_parameters(solve_of):
    mech, q, v, W6, W3, build = _matrix_weight_setup()
    B = len(q)
    # one controller whose setup-time matrices are placeholders, driven by per-tick matrices ...
    ctrl = build(np.eye(6), np.eye(3))
    twm = np.concatenate([W6.reshape(B, 36), W3.reshape(B, 9)], axis=1)
    res = solve_of(ctrl)(q, v, task_weight_matrix=twm)
    assert np.all(res.status == 1)
    # ... must equal, instance by instance, controllers BUILT with those matrices
    for i in range(B):
        ref = solve_of(build(W6[i], W3[i]))(q[i:i + 1], v[i:i + 1])
        assert ref.status[0] == 1
        np.testing.assert_allclose(res.vdot[i], ref.vdot[0], rtol=0, atol=1e-9)
        np.testing.assert_allclose(res.tau[i], ref.tau[0], rtol=0, atol=1e-7)
    # a broadcast row = the same matrices for every instance
    one = solve_of(ctrl)(q, v, task_weight_matrix=twm[0])
    np.testing.assert_allclose(one.vdot[0], res.vdot[0], rtol=0, atol=1e-12)
    # and the placeholder matrices give a different answer (the Parameter is really read)
    plain = solve_of(ctrl)(q, v)
    assert np.abs(plain.vdot - res.vdot).max() > 1e-4
    # shape validation: rows must match the batch
    with pytest.raises(ValueError):
        solve_of(ctrl)(q, v, task_weight_matrix=twm[:3])
    with pytest.raises(ValueError):
        solve_of(ctrl)(q, v, task_weight_matrix=twm[:, :40])


def test_matrix_weight_parameters_emulation():
    from emu import emu
    _check_matrix_parameters(lambda ctrl: (lambda q, v, **kw: emu.EmuController(ctrl.program).solve(q, v, **kw)))


@pytest.mark.gpu
def test_matrix_weight_parameters_gpu():
    _check_matrix_parameters(lambda ctrl: (lambda q, v, **kw: ctrl(q, v, check=False, **kw)))
