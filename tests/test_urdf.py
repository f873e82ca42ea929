"""URDF import: `parse_urdf` + fixed-joint removal reproduce hand-built mechanisms, and a
URDF-loaded robot runs through the controller (oracle on CPU, CUDA path on GPU)."""
import numpy as np
import pytest

import qpc_loader

qpc = qpc_loader.load()
from qpcontrol_jl_b200 import (MomentumBasedController, OSQPSettings, PointAccelerationTask, acrobot, parse_urdf,  # noqa: E402
                               center_of_mass_host)
from qpcontrol_jl_b200.mechanism import QUAT_FLOATING, REVOLUTE, PRISMATIC, _Builder  # noqa: E402
from qpcontrol_jl_b200.urdf import rpy_to_rot  # noqa: E402

import parity  # noqa: E402

ACROBOT_URDF = """<?xml version="1.0"?>
<robot name="Acrobot">
  <link name="base_link"/>
  <link name="upper_link">
    <inertial><origin xyz="0 0 -.5" rpy="0 0 0"/><mass value="1"/>
      <inertia ixx=".083" ixy="0" ixz="0" iyy=".083" iyz="0" izz=".001"/></inertial>
  </link>
  <link name="lower_link">
    <inertial><origin xyz="0 0 -1" rpy="0 0 0"/><mass value="1"/>
      <inertia ixx=".33" ixy="0" ixz="0" iyy=".33" iyz="0" izz=".001"/></inertial>
  </link>
  <joint name="shoulder" type="continuous">
    <parent link="base_link"/><child link="upper_link"/><origin xyz="0 .15 0"/><axis xyz="0 1 0"/>
  </joint>
  <joint name="elbow" type="continuous">
    <parent link="upper_link"/><child link="lower_link"/><origin xyz="0 .1 -1"/><axis xyz="0 1 0"/>
  </joint>
</robot>"""

CHAIN_URDF = """<robot name="chain">
  <link name="base"><inertial><origin xyz="0 0 0.1"/><mass value="5"/>
    <inertia ixx="0.1" iyy="0.2" izz="0.3" ixy="0.01" ixz="0" iyz="0"/></inertial></link>
  <link name="a"><inertial><origin xyz="0.1 0 -0.2" rpy="0.3 -0.2 0.5"/><mass value="2"/>
    <inertia ixx="0.02" iyy="0.03" izz="0.01" ixy="0" ixz="0.002" iyz="0"/></inertial></link>
  <link name="tool"><inertial><origin xyz="0 0.05 0"/><mass value="0.7"/>
    <inertia ixx="0.004" iyy="0.005" izz="0.006" ixy="0" ixz="0" iyz="0.001"/></inertial></link>
  <link name="b"><inertial><origin xyz="0 0 -0.3"/><mass value="1.5"/>
    <inertia ixx="0.05" iyy="0.05" izz="0.002" ixy="0" ixz="0" iyz="0"/></inertial></link>
  <joint name="j1" type="revolute"><parent link="base"/><child link="a"/>
    <origin xyz="0 0.1 0.4" rpy="0.1 0.2 -0.3"/><axis xyz="0 0 2"/><limit lower="-1" upper="1" effort="1" velocity="1"/></joint>
  <joint name="weld" type="fixed"><parent link="a"/><child link="tool"/><origin xyz="0.2 0 -0.4" rpy="0 0.5 0"/></joint>
  <joint name="j2" type="prismatic"><parent link="tool"/><child link="b"/><origin xyz="0 0 -0.1" rpy="0.2 0 0"/>
    <axis xyz="1 0 0"/></joint>
</robot>"""


def test_acrobot_urdf_matches_builtin_model():
    m, ref = parse_urdf(ACROBOT_URDF), acrobot()
    assert m.names == ref.names and m.joint_names == ref.joint_names
    for f in ("parent", "jtype"):
        assert np.array_equal(getattr(m, f), getattr(ref, f))
    for f in ("axis", "X_R", "X_p", "mass", "com", "inertia_com"):
        np.testing.assert_allclose(getattr(m, f), getattr(ref, f), atol=1e-15)


def test_fixed_joint_is_merged_and_children_reattached():
    m = parse_urdf(CHAIN_URDF, floating=True)
    assert m.names == ["base", "a", "b"] and m.joint_names == ["base_to_world", "j1", "j2"]
    assert list(m.jtype) == [QUAT_FLOATING, REVOLUTE, PRISMATIC] and list(m.parent) == [-1, 0, 1]
    assert m.nq == 9 and m.nv == 8
    np.testing.assert_allclose(m.axis[1], [0, 0, 1])  # normalised
    assert abs(m.mass[1] - 2.7) < 1e-15 and abs(m.total_mass - (5 + 2 + 0.7 + 1.5)) < 1e-12
    # j2's pose is now given in a's frame: weld o j2
    Rw, pw = rpy_to_rot([0, 0.5, 0]), np.array([0.2, 0, -0.4])
    np.testing.assert_allclose(m.X_R[2], Rw @ rpy_to_rot([0.2, 0, 0]), atol=1e-15)
    np.testing.assert_allclose(m.X_p[2], Rw @ np.array([0, 0, -0.1]) + pw, atol=1e-15)
    # composite inertia of (a + tool) about the composite centre of mass equals the two-body sum
    Ra = rpy_to_rot([0.3, -0.2, 0.5])
    Ia = Ra @ np.array([[0.02, 0, 0.002], [0, 0.03, 0], [0.002, 0, 0.01]]) @ Ra.T
    It = Rw @ np.array([[0.004, 0, 0], [0, 0.005, 0.001], [0, 0.001, 0.006]]) @ Rw.T
    ca, ct = np.array([0.1, 0, -0.2]), Rw @ np.array([0, 0.05, 0]) + pw
    c = (2 * ca + 0.7 * ct) / 2.7
    par = lambda mm, d: mm * (d @ d * np.eye(3) - np.outer(d, d))  # noqa: E731
    np.testing.assert_allclose(m.com[1], c, atol=1e-15)
    np.testing.assert_allclose(m.inertia_com[1], Ia + par(2, ca - c) + It + par(0.7, ct - c), atol=1e-15)


def test_merged_model_has_the_dynamics_of_the_unmerged_one(orc):
    """Same robot built by hand with `tool` as a separate (massless-jointed) body is not expressible without fixed
    joints, so compare against first principles: total mass, centre of mass and kinetic energy of the rigid pair."""
    m = parse_urdf(CHAIN_URDF, floating=False)
    assert list(m.jtype) == [REVOLUTE, PRISMATIC] and list(m.parent) == [-1, 0]
    om = orc.OracleMechanism(m)
    s = orc.OracleState(om)
    rng = np.random.default_rng(0)
    q, v = rng.uniform(-1, 1, m.nq), rng.standard_normal(m.nv)
    s.set(q, v)
    M = s.mass_matrix()
    assert np.allclose(M, M.T) and np.all(np.linalg.eigvalsh(M) > 0)
    # kinetic energy of body `a + tool` spinning about j1 (v[1] = 0): 1/2 w^2 (z' I_O z) with I_O about the joint axis
    v1 = np.array([1.3, 0.0])
    s.set(q, v1)
    z = np.array([0, 0, 1.0])
    I_O = m.inertia_com[0] + m.mass[0] * (m.com[0] @ m.com[0] * np.eye(3) - np.outer(m.com[0], m.com[0]))
    # plus body b carried rigidly (prismatic joint locked): its inertia about the same axis, expressed in a's frame
    Rb, pb = m.X_R[1], m.X_p[1] + m.X_R[1] @ (np.array([1.0, 0, 0]) * q[1])
    cb = Rb @ m.com[1] + pb
    I_b = Rb @ m.inertia_com[1] @ Rb.T + m.mass[1] * (cb @ cb * np.eye(3) - np.outer(cb, cb))
    ke = 0.5 * 1.3 ** 2 * (z @ (I_O + I_b) @ z)
    assert abs(0.5 * v1 @ s.mass_matrix() @ v1 - ke) < 1e-12


def test_rejects_unsupported_content():
    with pytest.raises(ValueError):
        parse_urdf(ACROBOT_URDF.replace('type="continuous"', 'type="planar"', 1))
    with pytest.raises(ValueError):
        parse_urdf("<robot><link name='a'/><link name='b'/></robot>")  # two roots


def test_urdf_robot_through_the_oracle_controller(orc):
    mech = parse_urdf(ACROBOT_URDF)
    low = MomentumBasedController(mech, OSQPSettings.acrobot_notebook())
    low.addtask(PointAccelerationTask(mech, -1, mech.nb - 1, (0.0, 0.0, -2.05)))
    for j in range(mech.nb):
        low.regularize(j, 1e-6)
    from qpcontrol_jl_b200 import scenarios
    q, v, des = scenarios.acrobot_random_inputs(mech, 32, seed=2)
    ref = orc.OracleController(low.program).solve_batch(q, v, desired=des)
    mech2, low2, task = scenarios.acrobot_point_task()
    ref2 = orc.OracleController(low2.program).solve_batch(q, v, desired=des)
    assert np.all(ref["status"] == 1)
    assert parity.rel_err(ref["tau"], ref2["tau"]).max() < 1e-12


@pytest.mark.gpu
def test_gpu_urdf_floating_chain(orc):
    """A URDF-loaded floating robot (fixed joint merged) through the CUDA path: torques match the oracle."""
    from qpcontrol_jl_b200 import JointAccelerationTask, MomentumRateTask
    mech = parse_urdf(CHAIN_URDF, floating=True)
    ctrl = MomentumBasedController(mech, OSQPSettings.test_suite(), floatingjoint=0)
    for pos in ([0.1, 0.1, -0.1], [-0.1, 0.1, -0.1], [0.1, -0.1, -0.1], [-0.1, -0.1, -0.1]):
        c = ctrl.addcontact(0, pos, (0.0, 0.0, 1.0), 0.8)
        c.maxnormalforce, c.weight = 1e4, 1e-3
    ctrl.addtask(MomentumRateTask(mech), 1.0)
    for j in (1, 2):
        ctrl.addtask(JointAccelerationTask(mech, j))
        ctrl.regularize(j, 0.05)
    ctrl.regularize(0, 0.05)
    rng = np.random.default_rng(3)
    q = np.stack([mech.rand_configuration(rng) for _ in range(16)])
    q[:, :4] = [1, 0, 0, 0]
    v = 0.1 * rng.standard_normal((16, mech.nv))
    res = ctrl(q, v)
    ref = orc.OracleController(ctrl.program).solve_batch(q, v)
    parity.assert_tick_parity(res, ref, ctrl.program)
