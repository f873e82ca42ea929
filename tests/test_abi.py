"""CPU-only: the C-ABI library loads and exports every symbol include/qpcontrol_b200.h declares; setup-time calls work
and validate arguments without a GPU; compute entry points fail loudly (no CPU fallback)."""
import ctypes as C
import os
import re

import numpy as np
import pytest

from qpcontrol_jl_b200 import MomentumBasedController, OSQPSettings, _lib, scenarios

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    text = open(os.path.join(ROOT, "include", "qpcontrol_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(qpc_[a-z0-9_]+)\s*\(", text)))


def test_library_exports_every_declared_symbol():
    lib = _lib.load()
    syms = declared_symbols()
    assert len(syms) >= 20
    for s in syms:
        assert hasattr(lib, s), f"{s} declared in include/qpcontrol_b200.h but not exported"
    assert set(_lib.SETUP_SYMBOLS + _lib.COMPUTE_SYMBOLS) <= set(syms)


def test_setup_calls_validate_arguments():
    lib = _lib.load()
    assert lib.qpc_version() >= 100
    mech = scenarios.atlas_like()
    arrs = [np.ascontiguousarray(a) for a in (mech.parent, mech.jtype, mech.axis, mech.X_R, mech.X_p, mech.mass,
                                              mech.com, mech.inertia_origin(), mech.gravity)]
    h = C.c_void_p(lib.qpc_mechanism_create(C.c_int32(mech.nb), *[_lib._p(a) for a in arrs]))
    assert h
    nb, nq, nv = C.c_int32(), C.c_int32(), C.c_int32()
    assert lib.qpc_mechanism_dims(h, C.byref(nb), C.byref(nq), C.byref(nv)) == 0
    assert (nb.value, nq.value, nv.value) == (31, 37, 36)
    c = C.c_void_p(lib.qpc_controller_create(h, C.c_int32(4), C.c_int32(0), None))
    assert c
    pos = np.zeros(3)
    assert lib.qpc_add_contact(c, C.c_int32(99), _lib._p(pos), _lib._p(pos), C.c_double(0.8)) < 0
    assert b"out of range" in lib.qpc_last_error()
    assert lib.qpc_add_task(c, C.c_int32(9), 0, 0, 0, None, 0, 0, C.c_double(0), None) < 0
    assert lib.qpc_add_task(c, C.c_int32(0), C.c_int32(-1), C.c_int32(5), C.c_int32(5), None, C.c_int32(-1),
                            C.c_int32(2), C.c_double(0), None) < 0  # matrix weight missing
    assert lib.qpc_add_task(c, C.c_int32(4), C.c_int32(-1), C.c_int32(-1), C.c_int32(-1), None, C.c_int32(7),
                            C.c_int32(0), C.c_double(0), None) == 0
    # unfinalized controllers cannot solve
    dims = [C.c_int32() for _ in range(7)]
    assert lib.qpc_controller_dims(c, *[C.byref(d) for d in dims]) < 0
    lib.qpc_controller_destroy(c)
    lib.qpc_mechanism_destroy(h)
    # parent must precede child
    bad = arrs[0].copy()
    bad[3] = 10
    assert not lib.qpc_mechanism_create(C.c_int32(mech.nb), _lib._p(bad), *[_lib._p(a) for a in arrs[1:]])


def test_default_settings_are_osqp_defaults():
    lib = _lib.load()
    s = _lib.qpc_settings()
    lib.qpc_default_settings(C.byref(s))
    d = OSQPSettings()
    for f in ("rho", "sigma", "alpha", "eps_abs", "eps_rel", "max_iter", "scaling", "check_termination"):
        assert getattr(s, f) == getattr(d, f)


def test_compute_without_gpu_fails_loudly():
    lib = _lib.load()
    if lib.qpc_device_count() > 0:
        pytest.skip("a GPU is visible")
    mech, low, ctrl, qnom = scenarios.atlas_standing()
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        ctrl(qnom[None], np.zeros((1, mech.nv)))


def test_recorded_program_matches_reference_structure():
    """The host mirror records what the notebook builds: 8 contacts, 2 foot tasks, 1 weighted linear-momentum task,
    1 pelvis task, 18 joint tasks, regularisation 0.05 on all 36 velocities (standing.jl:35-49)."""
    mech, low, ctrl, qnom = scenarios.atlas_standing()
    pr = low.program
    assert len(pr.contacts) == 8 and all(c.isenabled() for c in pr.contacts)
    kinds = [e.task.kind for e in pr.tasks]
    assert kinds.count(0) == 2 and kinds.count(6) == 1 and kinds.count(1) == 1 and kinds.count(4) == 18
    assert np.allclose(pr.reg, 0.05)
    assert pr.ndes == 12 + 3 + 3 + 18
    pr.contacts[0].disable()
    assert not pr.contacts[0].isenabled()
    with pytest.raises(ValueError):
        low.addtask(pr.tasks[0].task, np.eye(5))


def test_new_entry_points_validate_arguments_without_a_gpu():
    """qpc_set_warm_start / qpc_reset_warm_start / qpc_step_batch refuse an unfinalized controller and bad arguments;
    the qpc_batch_in mirror has the header's field order (task_weight / contact_geometry appended)."""
    lib = _lib.load()
    mech = scenarios.atlas_like()
    arrs = [np.ascontiguousarray(a) for a in (mech.parent, mech.jtype, mech.axis, mech.X_R, mech.X_p, mech.mass,
                                              mech.com, mech.inertia_origin(), mech.gravity)]
    h = C.c_void_p(lib.qpc_mechanism_create(C.c_int32(mech.nb), *[_lib._p(a) for a in arrs]))
    c = C.c_void_p(lib.qpc_controller_create(h, C.c_int32(4), C.c_int32(0), None))
    assert lib.qpc_set_warm_start(c, C.c_int32(1)) < 0 and b"not finalized" in lib.qpc_last_error()
    assert lib.qpc_reset_warm_start(c) < 0
    q, v = np.zeros((1, mech.nq)), np.zeros((1, mech.nv))
    assert lib.qpc_step_batch(c, C.c_int64(1), _lib._p(q), _lib._p(v), None, None, C.c_double(1e-3), C.c_int32(1),
                              C.c_int32(0), None) < 0
    assert lib.qpc_step_batch(None, C.c_int64(1), _lib._p(q), _lib._p(v), None, None, C.c_double(1e-3), C.c_int32(1),
                              C.c_int32(0), None) < 0
    lib.qpc_controller_destroy(c)
    lib.qpc_mechanism_destroy(h)
    names = [f[0] for f in _lib.qpc_batch_in._fields_]
    text = open(os.path.join(ROOT, "include", "qpcontrol_b200.h")).read()
    body = text[text.index("typedef struct {", text.index("Inputs of one batched tick")):text.index("} qpc_batch_in;")]
    body = re.sub(r"/\*.*?\*/", "", body, flags=re.S)
    declared = re.findall(r"(\w+);", body)
    assert names == declared


def test_host_mirror_rejects_malformed_tick_parameters():
    mech, low, ctrl, qnom = scenarios.atlas_standing()
    from qpcontrol_jl_b200._lib import _prep_tick_parameters

    class H:  # the two attributes _prep_tick_parameters reads
        program = low.program
        ncontacts = len(low.program.contacts)
    with pytest.raises(ValueError):
        _prep_tick_parameters(H, np.zeros(3), None)
    with pytest.raises(ValueError):
        _prep_tick_parameters(H, None, np.zeros((8, 6)))
    tw, cg = _prep_tick_parameters(H, np.ones(len(low.program.tasks)), np.zeros((2, 8, 7)))
    assert tw.flags.c_contiguous and cg.shape == (2, 8, 7)
