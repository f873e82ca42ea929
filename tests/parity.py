"""Shared parity metrics (north_star: torques and contact wrenches within 1e-5 relative, identical active sets)."""
import numpy as np

REL_TOL = 1e-5


def rel_err(a, b):
    a, b = np.asarray(a).reshape(len(a), -1), np.asarray(b).reshape(len(b), -1)
    if a.shape[1] == 0:
        return np.zeros(len(a))
    return np.abs(a - b).max(1) / np.maximum(1.0, np.abs(b).max(1))


def active_sets(wrenches, program, mech_R=None):
    """Solution active set: contacts whose normal-direction force exceeds 1e-6 * max(1, total)  (SURVEY 8(c))."""
    f = np.linalg.norm(wrenches[:, :, 3:], axis=2)
    tot = np.maximum(1.0, f.sum(1, keepdims=True))
    return f > 1e-6 * tot


def assert_tick_parity(res, ref, program, tol=REL_TOL):
    """res: BatchResult (device or emulation); ref: oracle dict."""
    ok_ref = (ref["status"] == 1) | (ref["status"] == 2)
    ok_res = (res.status == 1) | (res.status == 2)
    assert np.array_equal(ok_ref, ok_res), "accept/reject decision differs"
    k = ok_ref
    assert rel_err(res.tau[k], ref["tau"][k]).max(initial=0) < tol
    assert rel_err(res.vdot[k], ref["vd"][k]).max(initial=0) < tol
    assert rel_err(res.wrenches[k], ref["wrenches"][k]).max(initial=0) < tol
    if res.wrenches.shape[1]:
        assert np.array_equal(active_sets(res.wrenches[k], program), active_sets(ref["wrenches"][k], program))
