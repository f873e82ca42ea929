"""GPU tests of the drop-in boundary: the single-process multi-device entry point, several live controllers
on one device, page-locked vs pageable host buffers, and no allocation in steady state on the global-scratch QP path."""
import numpy as np
import pytest

from qpcontrol_jl_b200 import OSQPSettings, _lib, scenarios

pytestmark = pytest.mark.gpu


def _replicas(n, settings):
    """n replicas of the Atlas standing program, spread over the visible devices (all on device 0 on a 1-GPU box: the
    sharding / threading logic is the same)."""
    ndev = _lib.load().qpc_device_count()
    out = []
    for k in range(n):
        mech, low, ctrl, qnom = scenarios.atlas_standing(settings, device=k % ndev)
        out.append((mech, low, ctrl, qnom))
    return out


def test_solve_batch_multi_is_bit_identical_for_1_2_4_8_shards():
    """'results bit-identical for G = 1, 2, 4, 8' through ONE call from one process
    (qpc_solve_batch_multi: one host thread per controller / device)."""
    st = OSQPSettings.standing_notebook()
    reps = _replicas(8, st)
    mech, low, ctrl, qnom = reps[0]
    B = 1001  # not a multiple of 8: ragged shards
    q, v = scenarios.atlas_random_states(mech, qnom, B, seed=21)
    cm = scenarios.contact_masks(B, 8, seed=21)
    cw = np.full_like(cm, 1e-3)
    single = ctrl(q, v, cw, cm, check=False)
    devs = [r[1].finalize() for r in reps]
    for G in (1, 2, 4, 8):
        multi = _lib.solve_host_multi(devs[:G], q, v, contact_weight=cw, contact_maxnormalforce=cm)
        assert np.array_equal(multi.tau, single.tau), G
        assert np.array_equal(multi.wrenches, single.wrenches), G
        assert np.array_equal(multi.status, single.status) and np.array_equal(multi.iters, single.iters), G
    with pytest.raises(RuntimeError):
        _lib.solve_host_multi([devs[0], devs[0]], q, v)  # the same handle twice


def test_two_live_controllers_of_different_size_on_one_device():
    """Finalising a second, smaller controller must not lower the first one's dynamic shared-memory limits (they are per
    function and per device): interleaved ticks of both keep working."""
    mech, low, ctrl, qnom = scenarios.atlas_standing(OSQPSettings.standing_notebook())
    q, v = scenarios.atlas_random_states(mech, qnom, 64, seed=5)
    first = ctrl(q, v)
    amech, alow, atask = scenarios.acrobot_point_task()
    aq, av, ades = scenarios.acrobot_random_inputs(amech, 64, seed=2)
    a1 = alow(aq, av, ades)
    again = ctrl(q, v)
    a2 = alow(aq, av, ades)
    assert np.array_equal(first.tau, again.tau) and np.array_equal(a1.tau, a2.tau)
    # and the KKT kernels of the big controller (the hand-back path) still launch
    dev = low.finalize()
    dev.set_admm_warp(False)
    kkt = ctrl(q, v)
    dev.set_admm_warp(True)
    assert np.all(kkt.status == 1)


def test_pinned_and_pageable_host_buffers_give_identical_results():
    mech, low, ctrl, qnom = scenarios.atlas_standing(OSQPSettings.standing_notebook())
    B = 4608  # chunked path
    q, v = scenarios.atlas_random_states(mech, qnom, B, seed=17)
    pageable = ctrl(q, v)
    qp, vp = q.copy(), v.copy()
    _lib.pin_host_buffer(qp)
    _lib.pin_host_buffer(vp)
    try:
        pinned = ctrl(qp, vp)
    finally:
        _lib.unpin_host_buffer(qp)
        _lib.unpin_host_buffer(vp)
    assert np.array_equal(pageable.tau, pinned.tau) and np.array_equal(pageable.iters, pinned.iters)
    _lib.unpin_host_buffer(qp)  # unpinning twice is harmless


def test_global_scratch_qp_path_does_not_allocate_in_steady_state():
    """QPs whose matrices exceed shared memory run with a per-device scratch that is allocated once and kept."""
    import torch
    P, qv, A, l, u = scenarios.synthetic_qps(4, 200, 200, seed=5)
    st = OSQPSettings(eps_abs=1e-6, eps_rel=1e-6, max_iter=2000)
    n, m = 200, 200
    cuda = torch.device("cuda")
    dP, dq, dA, dl, du = (torch.from_numpy(np.ascontiguousarray(a)).to(cuda) for a in (P, qv, A, l, u))
    dx = torch.zeros(4, n, dtype=torch.float64, device=cuda)
    dy = torch.zeros(4, m, dtype=torch.float64, device=cuda)
    dstat = torch.zeros(4, dtype=torch.int32, device=cuda)
    diter = torch.zeros(4, dtype=torch.int32, device=cuda)
    dres = torch.zeros(4, 2, dtype=torch.float64, device=cuda)

    def run():
        _lib.solve_qp_batch_device(4, n, m, 0, dP, dq, dA, dl, du, None, None, st, dx, dy, dstat, diter, dres)
        torch.cuda.synchronize()

    run()
    free0, _ = torch.cuda.mem_get_info()
    run()
    run()
    free1, _ = torch.cuda.mem_get_info()
    assert free1 == free0
    assert np.all(dstat.cpu().numpy() == 1)


def test_direct_writes_into_pinned_outputs_equal_the_staged_copies():
    """Page-locked OUTPUT buffers are written by the epilogue kernel itself (api.cu: mapped_alias); pageable ones go
    through the staging buffer and a D2H copy per chunk.  Same bits, chunked path, odd batch."""
    import torch
    mech, low, ctrl, qnom = scenarios.atlas_standing(OSQPSettings.standing_notebook())
    B = 5001
    q, v = scenarios.atlas_random_states(mech, qnom, B, seed=19)
    dev = low.finalize()
    staged = dev.solve_host(q, v)  # fresh numpy arrays: pageable
    h = dev.h
    pin = lambda shape, dt=torch.float64: torch.empty(*shape, dtype=dt).pin_memory().numpy()  # noqa: E731
    from qpcontrol_jl_b200 import BatchResult
    res = BatchResult(tau=pin((B, h.nv)), vdot=pin((B, h.nv)), wrenches=pin((B, h.ncontacts, 6)),
                      status=pin((B,), torch.int32), iters=pin((B,), torch.int32), residuals=pin((B, 2)))
    res.tau[:] = np.nan
    res.wrenches[:] = np.nan
    dev.solve_host_into(q, v, res)
    assert np.array_equal(res.tau, staged.tau) and np.array_equal(res.vdot, staged.vdot)
    assert np.array_equal(res.wrenches, staged.wrenches) and np.array_equal(res.status, staged.status)
    assert np.array_equal(res.iters, staged.iters) and np.array_equal(res.residuals, staged.residuals)
