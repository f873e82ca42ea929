"""CPU-only: the N>1 host logic with world_size-2 gloo process groups -- contiguous shards, no
data-path collective, max-over-ranks timing, rank-0 gather.  The per-shard "solver" here is the CPU oracle (the
checker), so the test also pins that sharded results equal the unsharded ones bit for bit."""
import os
import socket
import sys

import numpy as np
import pytest
import torch.multiprocessing as mp

_ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if _ROOT not in sys.path:  # spawned workers import this module without conftest.py
    sys.path.insert(0, _ROOT)
import qpc_loader  # noqa: E402

qpc_loader.load()
from qpcontrol_jl_b200 import OSQPSettings, scenarios, sharding  # noqa: E402


def test_shard_ranges_partition_the_batch():
    for B in (0, 1, 7, 16, 16384, 65536 + 3):
        for world in (1, 2, 3, 4, 8):
            spans = [sharding.shard_range(B, r, world) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == B
            assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
            sizes = [hi - lo for lo, hi in spans]
            assert max(sizes) - min(sizes) <= 1
    with pytest.raises(ValueError):
        sharding.shard_range(8, 2, 2)


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _worker(rank, world, port, B, out_dir):
    os.environ.update(RANK=str(rank), WORLD_SIZE=str(world), LOCAL_RANK=str(rank), MASTER_ADDR="127.0.0.1",
                      MASTER_PORT=str(port))
    from emu import emu
    from qpcontrol_jl_b200 import sharding as sh
    sh.init_process_group("gloo")
    assert sh.env_rank_world() == (rank, world, rank)
    mech, low, ctrl, qnom = scenarios.atlas_standing(OSQPSettings.standing_notebook())
    q, v = scenarios.atlas_random_states(mech, qnom, B, seed=11)
    lo, hi = sh.shard_range(B, rank, world)
    sh.barrier()
    # the PRODUCT path's kernel bodies (kin.cuh assembly / inverse dynamics, admm_warp.cuh on CPU fibres), not the oracle
    res = emu.EmuController(low.program).solve_warp(q[lo:hi], v[lo:hi])
    tmax = sh.max_over_ranks([float(rank + 1), 0.5])
    assert tmax[0] == float(world)
    tau = sh.gather_rows(res.tau, B)
    if rank == 0:
        np.save(os.path.join(out_dir, "tau.npy"), tau)
    else:
        assert tau is None
    sh.barrier()
    import torch.distributed as dist
    dist.destroy_process_group()  # gloo's background threads must be torn down before the interpreter exits


def test_world_size_two_gloo_matches_unsharded(orc, tmp_path):
    B = 10
    port = _free_port()
    mp.spawn(_worker, args=(2, port, B, str(tmp_path)), nprocs=2, join=True)
    tau = np.load(tmp_path / "tau.npy")
    mech, low, ctrl, qnom = scenarios.atlas_standing(OSQPSettings.standing_notebook())
    q, v = scenarios.atlas_random_states(mech, qnom, B, seed=11)
    from emu import emu
    whole = emu.EmuController(low.program).solve_warp(q, v)
    assert np.array_equal(tau, whole.tau)  # bit-identical: the result of an instance does not depend on its shard
