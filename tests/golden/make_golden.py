#!/usr/bin/env python
"""Generates the golden fixtures in this directory from the CPU oracle (oracle/, fp64 restatement of the reference path).

The reference itself holds no golden vectors and cannot run here (Julia absent), so these fixtures pin
the ORACLE, not QPControl.jl: parity stays "unpinned" in the sense of DESIGN.md section 0.  They serve two purposes:
(1) the oracle cannot drift silently (tests/test_golden.py, CPU), (2) the GPU path is compared with committed numbers
and not only with whatever the oracle computes on the day (tests/test_golden.py, -m gpu).

    python tests/golden/make_golden.py        # rewrites the .npz files next to this script
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
import qpc_loader  # noqa: E402

qpc = qpc_loader.load()
from oracle import oracle as orc  # noqa: E402
from qpcontrol_jl_b200 import OSQPSettings, scenarios  # noqa: E402


def main():
    st = OSQPSettings.test_suite()
    # Atlas standing controller, BASELINE config 3 states (seed 3), first 32; and config 4 contact masks (seed 4)
    mech, low, ctrl, qnom = scenarios.atlas_standing(st)
    oc = orc.OracleController(low.program)
    oc.set_settings(st, warm_start=0)
    q, v = scenarios.atlas_random_states(mech, qnom, 32, seed=3)
    r = oc.solve_batch(q, v)
    np.savez_compressed(os.path.join(HERE, "atlas_standing_seed3.npz"), q=q, v=v, tau=r["tau"], vdot=r["vd"],
                        wrenches=r["wrenches"], status=r["status"])
    q, v = scenarios.atlas_random_states(mech, qnom, 32, seed=4)
    cm = scenarios.contact_masks(32, len(low.program.contacts), seed=4)
    cw = np.full_like(cm, 1e-3)
    oc.reset()
    r = oc.solve_batch(q, v, cweight=cw, cmaxnf=cm)
    np.savez_compressed(os.path.join(HERE, "atlas_contact_masks_seed4.npz"), q=q, v=v, cw=cw, cm=cm, tau=r["tau"],
                        vdot=r["vd"], wrenches=r["wrenches"], status=r["status"])
    # Acrobot PointAccelerationTask demo (BASELINE config 2 inputs, seed 2), first 64
    mech2, low2, task = scenarios.acrobot_point_task()
    q, v, des = scenarios.acrobot_random_inputs(mech2, 64, seed=2)
    r = orc.OracleController(low2.program).solve_batch(q, v, desired=des)
    np.savez_compressed(os.path.join(HERE, "acrobot_point_task_seed2.npz"), q=q, v=v, desired=des, tau=r["tau"],
                        vdot=r["vd"], status=r["status"])
    print("golden fixtures written to", HERE)


if __name__ == "__main__":
    main()
