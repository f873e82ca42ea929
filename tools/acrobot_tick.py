#!/usr/bin/env python
"""A few Acrobot ticks (BASELINE config 2) on cuda:0, device-resident -- the command ncu wraps for the thread-per-instance
kernel capture.   python tools/acrobot_tick.py [B] [reps]"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import torch
import qpc_loader
qpc = qpc_loader.load()
from qpcontrol_jl_b200 import scenarios

B = int(sys.argv[1]) if len(sys.argv) > 1 else 1 << 20
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 3
mech, low, task = scenarios.acrobot_point_task()
q, v, des = scenarios.acrobot_random_inputs(mech, B, seed=2)
dev = low.finalize()
dev.reserve(B)
cuda = torch.device("cuda", 0)
t = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(cuda)
dq, dv, dd = t(q), t(v), t(des)
nv = dev.dims["nv"]
out = dict(tau=torch.empty(B, nv, dtype=torch.float64, device=cuda), vdot=torch.empty(B, nv, dtype=torch.float64, device=cuda),
           status=torch.empty(B, dtype=torch.int32, device=cuda), iters=torch.empty(B, dtype=torch.int32, device=cuda))
stream = torch.cuda.current_stream().cuda_stream
ms = []
for _ in range(reps):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    dev.solve_device(B, dq, dv, out, desired=dd, stream=stream)
    e1.record()
    torch.cuda.synchronize()
    ms.append(e0.elapsed_time(e1))
print(f"acrobot B={B} ms per tick {ms} -> {B / (min(ms) * 1e-3) / 1e6:.1f} M solves/s; iters mean {out['iters'].double().mean().item():.1f}")
