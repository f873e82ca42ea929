#!/usr/bin/env python
"""Where the host-buffer tick (qpc_solve_batch, QPC_HOST_PTRS) spends its time beyond the kernels: the same 16,384-state
Atlas tick device-resident, from pinned host buffers with all outputs, with tau only, and the bare copies.
   python tools/e2e_probe.py"""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import torch
import ctypes as C
import qpc_loader
qpc = qpc_loader.load()
from qpcontrol_jl_b200 import OSQPSettings, scenarios, _lib

B = 16384
mech, low, ctrl, qnom = scenarios.atlas_standing(OSQPSettings.standing_notebook())
q, v = scenarios.atlas_random_states(mech, qnom, B, seed=3)
dev = low.finalize()
dev.reserve(B)
h = dev.h
nv, nc = h.nv, h.ncontacts
pin = lambda a: torch.from_numpy(np.ascontiguousarray(a)).pin_memory()
hq, hv = pin(q), pin(v)
res = qpc.BatchResult(tau=pin(np.empty((B, nv))).numpy(), vdot=pin(np.empty((B, nv))).numpy(), wrenches=pin(np.empty((B, nc, 6))).numpy(),
                      status=pin(np.empty(B, np.int32)).numpy(), iters=pin(np.empty(B, np.int32)).numpy(), residuals=pin(np.empty((B, 2))).numpy())

def timeit(fn, n=30, warm=5):
    for _ in range(warm): fn()
    torch.cuda.synchronize(); t0 = time.perf_counter()
    for _ in range(n): fn()
    torch.cuda.synchronize(); return 1e3 * (time.perf_counter() - t0) / n

full = timeit(lambda: dev.solve_host_into(hq.numpy(), hv.numpy(), res))
# tau only, straight through the C ABI (no Python marshalling beyond the struct)
bi = h.batch_in(hq.numpy(), hv.numpy(), None, None, None)
bo = _lib.qpc_batch_out(); bo.tau = res.tau.ctypes.data; bo.status = res.status.ctypes.data
call = lambda b_out: _lib.check(dev.lib, dev.lib.qpc_solve_batch(h.ctrl, C.c_int64(B), C.byref(bi), C.byref(b_out), C.c_int32(_lib.HOST_PTRS), None), "x")
tau_only = timeit(lambda: call(bo))
bo_all = _lib._batch_out(res)
raw_all = timeit(lambda: call(bo_all))
cuda = torch.device("cuda", 0)
dq, dv = hq.to(cuda), hv.to(cuda)
out = dict(tau=torch.empty(B, nv, dtype=torch.float64, device=cuda), vdot=torch.empty(B, nv, dtype=torch.float64, device=cuda),
           wrench=torch.empty(B, nc, 6, dtype=torch.float64, device=cuda), status=torch.empty(B, dtype=torch.int32, device=cuda),
           iters=torch.empty(B, dtype=torch.int32, device=cuda), residuals=torch.empty(B, 2, dtype=torch.float64, device=cuda))
stream = torch.cuda.current_stream().cuda_stream
devres = timeit(lambda: dev.solve_device(B, dq, dv, out, stream=stream))
def copies():
    dq.copy_(hq, non_blocking=True); dv.copy_(hv, non_blocking=True)
    for k, a in (("tau", res.tau), ("vdot", res.vdot), ("wrench", res.wrenches), ("status", res.status), ("iters", res.iters), ("residuals", res.residuals)):
        torch.from_numpy(a).copy_(out[k], non_blocking=True)
cp = timeit(copies)
print(f"device-resident tick {devres:.3f} ms | host buffers: python API {full:.3f}, C ABI all outputs {raw_all:.3f}, C ABI tau+status only {tau_only:.3f} | "
      f"bare copies (in {(hq.numel()+hv.numel())*8/1e6:.1f} MB, out {sum(a.nbytes for a in (res.tau,res.vdot,res.wrenches,res.status,res.iters,res.residuals))/1e6:.1f} MB) {cp:.3f} ms")
