import os, sys, time
sys.path.insert(0, os.getcwd())
import numpy as np, torch, qpc_loader
qpc = qpc_loader.load()
from qpcontrol_jl_b200 import OSQPSettings, scenarios
B = int(sys.argv[1])
mech, low, ctrl, qnom = scenarios.atlas_standing(OSQPSettings.standing_notebook())
q, v = scenarios.atlas_random_states(mech, qnom, B, seed=4)
cm = scenarios.contact_masks(B, 8, seed=4); cw = np.full_like(cm, 1e-3)
dev = low.finalize(); dev.reserve(B)
cuda = torch.device("cuda", 0)
t = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(cuda)
dq, dv, dcw, dcm = t(q), t(v), t(cw), t(cm)
nv = dev.dims["nv"]
out = dict(tau=torch.empty(B, nv, dtype=torch.float64, device=cuda), status=torch.empty(B, dtype=torch.int32, device=cuda))
stream = torch.cuda.current_stream().cuda_stream
flush = torch.empty(256 << 20, dtype=torch.uint8, device=cuda)
ms = []
for i in range(12):
    flush.zero_()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); dev.solve_device(B, dq, dv, out, contact_weight=dcw, contact_maxnormalforce=dcm, stream=stream); e1.record()
    torch.cuda.synchronize(); ms.append(e0.elapsed_time(e1))
print(f"B={B} nchunk={os.environ.get('QPC_NCHUNK','default')} median {np.median(ms[3:]):.3f} ms -> {B/np.median(ms[3:])/1e3:.2f} M solves/s")
