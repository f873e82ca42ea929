import sys; sys.path.insert(0,'.')
import qpc_loader; qpc=qpc_loader.load()
from qpcontrol_jl_b200 import _lib, scenarios, OSQPSettings
import numpy as np
for n,m in ((60,60),(64,64),(68,71)):
    P,qv,A,l,u=scenarios.synthetic_qps(4,n,m,seed=5)
    try:
        r=_lib.solve_qp_batch_host(P,qv,A,l,u,settings=OSQPSettings(eps_abs=1e-8,eps_rel=1e-8,max_iter=20000))
        print(n,m,"ok",r["status"],r["iters"])
    except Exception as e: print(n,m,e)
