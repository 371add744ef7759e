"""Lock-step projected L-BFGS (me_design.all_subdesigns) vs scipy L-BFGS-B run one start at a time, same starts."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, ccgp_b200, time
from ccgp_b200 import me_design
from scipy.optimize import minimize
z=np.load("convex-combination-of-gaussian-processes_b200/data/reference_designs.npz"); D_old=z["me_initial14"]
e=ccgp_b200.Engine(0); rng=np.random.default_rng(11)
P=8; params=np.column_stack([rng.uniform(0.3,0.7,P), rng.uniform(0.5,2.0,P), rng.uniform(3.0,6.0,P)])
ns=10
starts=me_design.random_lhd_starts(rng, P*ns, 7, 2)
for maxit in (100, 300):
    out=me_design.all_subdesigns(D_old, params, 7, 2, ns, rng, e, starts=starts, maxit=maxit)
    print("lockstep maxit", maxit, "iters", out["iterations"], np.round(out["values"],5))
    allv=out["all_values"]
sc=np.zeros((P,ns)); nit=np.zeros((P,ns))
for q in range(P):
    for s in range(ns):
        def f_and_g(x):
            pts=np.repeat(x[None,:],29,axis=0)
            for i in range(14):
                pts[1+2*i,i]+=1e-3; pts[2+2*i,i]-=1e-3
            nd=e.me_schur_batch(D_old, pts.reshape(-1,2,7).transpose(0,2,1), params[q:q+1])[0][:,0]
            return float(nd[0]), (nd[1::2]-nd[2::2])/2e-3
        r=minimize(f_and_g, starts[q*ns+s], jac=True, method="L-BFGS-B", bounds=[(-1.0,1.0)]*14, options=dict(maxiter=100))
        sc[q,s]=r.fun; nit[q,s]=r.nit
print("scipy best      ", np.round(sc.min(1),5), "mean nit", nit.mean())
print("per-start: lockstep better or equal (1e-4 rel) fraction", np.mean(allv <= sc*(1-1e-4)+0), " mean lock", allv.mean(), "mean scipy", sc.mean())
