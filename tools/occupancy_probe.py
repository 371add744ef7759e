"""How much does residency (CTAs/SM) matter?  n = 94 fits 5 candidates per SM, n = 100 fits 4."""
import os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import ccgp_b200
from ccgp_b200 import workloads, GAUSS_ANISO_LAMBDA, LOGSCALE
eng = ccgp_b200.Engine(0)
dev = torch.device("cuda", 0)
stream = torch.cuda.current_stream(dev)
eng.set_stream(stream.cuda_stream)
X, y, s2 = workloads.m1_design()
B = 1 << 17
cand = torch.from_numpy(np.asfortranarray(workloads.m1_candidates(B)).T.copy()).to(dev)
for n in (100,):
    eng.set_design(X[:n], y[:n])
    for cap in (0,):
        if cap: os.environ["CCGP_CTAS_PER_SM"] = str(cap)
        else: os.environ.pop("CCGP_CTAS_PER_SM", None)
        out = eng.nll_batch_dev(cand, GAUSS_ANISO_LAMBDA, s2, scale=LOGSCALE)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        eng.nll_batch_dev(cand, GAUSS_ANISO_LAMBDA, s2, scale=LOGSCALE, out_nll=out[0], out_beta=out[1], out_status=out[2])
        e1.record(stream); torch.cuda.synchronize()
        ms = e0.elapsed_time(e1)
        print("n=%3d cap=%d -> %s  %.3f ms  %.2f M evals/s" % (n, cap, eng.last_nll_config(), ms, B / ms / 1e3))
