"""One M1 launch of the NLL kernel (for ncu): python tools/prof_nll.py [batch]"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import ccgp_b200  # noqa: E402
from ccgp_b200 import workloads, GAUSS_ANISO_LAMBDA, LOGSCALE  # noqa: E402

B = int(sys.argv[1]) if len(sys.argv) > 1 else 1 << 17
eng = ccgp_b200.Engine(0)
X, y, s2 = workloads.m1_design()
eng.set_design(X, y)
th = workloads.m1_candidates(B)
for _ in range(2):
    nll, beta, st = eng.nll_batch(th, GAUSS_ANISO_LAMBDA, s2, scale=LOGSCALE)
print(eng.last_nll_config(), float(nll.min()))
