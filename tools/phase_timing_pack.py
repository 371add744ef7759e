"""Per-phase clock breakdown of the packed-residency kernel (factor_pack.cuh) on the M1 workload (debug aid).
Counters come from warp 0 of block 0 (it shares its sub-partition with the CTA's warps 4, 8, ..)."""
import ctypes as C
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import ccgp_b200  # noqa: E402
from ccgp_b200 import workloads, GAUSS_ANISO_LAMBDA, LOGSCALE  # noqa: E402

os.environ.setdefault("CCGP_KERNEL", "5")
eng = ccgp_b200.Engine(0)
X, y, s2 = workloads.m1_design()
eng.set_design(X, y)
B = 1 << 16
th = workloads.m1_candidates(B)
eng.nll_batch(th, GAUSS_ANISO_LAMBDA, s2, scale=LOGSCALE)
buf = (C.c_longlong * 32)()
eng._lib.ccgp_debug_phase_timing(eng._h, 1, None)
eng.nll_batch(th, GAUSS_ANISO_LAMBDA, s2, scale=LOGSCALE)
eng._lib.ccgp_debug_phase_timing(eng._h, 0, C.cast(buf, C.c_void_p))
print("variant", eng.last_nll_config())
ncand = max(buf[6], 1)
names = ["parameters + staging", "assemble the column", "panels + 8x8 Cholesky/inverse + solve", "scalars + output"]
tot = 0
for ph, nm in enumerate(names):
    tot += buf[ph]
    print("%-40s %9.0f clk per candidate" % (nm, buf[ph] / ncand))
print("total %.0f clk per candidate (%d candidates)" % (tot / ncand, ncand))
