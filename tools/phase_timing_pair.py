"""Per-phase clock breakdown of the two-warp DMMA kernel (factor_pair_kernel) on the M1 workload (debug aid).
Counters come from team 0 of block 0: warp A (diagonal chain) and warp B (updates)."""
import ctypes as C
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import ccgp_b200  # noqa: E402
from ccgp_b200 import workloads, GAUSS_ANISO_LAMBDA, LOGSCALE  # noqa: E402

os.environ.setdefault("CCGP_KERNEL", "3"); os.environ.setdefault("CCGP_TEAM_NW", "3")
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
ncand = max(buf[15], 1)
names = {0: ["top barrier", "build", "build barrier", "8x8 Cholesky + inverse", "barrier with B (waits for the lookahead)",
             "own solve of tile (c+1,c) + diagonal tile update", "-", "final barrier + scalars + output"],
         1: ["top barrier", "build", "build barrier", "last panel + lookahead (DMMA issue)", "barrier with A (waits for the chain)",
             "solve + store", "B-only barrier", "final barrier + next parameters"]}
for w in (0, 1):
    tot = 0
    for ph, nm in enumerate(names[w]):
        v = buf[w * 16 + ph]
        tot += v
        print("warp %s %-38s %9.0f clk per candidate" % ("AB"[w], nm, v / ncand))
    print("warp %s total %.0f clk per candidate (%d candidates)" % ("AB"[w], tot / ncand, ncand))
