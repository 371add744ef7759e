"""Per-phase clock breakdown of the DMMA factor kernel (factor_mma.cuh) on the M1 workload (debug aid).
Counters are read from block 0, lane 0 of warp 0 (diagonal warp) and warp 1 (first update warp)."""
import ctypes as C
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import ccgp_b200  # noqa: E402
from ccgp_b200 import workloads, GAUSS_ANISO_LAMBDA, LOGSCALE  # noqa: E402

os.environ.setdefault("CCGP_KERNEL", "4")   # the CTA-per-candidate DMMA kernel (factor_mma.cuh)
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
cfg = eng.last_nll_config()
ncand = B / (148 * cfg["ctas_per_sm"])
names = {0: ["build", "build barrier", "diag tile: last panel (2 DMMA)", "8x8 Cholesky + inverse + publish",
             "step barrier", "-", "scalars"],
         1: ["build", "build barrier", "issue (A) last panel + (B) lookahead", "wait for the diagonal warp",
             "solve (2 DMMA/tile) + store", "step barrier", "scalars"]}
print("variant", cfg, "candidates per CTA ~ %.1f" % ncand)
for w in (0, 1):
    tot = 0
    for ph, nm in enumerate(names[w]):
        v = buf[w * 16 + ph]
        tot += v
        print("warp %d %-40s %9.0f clk per candidate" % (w, nm, v / ncand))
    print("warp %d total %.0f clk per candidate" % (w, tot / ncand))
