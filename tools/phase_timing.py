"""Per-phase clock breakdown of the factor kernel on the M1 workload (debug aid).
Counters are read from block 0, lane 0 of warps 0 and 1."""
import ctypes as C
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import ccgp_b200  # noqa: E402
from ccgp_b200 import workloads, GAUSS_ANISO_LAMBDA, LOGSCALE  # noqa: E402

os.environ.setdefault("CCGP_NO_MMA", "1")   # the DFMA kernel (factor_engine.cuh)
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
names = ["build", "U1", "lookahead(tiles)+wait", "trsm", "diag block (warp 0)"]
print("variant", cfg, "candidates per CTA ~ %.1f" % ncand)
for w in (0, 1):
    tot = 0
    for ph, nm in enumerate(names):
        work, wait = buf[w * 16 + ph * 2], buf[w * 16 + ph * 2 + 1]
        tot += work + wait
        print("warp %d %-14s work %9.0f  barrier %9.0f   (clk per candidate)" % (w, nm, work / ncand, wait / ncand))
    print("warp %d total %.0f clk per candidate" % (w, tot / ncand))
