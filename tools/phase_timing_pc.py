"""Per-phase clock breakdown of the producer/consumer packed kernel (factor_pc.cuh) on the M1 workload (debug aid).
Counters: lane 0 of consumer warp 0 and of producer warp 8 (same sub-partition; the producer also serves consumer 4) of block 0.
usage: python tools/phase_timing_pc.py [split ...]"""
import ctypes as C
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import ccgp_b200  # noqa: E402
from ccgp_b200 import workloads, GAUSS_ANISO_LAMBDA, LOGSCALE  # noqa: E402

os.environ["CCGP_KERNEL"] = "6"
eng = ccgp_b200.Engine(0)
X, y, s2 = workloads.m1_design()
eng.set_design(X, y)
B = 1 << 16
th = workloads.m1_candidates(B)
for split in [int(a) for a in sys.argv[1:]] or [0, 1, 2]:
    os.environ["CCGP_PC_SPLIT"] = str(split)
    eng.nll_batch(th, GAUSS_ANISO_LAMBDA, s2, scale=LOGSCALE)
    buf = (C.c_longlong * 32)()
    eng._lib.ccgp_debug_phase_timing(eng._h, 1, None)
    eng.nll_batch(th, GAUSS_ANISO_LAMBDA, s2, scale=LOGSCALE)
    eng._lib.ccgp_debug_phase_timing(eng._h, 0, C.cast(buf, C.c_void_p))
    print("split", split, "variant", eng.last_nll_config())
    ncand = max(buf[7], 1)
    names = {0: "C: previous store -> step start", 1: "C: wait FULL", 2: "C: own tiles", 3: "C: load + panels", 4: "C: 8x8 Cholesky/inverse",
             5: "C: solve + store", 6: "C: scalars + output", 8: "P: loop overhead", 9: "P: wait LOADED (consumer 0)",
             10: "P: wait LOADED (consumer 4)", 11: "P: parameter transform (both)", 12: "P: assemble (both)"}
    for ph in sorted(names):
        print("  %-36s %9.0f clk per candidate" % (names[ph], buf[ph] / ncand))
    print("  consumer total %.0f, producer total %.0f clk per candidate of consumer 0 (%d candidates)" % (
        sum(buf[0:7]) / ncand, sum(buf[8:13]) / ncand, ncand))
