"""Where does a lock-step Metropolis step spend its time?  Wraps the engine calls of one ground-vibrations fit with timers.
usage: python tools/diag_gv_step.py [set index]"""
import os
import sys
import time
import collections

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tools"))
import ccgp_b200  # noqa: E402
import fit_gv  # noqa: E402
import fit_gv_sets  # noqa: E402

idx = int(sys.argv[1]) if len(sys.argv) > 1 else 1
S = fit_gv_sets.sets()
eng = ccgp_b200.Engine(0)
fit_gv.fit(eng, 2, N=200, samp_size=100, train=S[0][1], test=S[0][2])
if os.environ.get("DIAG_FULL_FIRST", "0") == "1":
    t0 = time.perf_counter()
    fit_gv.fit(eng, 32, seed=100, train=S[0][1], test=S[0][2])
    print("full fit of set 0 first: %.2f s" % (time.perf_counter() - t0))
acc = collections.defaultdict(list)


def wrap(name):
    f = getattr(eng, name)

    def g(*a, **k):
        t0 = time.perf_counter()
        r = f(*a, **k)
        acc[name].append(time.perf_counter() - t0)
        return r
    setattr(eng, name, g)


for nm in ("nll_batch", "rcond_batch", "set_design", "predict"):
    wrap(nm)
t0 = time.perf_counter()
r = fit_gv.fit(eng, 32, seed=100 + idx, train=S[idx][1], test=S[idx][2])
tot = time.perf_counter() - t0
print("set %d (%s): %.2f s total, CCGP_RINV_OLD=%s" % (idx, S[idx][0], tot, os.environ.get("CCGP_RINV_OLD", "0")))
for k, v in acc.items():
    v = np.array(v)
    print("  %-12s calls %5d  total %7.3f s  median %8.1f us  p99 %9.1f us  max %9.1f us" % (k, len(v), v.sum(), np.median(v) * 1e6, np.percentile(v, 99) * 1e6, v.max() * 1e6))
