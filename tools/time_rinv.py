"""R.Inv / rcond path: the tensor-path kernel (rinv_mma.cuh) against the substitution kernel it replaces (CCGP_RINV_OLD=1):
agreement and wall clock through the host API.  usage: python tools/time_rinv.py"""
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import ccgp_b200  # noqa: E402
from ccgp_b200 import workloads, GAUSS_ISO, GAUSS_ANISO_LAMBDA, LOGSCALE  # noqa: E402

eng = ccgp_b200.Engine(0)
rng = np.random.default_rng(11)


def best(fn, reps=3):
    fn()
    ts = []
    for _ in range(reps):
        t0 = time.perf_counter(); r = fn(); ts.append(time.perf_counter() - t0)
    return min(ts), r


for tag in ("m1", 14, 50, 64, 90, 97):
    if tag == "m1":
        X, y, s2 = workloads.m1_design(); fam = GAUSS_ANISO_LAMBDA; scale = LOGSCALE
        B = 16384; th = workloads.m1_candidates(B)
    else:
        n = tag; d = {14: 2, 50: 9, 64: 4, 90: 9, 97: 3}[n]
        X = rng.uniform(-1, 1, (n, d)); y = rng.normal(size=n); fam = GAUSS_ISO if d > 3 else GAUSS_ANISO_LAMBDA; scale = 0
        B = 16384
        t0 = 8.0 / d * n ** (1.0 / d)
        k = 3 if fam == GAUSS_ISO else d + 2
        th = np.column_stack([rng.uniform(0.2, 0.8, B)] + [rng.uniform(0.5 * t0, 1.5 * t0, B) for _ in range(k - 2)] + [rng.uniform(0.5, 3.0, B)])
    eng.set_design(X, y)
    n = X.shape[0]
    out = {}
    for name, env in (("old", "1"), ("mma", "0")):
        os.environ["CCGP_RINV_OLD"] = env
        t_rc, (rc, beta, st) = best(lambda: eng.rcond_batch(th, fam, scale=scale))
        Bi = 2048
        t_ri, (ri, b2, st2) = best(lambda: eng.rinv_batch(th[:Bi], fam, scale=scale))
        out[name] = (t_rc, rc, beta, st, t_ri, ri)
    o, m = out["old"], out["mma"]
    ok = (o[3] == 0)
    drc = np.max(np.abs(o[1][ok] - m[1][ok]) / o[1][ok])
    dbeta = np.max(np.abs(o[2][ok] - m[2][ok]) / np.maximum(1.0, np.abs(o[2][ok])))
    scale_ri = np.abs(o[5]).reshape(o[5].shape[0], -1).max(axis=1)
    dri = np.nanmax(np.abs(o[5] - m[5]).reshape(o[5].shape[0], -1).max(axis=1) / scale_ri)
    kap = 1.0 / np.minimum(o[1][:2048][ok[:2048]], 1.0)
    sym = np.nanmax(np.abs(m[5] - np.swapaxes(m[5], -1, -2)))
    print("n=%3d d=%d B=%d: rcond old %.2f ms (%.2f M/s) mma %.2f ms (%.2f M/s)  |  R.Inv x%d old %.2f ms mma %.2f ms  |  same status %s  max rel diff rcond %.1e beta %.1e R.Inv %.1e (max kappa_1 %.1e)  asym %.1e" % (
        n, X.shape[1], B, o[0] * 1e3, B / o[0] / 1e6, m[0] * 1e3, B / m[0] / 1e6, 2048, o[4] * 1e3, m[4] * 1e3,
        np.array_equal(o[3], m[3]), drc, dbeta, dri, kap.max(), sym), flush=True)
