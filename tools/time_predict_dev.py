"""Device time (CUDA events) of the predictive table next to the factor-only time of the same posterior rows and the
host-pointer wall clock, at the shapes BASELINE.json's configurations use (HE: n=64 d=4 T=14; GV: n=50/90 d=9 T=150/110;
M1 grid: n=100 d=2 T=625).  Usage: python tools/time_predict_dev.py"""
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import ccgp_b200  # noqa: E402
from ccgp_b200 import GAUSS_ISO, GAUSS_ANISO_LAMBDA  # noqa: E402

eng = ccgp_b200.Engine(0)
eng.set_stream(torch.cuda.current_stream().cuda_stream)
rng = np.random.default_rng(3)
dev = torch.device("cuda:0")


def ev_time(fn, reps=20):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


shapes = ((64, 4, GAUSS_ISO, 1000, 14), (50, 9, GAUSS_ISO, 1000, 150), (90, 9, GAUSS_ISO, 1000, 110),
          (14, 2, GAUSS_ANISO_LAMBDA, 1000, 625), (100, 2, GAUSS_ANISO_LAMBDA, 1000, 625), (100, 2, GAUSS_ANISO_LAMBDA, 1000, 16),
          (64, 4, GAUSS_ISO, 8000, 14), (50, 9, GAUSS_ISO, 8000, 150), (100, 2, GAUSS_ANISO_LAMBDA, 16, 10000))
for n, d, fam, S, T in shapes:
    X = rng.uniform(-1, 1, (n, d)); y = rng.normal(size=n)
    eng.set_design(X, y)
    k = eng.num_params(fam)
    th = 8.0 / d * n ** (1.0 / d)
    pars = np.column_stack([rng.uniform(0.2, 0.8, S)] + [rng.uniform(0.5 * th, 1.5 * th, S) for _ in range(k - 1)])
    Xn = rng.uniform(-1, 1, (T, d))
    pars_t = torch.tensor(np.ascontiguousarray(pars.T), device=dev)
    xn_t = torch.tensor(np.ascontiguousarray(Xn.T), device=dev)
    om = torch.empty(S * T, dtype=torch.float64, device=dev)
    ov = torch.empty(S * T, dtype=torch.float64, device=dev)
    st = torch.empty(S, dtype=torch.int32, device=dev)
    t_pred = ev_time(lambda: eng.predict_dev(pars_t, fam, xn_t, 1.0, om, ov, st))
    t_nll = ev_time(lambda: eng.nll_batch_dev(pars_t, fam, 1.0))
    # factors kept on the device (ccgp_factors_*): creation once, then only the site phase per table
    t0 = time.perf_counter()
    fac = eng.factors(pars, fam)
    t_create = time.perf_counter() - t0
    om2 = torch.empty(S * T, dtype=torch.float64, device=dev)
    ov2 = torch.empty(S * T, dtype=torch.float64, device=dev)
    t_fac = ev_time(lambda: fac.predict_dev(xn_t, 1.0, om2, ov2, st))
    same = bool(torch.equal(om, om2) and torch.equal(ov, ov2))
    finfo = fac.info()
    eng.predict(pars[:8], fam, Xn[:8], 1.0)
    ts = []
    for _ in range(5):
        t0 = time.perf_counter()
        m, v, s_ = eng.predict(pars, fam, Xn, 1.0)
        ts.append(time.perf_counter() - t0)
    tsf = []
    for _ in range(5):
        t0 = time.perf_counter()
        keep = fac.predict(Xn, 1.0)          # (kept: a result freed at once is trimmed off the heap and page-faults back in)
        tsf.append(time.perf_counter() - t0)
    fac.close()
    mref = om.cpu().numpy().reshape(S, T)
    print("n=%3d d=%d S=%d T=%d: predict kernel %7.1f us (%7.1f M pairs/s)  factor-only (NLL kernel) %6.1f us  host API %7.1f us (%7.1f M pairs/s)  maxdiff %.1e  finite %.3f"
          % (n, d, S, T, t_pred * 1e3, S * T / t_pred / 1e3, t_nll * 1e3, min(ts) * 1e6, S * T / min(ts) / 1e6,
             float(np.nanmax(np.abs(mref - np.asarray(m).T))), np.isfinite(m).mean()), flush=True)
    print("      stored factors (%.1f MB, created in %.0f us): site phase only %7.1f us (%7.1f M pairs/s)  host API %7.1f us (%7.1f M pairs/s)  bit-identical %s"
          % (finfo["device_bytes"] / 1e6, t_create * 1e6, t_fac * 1e3, S * T / t_fac / 1e3, min(tsf) * 1e6, S * T / min(tsf) / 1e6, same), flush=True)
