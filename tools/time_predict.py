"""Wall-clock of the predictive table (S posterior rows x T test sites) through the host API, several design sizes.
FLOP model: per (row, site) n(3d+4) + 2 n^2 + 6n, + n^3/3 per row."""
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import ccgp_b200  # noqa: E402
from ccgp_b200 import GAUSS_ISO, GAUSS_ANISO_LAMBDA  # noqa: E402

eng = ccgp_b200.Engine(0)
rng = np.random.default_rng(3)
for n, d, fam, S, T in ((14, 2, GAUSS_ANISO_LAMBDA, 1000, 625), (100, 2, GAUSS_ANISO_LAMBDA, 1000, 625), (64, 4, GAUSS_ISO, 1000, 14),
                        (50, 9, GAUSS_ISO, 1000, 150), (90, 9, GAUSS_ISO, 1000, 110), (100, 2, GAUSS_ANISO_LAMBDA, 4000, 2500)):
    X = rng.uniform(-1, 1, (n, d)); y = rng.normal(size=n)
    eng.set_design(X, y)
    k = eng.num_params(fam)
    th = 8.0 / d * n ** (1.0 / d)
    pars = np.column_stack([rng.uniform(0.2, 0.8, S)] + [rng.uniform(0.5 * th, 1.5 * th, S) for _ in range(k - 1)])
    Xn = rng.uniform(-1, 1, (T, d))
    eng.predict(pars[:8], fam, Xn[:8], 1.0)
    ts = []
    for _ in range(3):
        t0 = time.perf_counter()
        m, v, st = eng.predict(pars, fam, Xn, 1.0)
        ts.append(time.perf_counter() - t0)
    dt = min(ts)
    flop = S * T * (n * (3 * d + 4) + 2.0 * n * n + 6 * n) + S * n ** 3 / 3.0
    print("n=%3d d=%d S=%d T=%d: %8.2f ms  %7.2f M (row,site)/s  %6.3f TFLOP/s algorithmic  finite %.3f" % (
        n, d, S, T, dt * 1e3, S * T / dt / 1e6, flop / dt / 1e12, np.isfinite(m).mean()))
