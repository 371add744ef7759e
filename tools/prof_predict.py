"""One predictive table (n=100, S=1000 posterior rows x T=625 sites) for ncu: python tools/prof_predict.py"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import ccgp_b200  # noqa: E402
from ccgp_b200 import GAUSS_ANISO_LAMBDA  # noqa: E402

eng = ccgp_b200.Engine(0)
rng = np.random.default_rng(3)
n, d, S, T = 100, 2, 1000, 625
X = rng.uniform(-1, 1, (n, d)); y = rng.normal(size=n)
eng.set_design(X, y)
pars = np.column_stack([rng.uniform(0.2, 0.8, S)] + [rng.uniform(20, 60, S) for _ in range(2)] + [rng.uniform(0.5, 2, S)])
Xn = rng.uniform(-1, 1, (T, d))
for _ in range(2):
    m, v, st = eng.predict(pars, GAUSS_ANISO_LAMBDA, Xn, 1.0)
print(float(m.mean()), int(st.sum()))
