"""Tiny end-to-end pass over every kernel (for compute-sanitizer runs)."""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import ccgp_b200
from ccgp_b200 import GAUSS_ISO, GAUSS_ANISO_LAMBDA, MEAN_ZERO_PLUS_TAU2
eng = ccgp_b200.Engine(0)
rng = np.random.default_rng(0)
for n, d, fam in ((20, 2, GAUSS_ANISO_LAMBDA), (70, 3, GAUSS_ISO), (100, 2, GAUSS_ANISO_LAMBDA)):
    X = rng.uniform(-1, 1, (n, d)); y = rng.normal(size=n)
    eng.set_design(X, y)
    k = eng.num_params(fam)
    cand = np.column_stack([rng.uniform(0.2, 0.8, 24)] + [rng.uniform(20, 40, 24) for _ in range(k - 1)])
    nll, beta, st = eng.nll_batch(cand, fam, 1.0)
    nll2, _, _ = eng.nll_batch(cand, fam, 1.0, mean_mode=MEAN_ZERO_PLUS_TAU2, tau=10.0)
    m, v, _ = eng.predict(cand[:3], fam, rng.uniform(-1, 1, (9, d)), 1.0)
    ri, _, _ = eng.rinv_batch(cand[:2], fam)
    print(n, d, np.isfinite(nll).all(), np.isfinite(nll2).all(), np.isfinite(m).all(), np.isfinite(ri).all())
D_old = rng.uniform(-1, 1, (14, 2)); pool = rng.uniform(-1, 1, (40, 7, 2))
nd, ld, st = eng.me_schur_batch(D_old, pool, [[0.5, 1.0, 4.0], [0.3, 2.0, 3.0]])
os.environ["CCGP_ME_GENERIC"] = "1"
nd2, _, _ = eng.me_schur_batch(D_old, pool, [[0.5, 1.0, 4.0], [0.3, 2.0, 3.0]])
print("me", np.abs(nd - nd2).max())
out, st = eng.subset_logdet_batch(rng.uniform(-1, 1, (200, 2)), rng.integers(0, 200, (10, 21)).astype(np.int32), GAUSS_ISO, [0.5, 30.0, 60.0])
os.environ["CCGP_FORCE_BIG"] = "1"
X = rng.uniform(-1, 1, (70, 2)); y = rng.normal(size=70)
eng.set_design(X, y)
nll, _, _ = eng.nll_batch(np.column_stack([rng.uniform(0.2, 0.8, 3), rng.uniform(20, 40, 3), rng.uniform(20, 40, 3)]), GAUSS_ISO, 1.0)
print("big", nll)
eng.close()
