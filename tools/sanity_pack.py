"""Packed kernel (factor_pack.cuh) against the shipped kernel choice on small batches of synthetic designs.
usage: python tools/sanity_pack.py [B] [n,d,logscale ...]      e.g.  python tools/sanity_pack.py 4096 100,2,1 90,2,0"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import ccgp_b200  # noqa: E402
from ccgp_b200 import workloads, GAUSS_ANISO_LAMBDA, LOGSCALE  # noqa: E402

B = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
cases = [tuple(int(v) for v in a.split(",")) for a in sys.argv[2:]] or [(100, 2, 1)]
eng = ccgp_b200.Engine(0)
rng = np.random.default_rng(11)
for n, d, logs in cases:
    if (n, d) == (100, 2):
        X, y, s2 = workloads.m1_design()
    else:
        X = rng.uniform(-1, 1, (n, d)); y = rng.normal(size=n); s2 = 1.0
    eng.set_design(X, y)
    if logs:
        th = np.column_stack([rng.normal(np.log(20.0 / d), 0.5, B) for _ in range(d)] + [rng.normal(0, 1, B), rng.normal(0, 1, B)])
        scale = LOGSCALE
    else:
        t0 = 8.0 / d * n ** (1.0 / d)
        th = np.column_stack([rng.uniform(0.2, 0.8, B)] + [rng.uniform(0.5 * t0, 1.5 * t0, B) for _ in range(d)] + [rng.uniform(0.5, 3.0, B)])
        scale = 0
    os.environ.pop("CCGP_KERNEL", None)
    ref, rb, rs = eng.nll_batch(th, GAUSS_ANISO_LAMBDA, s2, scale=scale)
    cfg0 = eng.last_nll_config()["variant"]
    for kern in os.environ.get("SANITY_KERNELS", "5").split(","):
      os.environ["CCGP_KERNEL"] = kern
      for rep in range(2):
          v, b, s = eng.nll_batch(th, GAUSS_ANISO_LAMBDA, s2, scale=scale)
          ok = np.isfinite(ref) & np.isfinite(v)
          nbad = int(np.sum(~(np.abs(v - ref) <= 1e-9 * np.abs(ref)) & ~(np.isnan(v) & np.isnan(ref))))
          print("n=%d d=%d log=%d ref variant %s kernel %s: maxrel nll %.2e beta %.2e, NaN sets equal %s, wrong %d of %d (finite %d)" % (
              n, d, logs, cfg0, eng.last_nll_config(), np.max(np.abs(v[ok] - ref[ok]) / np.maximum(1, np.abs(ref[ok]))),
              np.max(np.abs(b[ok] - rb[ok]) / np.maximum(1, np.abs(rb[ok]))), np.array_equal(np.isfinite(ref), np.isfinite(v)), nbad, B, ok.sum()))
