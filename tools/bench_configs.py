"""Throughput on every configuration BASELINE.json lists (SURVEY 8d "secondary inputs"), through the host-pointer
C ABI (copies included), on the shipped designs.  Run on the B200 box: python tools/bench_configs.py
-> one JSON object per line (also written to gpurun_out/configs.jsonl)."""
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import ccgp_b200  # noqa: E402
from ccgp_b200 import workloads, GAUSS_ISO, GAUSS_ANISO_LAMBDA, MATERN1D, LOGSCALE, MEAN_ZERO_PLUS_TAU2  # noqa: E402
from ccgp_b200 import reference_api as api  # noqa: E402

eng = ccgp_b200.Engine(0)
D = workloads.designs()
rng = np.random.default_rng(0)
out = []


def best_of(fn, reps=3):
    fn()
    ts = []
    for _ in range(reps):
        t0 = time.perf_counter()
        r = fn()
        ts.append(time.perf_counter() - t0)
    return min(ts), r


def record(name, units, unit, dt, extra=None):
    row = dict(config=name, value=units / dt, unit=unit + "/s", seconds=dt, units=units)
    row.update(extra or {})
    out.append(row)
    print(json.dumps(row), flush=True)


# configs[1]: the reference's own hyper-prior sweep ([V]:588-599): 60 x 1728 Halton candidates, n = 14, tau = 100
X14 = D["maximin14"]
y14 = workloads.test_function_4(X14)
dt, r = best_of(lambda: api.choose_hyperpars(X14, y14, D["hyperpars_2d"], 0.7, engine=eng))
record("C1 choose.hyperpars sweep, maximin 14 pts, 60 x 1728, tau=100 (incl. qigamma grid on the host)", 60 * 1728, "evals", dt,
       dict(argmax_row=[float(v) for v in r["pars"]]))
eng.set_design(X14, y14)
cand = np.vstack([api.sweep_candidates(h[0:2], h[2:4], 1728) for h in D["hyperpars_2d"]])
dt, _ = best_of(lambda: eng.nll_batch(cand, GAUSS_ISO, 0.7, mean_mode=MEAN_ZERO_PLUS_TAU2, tau=100.0))
record("C1 sweep, likelihood calls only", cand.shape[0], "evals", dt)
# configs[1]: anisotropic NLL batches on the 14- and 100-point designs
for key, B in (("maximin14", 1 << 20), ("maximin100", 1 << 20)):
    X = D[key]
    y = workloads.test_function_4(X)
    eng.set_design(X, y)
    th = workloads.m1_candidates(B)
    if key == "maximin14":
        th = th - np.array([np.log(4.0), np.log(4.0), 0.0, 0.0])          # coarser design: smaller scales keep R well conditioned
    dt, r = best_of(lambda: eng.nll_batch(th, GAUSS_ANISO_LAMBDA, 1.0, scale=LOGSCALE))
    record("C1 anisotropic logpost batch, %s" % key, B, "evals", dt, dict(not_pd=int((r[2] != 0).sum())))
# configs[2]: heat exchanger, n = 64, d = 4: 624 x 1000 sweep (tau = 50) and the S x T predictive table
he, het = D["he_train"], D["he_test"]
eng.set_design(he[:, :4], he[:, 4])
candh = np.vstack([api.sweep_candidates(h[0:2], h[2:4], 1000) for h in D["he_hyperpars"]])
dt, _ = best_of(lambda: eng.nll_batch(candh, GAUSS_ISO, 30.0, mean_mode=MEAN_ZERO_PLUS_TAU2, tau=50.0))
record("C2 heat exchanger sweep, n=64 d=4, 624 x 1000, tau=50", candh.shape[0], "evals", dt)
pars = np.column_stack([rng.uniform(0.2, 0.8, 1000), 1 / rng.gamma(3, 1.0, 1000) + 0.5, 1 / rng.gamma(5, 1 / 40.0, 1000)])
dt, _ = best_of(lambda: eng.predict(pars, GAUSS_ISO, het[:, :4], 30.0))
record("C2 heat exchanger predictive table, S=1000 x T=14", 1000 * het.shape[0], "(row,site)", dt)
# configs[3]: ground vibrations, n = 50 / 90, d = 9
for tag, B in (("gv50", 1 << 18), ("gv90", 1 << 18)):
    tr, te = D[tag + "_train1"], D[tag + "_test1"]
    eng.set_design(tr[:, :9], tr[:, 9])
    cg = np.column_stack([rng.uniform(0.05, 0.95, B), 0.06 / rng.gamma(3, 1.0, B), 1 / rng.gamma(5, 0.5, B)])
    dt, r = best_of(lambda: eng.nll_batch(cg, GAUSS_ISO, 13.0))
    record("C3 ground vibrations NLL batch, %s (n=%d, d=9)" % (tag, tr.shape[0]), B, "evals", dt, dict(not_pd=int((r[2] != 0).sum())))
    dt, _ = best_of(lambda: eng.predict(cg[:1000], GAUSS_ISO, te[:, :9], 13.0))
    record("C3 ground vibrations predictive table, %s, S=1000 x T=%d" % (tag, te.shape[0]), 1000 * te.shape[0], "(row,site)", dt)
# configs[0]: 1-D Matern(5) + Matern designs, n = 8, 201 designs
d1 = D["design1d"]
B = 1 << 16
c1 = np.column_stack([rng.uniform(0.1, 0.9, B), rng.uniform(0.05, 0.3, B), rng.uniform(0.3, 1.5, B)])
def run_1d():
    for row in d1[:16]:
        eng.set_design(row[:, None], np.sin(6 * row))
        eng.nll_batch(c1, MATERN1D, 1.0)
dt, _ = best_of(run_1d, reps=2)
record("C0 1-D Matern(nu=5)+Matern likelihood, n=8, 16 designs x 65536 candidates", 16 * B, "evals", dt)
# configs[4]: ME second batch over All_Subdesigns x posterior draws; subset log-dets on the synthetic pool
D_old, pool = workloads.me_pool()
pp = workloads.me_params(1000)
dt, _ = best_of(lambda: eng.me_argmin(D_old, pool, pp))
record("C4 ME-A: 1000 All_Subdesigns blocks x 1000 parameter rows, argmin per row", 1000 * 1000, "dets", dt)
P = workloads.synthetic_pool(2048)
for m in (7, 21, 64, 128, 216):      # the shared-memory kernels end at n ~ 220 (m = 256 of SURVEY ME-B needs the HBM path: next)
    C = 1 << 16 if m <= 64 else 1 << 13
    idx = np.asfortranarray(np.stack([rng.choice(2048, size=m, replace=False) for _ in range(C)]).astype(np.int32))   # the ABI's (R's) layout
    dt, r = best_of(lambda: eng.subset_logdet_batch(P, idx, GAUSS_ANISO_LAMBDA, [0.5, 30.0, 60.0, 2.0]))
    record("C4 ME-B: subset log-dets, m=%d of 2048 synthetic points" % m, C, "logdets", dt, dict(failed=int((r[1] != 0).sum())))
dt, r = best_of(lambda: eng.kmedoids_pam(D["me_all_subdesigns"].reshape(-1, 2), 7), reps=2)
record("C4 7-medoids of the 7000 All_Subdesigns points", 1, "clusterings", dt, dict(medoid_rows_1based=sorted(int(v) + 1 for v in r[0])))
os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
with open(os.path.join(ROOT, "gpurun_out", "configs.jsonl"), "w") as f:
    for row in out:
        f.write(json.dumps(row) + "\n")
