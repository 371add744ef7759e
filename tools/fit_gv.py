"""End-to-end fit of the ground-vibrations emulator ([G]:689-762) with the lock-step drivers: Laplace start, C Metropolis
chains in lock-step (one batched logpost per step), predictive table of the pooled posterior sample at the 150 test
sites -- compared with the reference's own stored result for this training set (`Size 50 Results 1.txt`).
usage: python tools/fit_gv.py [chains]"""
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import ccgp_b200  # noqa: E402
from ccgp_b200 import reference_api as api, samplers, workloads  # noqa: E402


def fit(eng, C, seed=1, N=5000, samp_size=1000, batch_size=20, alpha=0.5, train=None, test=None):
    D = workloads.designs()
    tr, te = (D["gv50_train1"], D["gv50_test1"]) if train is None else (train, test)
    X, y = tr[:, :9], tr[:, 9]
    sigma2 = float(np.var(y, ddof=1))                       # stand-in for mlegp(D.train, y.train)$sig2 ([G]:720-721)
    fn = lambda th: api.logpost_batch(X, th, y, sigma2, script="G", engine=eng)   # noqa: E731
    rng = np.random.default_rng(seed)
    t0 = time.perf_counter()
    lap = samplers.laplace_batch(fn, np.array([[1.0, 1.0, 0.0]]))                  # start <- c(1,1,0) ([G]:691)
    t1 = time.perf_counter()
    chains = samplers.Metro_multichain(np.repeat(lap["mode"], C, 0), lap["var"][0], N, samp_size, batch_size, alpha, fn, rng)
    t2 = time.perf_counter()
    th = np.vstack([c["sample"] for c in chains])
    nat = api.transform(th, ccgp_b200.GAUSS_ISO, 9)
    mean, var = api.predict_post_batch(te[:, :9], X, y, nat, sigma2, script="G", engine=eng)
    t3 = time.perf_counter()
    yhat = mean.mean(axis=1)                                 # `prediction`: mean of the per-sample predictive means ([G]:~620)
    props = sum(c["n_proposals"] for c in chains)
    return dict(yhat=yhat, y_true=te[:, 9], seconds=dict(laplace=t1 - t0, metro=t2 - t1, predict=t3 - t2),
                laplace_evals=lap["evals"], proposals=props, accepted=sum(c["n_accept"] for c in chains),
                samples=th.shape[0], mode=lap["mode"][0])


if __name__ == "__main__":
    C = int(sys.argv[1]) if len(sys.argv) > 1 else 64
    eng = ccgp_b200.Engine(0)
    fit(eng, 2, N=200, samp_size=100)                        # warm-up
    r = fit(eng, C)
    ref = np.load(os.path.join(ROOT, "tests", "golden", "gv50_results1.npz"))
    rm = lambda a, b: float(np.sqrt(np.mean((a - b) ** 2)))  # noqa: E731
    print("chains %d: laplace %.2f s (%d evals), Metro %.2f s (%d proposals, %d accepted, %.0f logpost evals/s), predict %.3f s (%d rows x 150 sites)" % (
        C, r["seconds"]["laplace"], r["laplace_evals"], r["seconds"]["metro"], r["proposals"], r["accepted"],
        r["proposals"] / r["seconds"]["metro"], r["seconds"]["predict"], r["samples"]))
    print("RMSPE vs y.true: ours %.4f, reference's stored run %.4f; RMS(ours - stored y.hat.Combined) %.4f; corr %.4f; sd(y.true) %.3f" % (
        rm(r["yhat"], r["y_true"]), rm(ref["y_hat_combined"], ref["y_true"]), rm(r["yhat"], ref["y_hat_combined"]),
        float(np.corrcoef(r["yhat"], ref["y_hat_combined"])[0, 1]), float(np.std(r["y_true"]))))
