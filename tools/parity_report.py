"""Observed parity errors of the CUDA path on the golden cases, with the kappa_1(R) of the worst row
(every gate looser than 1e-10 must state the observed error and the row's kappa).
Run on the GPU box:  python tools/parity_report.py > profiles/r02_parity_errors.txt
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import ccgp_b200  # noqa: E402
from ccgp_b200 import GAUSS_ISO, GAUSS_ANISO_LAMBDA, MEAN_ZERO_PLUS_TAU2  # noqa: E402
from ccgp_b200 import reference_api as api  # noqa: E402

G = dict(np.load(os.path.join(ROOT, "tests", "golden", "golden_cases.npz")))
G2 = dict(np.load(os.path.join(ROOT, "tests", "golden", "golden_r02.npz")))
D = dict(np.load(os.path.join(ROOT, "convex-combination-of-gaussian-processes_b200", "data", "reference_designs.npz")))
GV = dict(np.load(os.path.join(ROOT, "tests", "golden", "gv_sets.npz")))


def rel(a, b):
    return np.abs(np.asarray(a) - np.asarray(b)) / np.maximum(np.abs(b), 1.0)


def worst(name, err, kappa=None, note=""):
    err = np.asarray(err).reshape(-1)
    i = int(np.nanargmax(err))
    k = "" if kappa is None else "  kappa_1 of that row %.2e (max over rows %.2e)" % (np.asarray(kappa).reshape(-1)[i % np.size(kappa)], np.max(kappa))
    print("%-58s max %.2e%s %s" % (name, err[i], k, note))


def main():
    eng = ccgp_b200.Engine(0)
    print("# observed |gpu - oracle| on the golden cases (relative = |a-b|/max(|b|,1) unless stated); reference-faithful oracle")
    # GLS NLL / beta per configuration
    cases = [("c1n100", D["maximin100"], G["c1n100_y"], G["c1n100_nat"], GAUSS_ANISO_LAMBDA, 1.0),
             ("c1n14gls", D["maximin14"], G["c1n14_y"], G["c1n14_nat"], GAUSS_ISO, 0.7),
             ("c2gls", D["he_train"][:, :4], D["he_train"][:, 4], G["c2_nat"], GAUSS_ISO, 30.0),
             ("gv50", D["gv50_train1"][:, :9], D["gv50_train1"][:, 9], G["gv50_nat"], GAUSS_ISO, 13.0),
             ("gv90", D["gv90_train1"][:, :9], D["gv90_train1"][:, 9], G["gv90_nat"], GAUSS_ISO, 13.0)]
    for tag, X, y, nat, fam, s2 in cases:
        eng.set_design(X, y)
        nll, beta, st = eng.nll_batch(nat, fam, s2)
        kap = G[tag + "_kappa"]
        ok = (kap <= 1e6) & (st == 0)
        worst("%s NLL  (rows kappa<=1e6: %d of %d)" % (tag, ok.sum(), len(kap)), rel(-nll[ok], G[tag + "_ref"][ok]), kap[ok])
        worst("%s beta" % tag, rel(beta[ok], G[tag + "_beta"][ok]), kap[ok])
        hi = (kap > 1e6) & (st == 0)
        if hi.any():
            worst("%s NLL  (rows kappa>1e6: %d)" % (tag, hi.sum()), rel(-nll[hi], G[tag + "_ref"][hi]), kap[hi])
            worst("%s beta (rows kappa>1e6)" % tag, rel(beta[hi], G[tag + "_beta"][hi]), kap[hi])
    # tau^2 variant vs minimal / truth
    for tag, X, y, nat, s2, tau in (("c1n14", D["maximin14"], G["c1n14_y"], G["c1n14_nat"], 0.7, 100.0),
                                    ("c2tau", D["he_train"][:, :4], D["he_train"][:, 4], G["c2_nat"], 30.0, 50.0)):
        eng.set_design(X, y)
        nll, _, st = eng.nll_batch(nat, GAUSS_ISO, s2, mean_mode=MEAN_ZERO_PLUS_TAU2, tau=tau)
        worst("%s tau^2 NLL vs minimal oracle" % tag, rel(-nll, G[tag + "_minimal"]), G[tag + "_kappa"])
        worst("%s tau^2 NLL vs reference-faithful oracle" % tag, rel(-nll, G[tag + "_ref"]), G[tag + "_kappa"], "(the reference's own error)")
        okt = ~np.isnan(G[tag + "_truth"])
        worst("%s tau^2 NLL vs 50-digit truth" % tag, rel(-nll[okt], G[tag + "_truth"][okt]), G[tag + "_kappa"][okt])
    # likeli.hyperpars
    X = D["maximin14"]
    hp = D["hyperpars_2d"]
    N = int(G["c1n14_likeli_N"])
    errs = []
    for i, want in zip(G["c1n14_likeli_rows"].astype(int), G["c1n14_likeli"]):
        got = api.likeli_hyperpars(X, G["c1n14_y"], hp[i, 0:2], hp[i, 2:4], 0.7, N=N, tau=100.0, engine=eng)
        errs.append(abs(np.log(got) - np.log(want)))
    worst("likeli.hyperpars log-value vs reference-faithful oracle", errs, None, "(absolute in log)")
    # predictive tables
    eng.set_design(D["maximin14"], G["pred14_y"])
    m, v, _ = eng.predict(G["pred14_pars"], GAUSS_ANISO_LAMBDA, G["pred14_Xnew"], 0.9)
    worst("pred14 mean", rel(m, G["pred14_mean"]))
    worst("pred14 var / sigma2 (absolute)", np.abs(v - G["pred14_var"]) / 0.9)
    he, het = D["he_train"], D["he_test"]
    eng.set_design(he[:, :4], he[:, 4])
    m, v, _ = eng.predict(G["predHE_pars"], GAUSS_ISO, het[:, :4], 30.0)
    kap = G["c2gls_kappa"][:4]
    worst("predHE mean (n=64, d=4)", rel(m, G["predHE_mean"]).max(axis=0), kap)
    worst("predHE var / sigma2 (absolute)", (np.abs(v - G["predHE_var"]) / 30.0).max(axis=0), kap)
    tr, te = D["gv50_train1"], D["gv50_test1"]
    eng.set_design(tr[:, :9], tr[:, 9])
    m, v, _ = eng.predict(G["predGV_pars"], GAUSS_ISO, te[:20, :9], 13.0)
    kap = G["gv50_kappa"][:3]
    worst("predGV mean (n=50, d=9)", rel(m, G["predGV_mean"]).max(axis=0), kap)
    worst("predGV var / sigma2 (absolute)", (np.abs(v - G["predGV_var"]) / 13.0).max(axis=0), kap)
    for size, count in ((50, 9), (90, 8)):
        em, ev, en, eb, kk = [], [], [], [], []
        for i in range(1, count + 1):
            trn, tst = GV["train%d_%d" % (size, i)], GV["test%d_%d" % (size, i)]
            tag = "gv%d_%d_" % (size, i)
            eng.set_design(trn[:, :9], trn[:, 9])
            nll, beta, st = eng.nll_batch(G2[tag + "nat"], GAUSS_ISO, 13.0)
            m, v, _ = eng.predict(G2[tag + "nat"][:2], GAUSS_ISO, tst[:12, :9], 13.0)
            en.extend(rel(-nll, G2[tag + "ref"])); eb.extend(rel(beta, G2[tag + "beta"])); kk.extend(G2[tag + "kappa"])
            em.extend(rel(m, G2[tag + "pred_mean"]).max(axis=0)); ev.extend((np.abs(v - G2[tag + "pred_var"]) / 13.0).max(axis=0))
        kk = np.array(kk)
        worst("GV n=%d all %d sets: NLL" % (size, count), en, kk)
        worst("GV n=%d all sets: beta" % size, eb, kk)
        worst("GV n=%d all sets: predictive mean" % size, em, np.array(kk).reshape(count, 8)[:, :2].reshape(-1))
        worst("GV n=%d all sets: predictive var / sigma2" % size, ev, np.array(kk).reshape(count, 8)[:, :2].reshape(-1))
    # ME
    D_old, pool = D["me_initial14"], D["me_all_subdesigns"]
    nd, ld, st = eng.me_schur_batch(D_old, pool[:200], G["me_params"])
    worst("ME -det, 200 designs x 6 rows (relative to each value)", np.abs(nd - G["me_negdet_200"]) / np.abs(G["me_negdet_200"]))
    print("ME values span %.2e .. %.2e; relative error of a determinant of a 7x7 Schur complement ~ 7 * kappa(S) * eps" % (
        np.abs(G["me_negdet_200"]).min(), np.abs(G["me_negdet_200"]).max()))
    eng.close()


if __name__ == "__main__":
    main()
