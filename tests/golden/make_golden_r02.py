"""Round-2 fixtures (run in the BUILD container, where /root/reference exists):

  OPENBLAS_NUM_THREADS=1 python tests/golden/make_golden_r02.py

1. gv_sets.npz -- all 17 Ground-Vibrations training/test pairs ([G]:710-718: 9 of size 50, 8 of size 90), the
   inputs of BASELINE configs[3]; reference_designs.npz only carries sample 1 of each size.
2. golden_r02.npz -- oracle outputs for the configurations VERDICT r01 lists as untested: every GV set (NLL rows
   + a predictive table, n = 50 and n = 90), the heat-exchanger `choose.hyperpars` sweep on a subset of the 624
   hyper-prior rows with the script's own N = 1000, tau = 50 ([H]:549-595), the ME criterion over the full
   1000-design pool x 64 parameter rows ([M]:869-877, 944-945), subset log-dets of size 128 and 256,
   `solve(R)` / beta.MLE on several rows ([A]:448-452).
Like golden_cases.npz these pin the ORACLE (no R in the image): parity w.r.t. real R stays unpinned until
tests/test_r_crosscheck.py has run on a machine with R.
"""
import os
import sys
from multiprocessing import Pool

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
from oracle import ccgp_oracle as orc  # noqa: E402
from make_golden import rtable  # noqa: E402

OUT = os.path.dirname(os.path.abspath(__file__))
HE_ROWS = [0, 77, 155, 311, 468, 623]


def gv_sets():
    gv = "Ground Vibrations Emulator"
    d = {}
    for size, count in ((50, 9), (90, 8)):
        for i in range(1, count + 1):
            d["train%d_%d" % (size, i)] = rtable("%s/Training Sets/Training Set Size %d Sample %d.txt" % (gv, size, i))
            d["test%d_%d" % (size, i)] = rtable("%s/Test Sets/Test Set Size %d Sample %d.txt" % (gv, size, i))
    return d


def _he_row(args):
    X, y, a1b1, a2b2 = args
    pars = orc.sweep_candidates(a1b1, a2b2, 1000)
    ref = np.array([orc.cond_loglike_reference(X, y, 30.0, orc.FAMILY_ISO, row, 50.0) for row in pars])
    mini = np.array([orc.loglik_minimal(X, y, 30.0, orc.FAMILY_ISO, row, "tau", 50.0)["loglik"] for row in pars])
    return ref, mini


def _me_rows(args):
    D_old, pool, prm = args
    return orc.me_schur_negdet_batch(D_old, pool, prm)


def main():
    D = dict(np.load(os.path.join(os.path.dirname(os.path.dirname(OUT)), "convex-combination-of-gaussian-processes_b200", "data", "reference_designs.npz")))
    GV = gv_sets()
    np.savez_compressed(os.path.join(OUT, "gv_sets.npz"), **GV)
    G = {}
    pool = Pool(os.cpu_count() or 1)

    # ---- C3: every GV set: 8 candidates from the script's prior region + a small predictive table ----------
    for size, count in ((50, 9), (90, 8)):
        for i in range(1, count + 1):
            tr, te = GV["train%d_%d" % (size, i)], GV["test%d_%d" % (size, i)]
            X, y = tr[:, :9], tr[:, 9]
            rng = np.random.default_rng(100 * size + i)
            B = 8
            nat = np.column_stack([rng.uniform(0.05, 0.95, B), 0.06 / rng.gamma(3, 1.0, B), 1.0 / rng.gamma(5, 1 / 2.0, B)])
            ref = [orc.loglik_reference(X, y, 13.0, orc.FAMILY_ISO, q) for q in nat]
            tag = "gv%d_%d_" % (size, i)
            G[tag + "nat"] = nat
            G[tag + "ref"] = np.array([r["loglik"] for r in ref])
            G[tag + "beta"] = np.array([r["beta"] for r in ref])
            G[tag + "kappa"] = np.array([orc.cond1(orc.Mixed_corr_matrix_direct(X, orc.FAMILY_ISO, q)) for q in nat])
            m, v = orc.predict_table(X, y, 13.0, orc.FAMILY_ISO, nat[:2], te[:12, :9])
            G[tag + "pred_mean"], G[tag + "pred_var"] = m, v

    # ---- C2: choose.hyperpars on the heat exchanger, N = 1000, tau = 50, sigma2 = 30 -------------------------
    he, hp = D["he_train"], D["he_hyperpars"]
    X, y = he[:, :4], he[:, 4]
    res = pool.map(_he_row, [(X, y, hp[r, 0:2], hp[r, 2:4]) for r in HE_ROWS])
    G["he_rows"] = np.array(HE_ROWS)
    G["he_ref_loglik"] = np.stack([r[0] for r in res])            # (rows, 1000) reference-faithful (direct Sigma)
    G["he_min_loglik"] = np.stack([r[1] for r in res])            # (rows, 1000) minimal (Sherman-Morrison form)
    G["he_kappa_first8"] = np.array([orc.cond1(orc.Mixed_corr_matrix_direct(X, orc.FAMILY_ISO, q))
                                     for q in orc.sweep_candidates(hp[0, 0:2], hp[0, 2:4], 1000)[:8]])

    # ---- ME: full pool x 64 parameter rows -----------------------------------------------------------------
    D_old, pl = D["me_initial14"], D["me_all_subdesigns"]
    rng = np.random.default_rng(7007)
    Q = 63
    prm = np.vstack([[0.5, 1.0, 4.0], np.column_stack([rng.uniform(0, 1, Q), 1.0 / rng.gamma(3, 1 / 2.0, Q),
                                                       1.0 / rng.gamma(5, 1 / 16.0, Q)])])
    parts = pool.map(_me_rows, [(D_old, pl, prm[i:i + 8]) for i in range(0, 64, 8)])
    nd = np.hstack(parts)
    G["me64_params"] = prm
    G["me64_argmin"] = nd.argmin(axis=0)
    G["me64_min"] = nd.min(axis=0)
    srt = np.sort(nd, axis=0)
    G["me64_gap"] = (srt[1] - srt[0]) / np.abs(srt[0])            # relative gap between the best two designs

    # ---- subset log-dets, m = 128 and 256 (SURVEY 8d ME-B) ------------------------------------------------------
    gc = dict(np.load(os.path.join(OUT, "golden_cases.npz")))
    lhs, par = gc["sub_pool"], gc["sub_params"]
    rng = np.random.default_rng(256)
    for m_ in (128, 256):
        idx = np.array([rng.choice(lhs.shape[0], m_, replace=False) for _ in range(8)], dtype=np.int32)
        G["sub_idx_%d" % m_] = idx
        G["sub_logdet_%d" % m_] = np.array([orc.subset_logdet(lhs, ix, orc.FAMILY_ANISO_LAMBDA, par) for ix in idx])

    # ---- solve(R), beta.MLE on several rows: n = 14 aniso, n = 64 iso ----------------------------------------
    X14 = D["maximin14"]
    y14 = orc.test_function(4, X14[:, 0], X14[:, 1])
    p14 = gc["pred14_pars"]
    rr = [orc.loglik_reference(X14, y14, 0.9, orc.FAMILY_ANISO_LAMBDA, q) for q in p14]
    G["rinv14_all"] = np.stack([r["R_inv"] for r in rr])
    G["rinv14_all_beta"] = np.array([r["beta"] for r in rr])
    G["rinv14_all_kappa"] = np.array([orc.cond1(orc.Mixed_corr_matrix_direct(X14, orc.FAMILY_ANISO_LAMBDA, q)) for q in p14])
    p64 = gc["c2_nat"][:3]
    rr = [orc.loglik_reference(X, y, 30.0, orc.FAMILY_ISO, q) for q in p64]
    G["rinv64"] = np.stack([r["R_inv"] for r in rr])
    G["rinv64_beta"] = np.array([r["beta"] for r in rr])
    G["rinv64_kappa"] = gc["c2gls_kappa"][:3]

    pool.close()
    np.savez_compressed(os.path.join(OUT, "golden_r02.npz"), **G)
    print("wrote", len(GV), "GV arrays and", len(G), "golden arrays")
    print("ME best-two relative gap: min %.2e" % G["me64_gap"].min())
    print("HE ref-vs-minimal per-candidate max abs diff %.2e" % np.max(np.abs(G["he_ref_loglik"] - G["he_min_loglik"])))


if __name__ == "__main__":
    main()
