"""Generate the committed fixtures under tests/golden/ (run in the BUILD container,
where /root/reference exists; the GPU box never runs this).

  python tests/golden/make_golden.py

1. reference_designs.npz -- the reference's shipped *data inputs* (designs,
   hyper-prior grids, training/test sets) parsed from /root/reference's .txt
   files, so tests and bench.py can run where /root/reference is absent.
2. golden_cases.npz -- oracle outputs (reference-faithful path, minimal path,
   and a 50-digit mpmath truth on a few rows) on seeded inputs.

The reference publishes no numerical outputs for this path and R is not in the
container, so these vectors pin the ORACLE (parity unpinned w.r.t. real R).
"""
import os
import sys
import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import ccgp_oracle as orc  # noqa: E402

REF = "/root/reference"
OUT = os.path.dirname(os.path.abspath(__file__))


def rt(path, **kw):
    return np.loadtxt(os.path.join(REF, path), **kw)


def rtable(path):
    """R write.table output with header + row names -> float matrix."""
    rows = []
    with open(os.path.join(REF, path)) as fh:
        next(fh)
        for line in fh:
            parts = line.split()
            if parts:
                rows.append([float(v) for v in parts[1:]])
    return np.array(rows)


def designs():
    d = {}
    d["maximin14"] = rt("2D Codes and Designs/maximin 14 pts.txt")
    d["maximin100"] = rt("2D Codes and Designs/maximin 100 pts.txt")
    d["hyperpars_2d"] = rtable("2D Codes and Designs/hyperpars.matrix.txt")
    d["me_initial14"] = rt("Batch Sequential ME Designs/Initial ME Design.txt")
    d["me_all_subdesigns"] = rtable("Batch Sequential ME Designs/All_Subdesigns.txt").reshape(1000, 7, 2)
    d["me_plugin21"] = rtable("Batch Sequential ME Designs/Plug-in ME 14 plus 7 Design.txt")
    d["me_kmedoids21"] = rt("Batch Sequential ME Designs/k-medoids ME Design.txt", skiprows=1)
    d["me_maximin21"] = rt("Batch Sequential ME Designs/maximin 21 pts.txt")
    d["he_train"] = rtable("Heat Exchanger Emulator/Qian Training Set.txt")
    d["he_test"] = rtable("Heat Exchanger Emulator/Qian Test Set.txt")
    d["he_hyperpars"] = rtable("Heat Exchanger Emulator/hyperpars.matrix.txt")
    gv = "Ground Vibrations Emulator"
    d["gv50_train1"] = rtable(gv + "/Training Sets/Training Set Size 50 Sample 1.txt")
    d["gv50_test1"] = rtable(gv + "/Test Sets/Test Set Size 50 Sample 1.txt")
    d["gv90_train1"] = rtable(gv + "/Training Sets/Training Set Size 90 Sample 1.txt")
    d["gv90_test1"] = rtable(gv + "/Test Sets/Test Set Size 90 Sample 1.txt")
    d["design1d"] = rtable("1D Codes and Designs/1D Combined GP Simulation Designs.txt")
    return d


def nll_case(X, y, sigma2, family, nat, mean_mode, tau, n_truth):
    B = nat.shape[0]
    ref = np.empty(B)
    beta = np.empty(B)
    mini = np.empty(B)
    kap = np.empty(B)
    for b in range(B):
        if mean_mode == "gls":
            r = orc.loglik_reference(X, y, sigma2, family, nat[b])
            ref[b], beta[b] = r["loglik"], r["beta"]
        else:
            ref[b] = orc.cond_loglike_reference(X, y, sigma2, family, nat[b], tau)
            beta[b] = np.nan
        m = orc.loglik_minimal(X, y, sigma2, family, nat[b], mean_mode, tau)
        mini[b] = m["loglik"]
        if mean_mode != "gls":
            beta[b] = m["beta"]
        kap[b] = orc.cond1(orc.Mixed_corr_matrix_direct(X, family, nat[b]))
    truth = np.full(B, np.nan)
    for b in range(min(n_truth, B)):
        truth[b], _ = orc.loglik_truth(X, y, sigma2, family, nat[b], mean_mode, tau)
    return dict(ref=ref, beta=beta, minimal=mini, truth=truth, kappa=kap)


def main():
    D = designs()
    np.savez_compressed(os.path.join(os.path.dirname(os.path.dirname(OUT)), "convex-combination-of-gaussian-processes_b200", "data", "reference_designs.npz"), **D)
    G = {}

    # ---- C1: n=100 aniso, bench distribution (SURVEY 8d M1 primary) -------------
    X = D["maximin100"]
    y = orc.test_function(4, X[:, 0], X[:, 1])
    rng = np.random.default_rng(20131)
    B = 48
    theta = np.column_stack([rng.normal(np.log(20), 0.5, B), rng.normal(np.log(20), 0.5, B),
                             rng.normal(0, 1, B), rng.normal(0, 0.5, B)])
    nat = np.array([orc.transform_theta(orc.FAMILY_ANISO_LAMBDA, t, 2) for t in theta])
    r = nll_case(X, y, 1.0, orc.FAMILY_ANISO_LAMBDA, nat, "gls", 0.0, 3)
    G.update({"c1n100_theta": theta, "c1n100_nat": nat, "c1n100_y": y,
              **{"c1n100_" + k: v for k, v in r.items()}})
    # logpost val (adds Jacobian + [A] prior) on the first 8
    G["c1n100_logpost_val"] = np.array([orc.logpost(X, t, y, 1.0, orc.FAMILY_ANISO_LAMBDA, "A")["val"] for t in theta[:8]])

    # ---- C1: n=14 iso sweep (the reference's own empirical-Bayes loop) ----------
    X = D["maximin14"]
    y = orc.test_function(4, X[:, 0], X[:, 1])
    hp = D["hyperpars_2d"]
    cands = np.vstack([orc.sweep_candidates(hp[i, 0:2], hp[i, 2:4], 1728)[:24] for i in (0, 17, 59)])
    r = nll_case(X, y, 0.7, orc.FAMILY_ISO, cands, "tau", 100.0, 6)
    G.update({"c1n14_nat": cands, "c1n14_y": y, **{"c1n14_" + k: v for k, v in r.items()}})
    r = nll_case(X, y, 0.7, orc.FAMILY_ISO, cands, "gls", 0.0, 6)
    G.update({"c1n14gls_" + k: v for k, v in r.items()})
    # likeli.hyperpars for 2 grid rows with reduced N (kept small: oracle loops)
    G["c1n14_likeli_rows"] = np.array([0, 17])
    G["c1n14_likeli_N"] = np.array(256)
    G["c1n14_likeli"] = np.array([orc.likeli_hyperpars(X, y, hp[i, 0:2], hp[i, 2:4], 0.7, N=256, tau=100.0) for i in (0, 17)])
    G["halton_first16"] = orc.halton_base2(16)

    # ---- C2: heat exchanger n=64 d=4 iso ----------------------------------------
    he = D["he_train"]
    X, y = he[:, :4], he[:, 4]
    rng = np.random.default_rng(64)
    B = 24
    nat = np.column_stack([rng.uniform(0.05, 0.95, B), 1.0 / rng.gamma(3, 1 / 1.0, B), 1.0 / rng.gamma(5, 1 / 40.0, B)])
    r = nll_case(X, y, 30.0, orc.FAMILY_ISO, nat, "gls", 0.0, 2)
    G.update({"c2_nat": nat, **{"c2gls_" + k: v for k, v in r.items()}})
    r = nll_case(X, y, 30.0, orc.FAMILY_ISO, nat, "tau", 50.0, 2)
    G.update({"c2tau_" + k: v for k, v in r.items()})

    # ---- C3: ground vibrations n=50 / n=90, d=9 iso -----------------------------
    for tag in ("gv50", "gv90"):
        tr = D[tag + "_train1"]
        X, y = tr[:, :9], tr[:, 9]
        rng = np.random.default_rng(1)
        B = 16
        nat = np.column_stack([rng.uniform(0.05, 0.95, B), 0.02 / rng.gamma(3, 1.0, B) * 3, 1.0 / rng.gamma(5, 1 / 2.0, B)])
        r = nll_case(X, y, 13.0, orc.FAMILY_ISO, nat, "gls", 0.0, 1)
        G.update({tag + "_nat": nat, **{tag + "_" + k: v for k, v in r.items()}})

    # ---- predict tables ---------------------------------------------------------
    X = D["maximin14"]
    y = orc.test_function(4, X[:, 0], X[:, 1])
    rng = np.random.default_rng(5)
    S = 6
    pars = np.column_stack([rng.uniform(0.2, 0.8, S), rng.uniform(1, 6, S), rng.uniform(1, 6, S), rng.uniform(0.5, 3, S)])
    u = np.linspace(0, 1, 5)
    Xnew = np.array([[a, b] for b in u for a in u])        # expand.grid(u,u) ordering
    m, v = orc.predict_table(X, y, 0.9, orc.FAMILY_ANISO_LAMBDA, pars, Xnew)
    G.update({"pred14_pars": pars, "pred14_Xnew": Xnew, "pred14_mean": m, "pred14_var": v, "pred14_y": y})
    # quirk Q2 ([V]): matrix from (p, theta1, lambda), vector from (p, theta1, theta1*(1+lambda))
    parsV = np.column_stack([rng.uniform(0.2, 0.8, S), rng.uniform(1, 4, S), rng.uniform(2, 8, S)])
    parsVvec = np.column_stack([parsV[:, 0], parsV[:, 1], parsV[:, 1] * (1 + parsV[:, 2])])
    m, v = orc.predict_table(X, y, 0.9, orc.FAMILY_ISO_RAW2, parsV, Xnew, orc.FAMILY_ISO, parsVvec)
    G.update({"predV_pars": parsV, "predV_parsvec": parsVvec, "predV_mean": m, "predV_var": v})
    he, het = D["he_train"], D["he_test"]
    parsH = G["c2_nat"][:4]
    m, v = orc.predict_table(he[:, :4], he[:, 4], 30.0, orc.FAMILY_ISO, parsH, het[:, :4])
    G.update({"predHE_pars": parsH, "predHE_mean": m, "predHE_var": v})
    tr, te = D["gv50_train1"], D["gv50_test1"]
    parsG = G["gv50_nat"][:3]
    m, v = orc.predict_table(tr[:, :9], tr[:, 9], 13.0, orc.FAMILY_ISO, parsG, te[:20, :9])
    G.update({"predGV_pars": parsG, "predGV_mean": m, "predGV_var": v})

    # ---- R.Inv / factors on one candidate ---------------------------------------
    r = orc.loglik_reference(D["maximin14"], y, 0.9, orc.FAMILY_ANISO_LAMBDA, pars[0])
    G["rinv14"] = r["R_inv"]
    mf, vf1, vf2 = orc.factors(r["R_inv"], r["beta"], y)
    G.update({"rinv14_beta": np.array(r["beta"]), "rinv14_mean_factor": mf, "rinv14_var_factor1": vf1, "rinv14_var_factor2": np.array(vf2)})

    # ---- ME: Schur negdets over pool x params, Entropy --------------------------
    D_old = D["me_initial14"]
    pool = D["me_all_subdesigns"]
    rng = np.random.default_rng(7)
    Q = 5
    params = np.vstack([[0.5, 1.0, 4.0],
                        np.column_stack([rng.uniform(0, 1, Q), 1.0 / rng.gamma(3, 1 / 2.0, Q), 1.0 / rng.gamma(5, 1 / 16.0, Q)])])
    nd = orc.me_schur_negdet_batch(D_old, pool[:200], params)
    G.update({"me_params": params, "me_negdet_200": nd, "me_argmin_200": nd.argmin(axis=0)})
    nd_full = orc.me_schur_negdet_batch(D_old, pool, params[:1])
    G.update({"me_negdet_full_prior": nd_full[:, 0], "me_argmin_full_prior": np.array(int(nd_full[:, 0].argmin()))})
    G["entropy_initial14"] = np.array([orc.Entropy(D_old, *q) for q in params])
    G["entropy_pool21"] = np.array([orc.Entropy(np.vstack([D_old, pool[c]]), 0.5, 1.0, 4.0) for c in range(32)])

    # ---- subset log-dets on a synthetic pool ------------------------------------
    rng = np.random.default_rng(2048)
    N = 2048
    lhs = (np.column_stack([rng.permutation(N), rng.permutation(N)]) + rng.uniform(size=(N, 2))) / N * 2 - 1
    G["sub_pool"] = lhs
    par = np.array([0.5, 30.0, 60.0, 2.0])
    G["sub_params"] = par
    for m_ in (7, 21, 64):
        idx = np.array([rng.choice(N, m_, replace=False) for _ in range(12)], dtype=np.int32)
        G["sub_idx_%d" % m_] = idx
        G["sub_logdet_%d" % m_] = np.array([orc.subset_logdet(lhs, ix, orc.FAMILY_ANISO_LAMBDA, par) for ix in idx])

    # ---- C0: the 1-D scripts (n = 8, nu = 5): Matern+Matern and Matern+spline ------------
    rng = np.random.default_rng(8)
    X1 = D["design1d"][3].reshape(-1, 1)                       # one of the shipped 8-point designs
    y1 = np.sin(2 * np.pi * X1[:, 0]) + 0.5 * np.cos(9 * X1[:, 0])
    B = 16
    # the scripts' priors (theta1 ~ IG(3,2), theta2 ~ IG(5,16)) put most mass where an 8-point Matern(5) Gram
    # matrix has kappa > 1e10; the parity rows use shorter ranges so that kappa_1(R) <= 1e6 (stated gate)
    nat1 = np.column_stack([rng.uniform(0.1, 0.9, B), rng.uniform(0.04, 0.2, B), rng.uniform(0.08, 0.4, B)])
    Xn1 = np.linspace(0, 1, 11).reshape(-1, 1)
    G.update({"d1_X": X1, "d1_y": y1, "d1_nat": nat1, "d1_Xnew": Xn1})
    for tag, fam in (("d1mm", orc.FAMILY_MATERN1D), ("d1ms", orc.FAMILY_MATERN_SPLINE1D)):
        r = nll_case(X1, y1, 0.8, fam, nat1, "gls", 0.0, 0)
        G.update({tag + "_" + k: v for k, v in r.items()})
        m, v = orc.predict_table(X1, y1, 0.8, fam, nat1[:4], Xn1)
        G.update({tag + "_pred_mean": m, tag + "_pred_var": v})
        G[tag + "_R"] = orc.Mixed_corr_matrix(X1, fam, nat1[0])

    np.savez_compressed(os.path.join(OUT, "golden_cases.npz"), **G)
    print("wrote", len(D), "design arrays and", len(G), "golden arrays")
    for k in ("c1n100", "c1n14", "c1n14gls", "c2gls", "c2tau", "gv50", "gv90"):
        ref, mini, tr, kap = G[k + "_ref"], G[k + "_minimal"], G[k + "_truth"], G[k + "_kappa"]
        rel = np.abs(ref - mini) / np.maximum(np.abs(ref), 1)
        ok = ~np.isnan(tr)
        print("%-9s kappa med %.2e max %.2e | ref-vs-min max rel %.2e | ref-vs-truth %.2e | min-vs-truth %.2e" % (
            k, np.median(kap), kap.max(), np.nanmax(rel),
            np.max(np.abs(ref[ok] - tr[ok]) / np.maximum(np.abs(tr[ok]), 1)),
            np.max(np.abs(mini[ok] - tr[ok]) / np.maximum(np.abs(tr[ok]), 1))))


if __name__ == "__main__":
    main()
