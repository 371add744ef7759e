"""Error-versus-conditioning study of the NLL path (SURVEY 8d parity gate).

Design: `maximin 100 pts` ([-1,1]^2), y = simulator 4 ([A]:338), sigma2 = 1, family GAUSS_ANISO_LAMBDA --
the headline configuration.  Candidates are real-line rows (psi1, psi2, phi, zeta) of `logpost` ([A]:433-442).

  make  (CPU, no GPU needed; run once, output committed as tests/golden/kappa_study.npz)
        * curve set: draws from the script's own prior ([A]:462: psi ~ N(-1, 1), lambda ~ IG(4, 4), p ~ U(0,1))
          with psi shifted by U(0, 5.5) so that kappa_1(R) covers 1e2 .. 1e17, `PER_DECADE` rows per decade of
          kappa_1; for each row: kappa_1 (dgecon, what R's rcond uses), the reference-faithful oracle
          (solve + dmnorm, NA by the rcond < eps rule of [A]:448-449), the minimal oracle, the 50-digit truth.
        * NA set: `N_NA` unshifted prior draws + `N_NA` shifted ones; per row kappa_1, the reference's NA flag,
          and the smallest Cholesky pivot of the direct-difference matrix (host model of the kernel's flag).
  gpu   (GPU box) evaluates the kernel on both sets and writes the report (profiles/kappa_curve.txt).
"""
import os
import sys
import time
from multiprocessing import Pool

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
FIXTURE = os.path.join(ROOT, "tests", "golden", "kappa_study.npz")
PER_DECADE = 5
N_NA = 4000
DECADES = list(range(1, 18))          # bucket k: 10^k <= kappa_1 < 10^(k+1)


def _design():
    import ccgp_b200  # noqa: F401  (registers the package alias)
    from ccgp_b200 import workloads
    return workloads.m1_design()


def prior_draws(rng, B, shift_max):
    """[A]:462 as a sampler: log.prior + log.jacob = N(-1,1) on psi1, psi2; IG(4,4) on lambda = e^zeta; U(0,1) on p."""
    psi = rng.normal(-1.0, 1.0, (B, 2)) + rng.uniform(0.0, shift_max, (B, 1))
    p = rng.uniform(0.0, 1.0, B)
    lam = 1.0 / rng.gamma(4.0, 1.0 / 4.0, B)
    return np.column_stack([psi, np.log(p / (1.0 - p)), np.log(lam)])


def _cheap(row):
    from oracle import ccgp_oracle as orc
    X, y, s2 = _design()
    nat = orc.transform_theta(orc.FAMILY_ANISO_LAMBDA, row, 2)
    R = orc.Mixed_corr_matrix(X, orc.FAMILY_ANISO_LAMBDA, nat)
    kap = orc.cond1(R)
    ref = orc.loglik_reference(X, y, s2, orc.FAMILY_ANISO_LAMBDA, nat)
    Rd = orc.Mixed_corr_matrix_direct(X, orc.FAMILY_ANISO_LAMBDA, nat)
    # smallest pivot of an unblocked Cholesky of the direct-difference matrix (<= 0: breakdown)
    try:
        L = np.linalg.cholesky(Rd)
        minpiv = float(np.min(np.diag(L)) ** 2)
    except np.linalg.LinAlgError:
        minpiv = 0.0
    return kap, ref["loglik"], ref["beta"], ref["status"], minpiv


def _truth(row):
    from oracle import ccgp_oracle as orc
    X, y, s2 = _design()
    nat = orc.transform_theta(orc.FAMILY_ANISO_LAMBDA, row, 2)
    mn = orc.loglik_minimal(X, y, s2, orc.FAMILY_ANISO_LAMBDA, nat)
    tr = orc.loglik_truth(X, y, s2, orc.FAMILY_ANISO_LAMBDA, nat)
    return mn["loglik"], mn["beta"], tr[0], tr[1]


def make():
    rng = np.random.default_rng(448)
    workers = os.cpu_count() or 1
    t0 = time.time()
    with Pool(workers) as pool:
        na_rows = np.vstack([prior_draws(rng, N_NA, 0.0), prior_draws(rng, N_NA, 5.5)])
        na = np.array(pool.map(_cheap, list(na_rows), chunksize=64))
        print("NA set: %d rows in %.0f s" % (len(na_rows), time.time() - t0))
        kap = na[:, 0]
        dec = np.floor(np.log10(np.maximum(kap, 1.0))).astype(int)
        pick = []
        for k in DECADES:
            idx = np.flatnonzero(dec == k)[:PER_DECADE]
            pick.extend(idx.tolist())
        pick = np.array(pick)
        curve_rows = na_rows[pick]
        t0 = time.time()
        tr = np.array(pool.map(_truth, list(curve_rows), chunksize=1))
        print("curve set: %d rows, truth in %.0f s" % (len(curve_rows), time.time() - t0))
    np.savez_compressed(
        FIXTURE,
        curve_theta=curve_rows, curve_kappa=kap[pick], curve_ref_ll=na[pick, 1], curve_ref_beta=na[pick, 2],
        curve_ref_status=na[pick, 3].astype(np.int32), curve_min_ll=tr[:, 0], curve_min_beta=tr[:, 1],
        curve_truth_ll=tr[:, 2], curve_truth_beta=tr[:, 3],
        na_theta=na_rows, na_kappa=kap, na_ref_status=na[:, 3].astype(np.int32), na_minpiv=na[:, 4],
        na_shifted=np.concatenate([np.zeros(N_NA, np.int32), np.ones(N_NA, np.int32)]))
    print("wrote", FIXTURE)


def _rel(a, b):
    return np.abs(a - b) / np.maximum(np.abs(b), 1.0)


def report(gpu_nll, gpu_beta, gpu_st, na_st, fx, na_rcond=None):
    """Text of profiles/kappa_curve.txt from the kernel's outputs on the two sets."""
    out = []
    w = out.append
    kap = fx["curve_kappa"]
    ref_ll, tr_ll, mn_ll = fx["curve_ref_ll"], fx["curve_truth_ll"], fx["curve_min_ll"]
    ref_b, tr_b = fx["curve_ref_beta"], fx["curve_truth_beta"]
    ref_st = fx["curve_ref_status"]
    w("# error vs kappa_1(R): maximin 100 pts, simulator 4, sigma2 = 1, GAUSS_ANISO_LAMBDA; %d rows per decade" % PER_DECADE)
    w("# loglik errors are relative (|a-b| / max(|b|,1)); 'ref' = reference-faithful oracle (solve + dmnorm, what R runs),")
    w("# 'truth' = 50-digit mpmath; NA(ref) = R's rcond < 2.2e-16 rule ([A]:448-449); NA(gpu) = kernel status != 0")
    w("%-11s %4s %10s %10s %10s %10s %10s %8s %8s" % ("kappa_1", "rows", "gpu-truth", "ref-truth", "min-truth", "gpu-ref",
                                                      "b:gpu-tru", "NA(ref)", "NA(gpu)"))
    dec = np.floor(np.log10(kap)).astype(int)
    for k in DECADES:
        m = dec == k
        if not m.any():
            continue
        okg = m & (gpu_st == 0)
        okr = m & (ref_st == 0)
        both = okg & okr

        def mx(a):
            return ("%10.2e" % np.max(a)) if a.size else "%10s" % "-"
        w("1e%02d..1e%02d %4d %s %s %s %s %s %8d %8d" % (
            k, k + 1, m.sum(), mx(_rel(-gpu_nll[okg], tr_ll[okg])), mx(_rel(ref_ll[okr], tr_ll[okr])),
            mx(_rel(mn_ll[m], tr_ll[m])), mx(_rel(-gpu_nll[both], ref_ll[both])), mx(_rel(gpu_beta[okg], tr_b[okg])),
            (ref_st[m] != 0).sum(), (gpu_st[m] != 0).sum()))
    w("")
    w("# NA-set agreement on %d prior draws ([A]:462 as a sampler) + %d draws with psi shifted by U(0,5.5)" % (N_NA, N_NA))
    nk = fx["na_kappa"]
    nref = fx["na_ref_status"] != 0
    ngpu = na_st != 0
    w("%-11s %6s %8s %8s %10s %10s" % ("kappa_1", "rows", "NA(ref)", "NA(gpu)", "gpu-only", "ref-only"))
    nd = np.floor(np.log10(np.maximum(nk, 1.0))).astype(int)
    for k in sorted(set(nd.tolist())):
        m = nd == k
        w("1e%02d..1e%02d %6d %8d %8d %10d %10d" % (k, k + 1, m.sum(), nref[m].sum(), ngpu[m].sum(),
                                                     (ngpu & ~nref & m).sum(), (nref & ~ngpu & m).sum()))
    w("total       %6d %8d %8d %10d %10d" % (len(nk), nref.sum(), ngpu.sum(), (ngpu & ~nref).sum(), (nref & ~ngpu).sum()))
    for tag, sel in (("unshifted prior draws", fx["na_shifted"] == 0), ("shifted draws", fx["na_shifted"] == 1)):
        w("%s: %d rows, NA(ref) %d, NA(gpu) %d, disagree %d" % (tag, sel.sum(), nref[sel].sum(), ngpu[sel].sum(),
                                                                 (nref[sel] != ngpu[sel]).sum()))
    if na_rcond is not None:
        w("")
        w("# the same with the wrappers' rule (logpost / logpost.batch): NA when status != 0 OR ccgp_rcond_batch < 2.220446e-16")
        nrc = ngpu | ~(na_rcond >= 2.220446049250313e-16)
        w("%-11s %6s %8s %8s %10s %10s" % ("kappa_1", "rows", "NA(ref)", "NA(gpu)", "gpu-only", "ref-only"))
        for k in sorted(set(nd.tolist())):
            m = nd == k
            if (nref[m] != nrc[m]).any() or (nref[m].any() and not nref[m].all()):
                w("1e%02d..1e%02d %6d %8d %8d %10d %10d" % (k, k + 1, m.sum(), nref[m].sum(), nrc[m].sum(),
                                                             (nrc & ~nref & m).sum(), (nref & ~nrc & m).sum()))
        w("total       %6d %8d %8d %10d %10d" % (len(nk), nref.sum(), nrc.sum(), (nrc & ~nref).sum(), (nref & ~nrc).sum()))
        fin = np.isfinite(na_rcond) & (na_rcond > 0) & np.isfinite(nk)
        ratio = (1.0 / na_rcond[fin]) / nk[fin]
        w("exact kappa_1 (gpu, from the explicit inverse) / dgecon estimate (oracle): median %.3f, 1%% %.3f, 99%% %.3f" % (
            np.median(ratio), np.quantile(ratio, 0.01), np.quantile(ratio, 0.99)))
    return "\n".join(out) + "\n"


def gpu(out_path):
    import ccgp_b200
    from ccgp_b200 import GAUSS_ANISO_LAMBDA, LOGSCALE
    fx = dict(np.load(FIXTURE))
    X, y, s2 = _design()
    eng = ccgp_b200.Engine(0)
    eng.set_design(X, y)
    nll, beta, st = eng.nll_batch(fx["curve_theta"], GAUSS_ANISO_LAMBDA, s2, scale=LOGSCALE)
    _, _, na_st = eng.nll_batch(fx["na_theta"], GAUSS_ANISO_LAMBDA, s2, scale=LOGSCALE)
    na_rc, _, _ = eng.rcond_batch(fx["na_theta"], GAUSS_ANISO_LAMBDA, scale=LOGSCALE)
    txt = report(nll, beta, st, na_st, fx, na_rc)
    eng.close()
    with open(out_path, "w") as f:
        f.write(txt)
    print(txt)


if __name__ == "__main__":
    if len(sys.argv) > 1 and sys.argv[1] == "make":
        make()
    else:
        gpu(sys.argv[2] if len(sys.argv) > 2 else os.path.join(ROOT, "gpurun_out", "kappa_curve.txt"))
