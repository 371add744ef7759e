"""Host-side mirror of the reference's closures for the hot path, same names
(dots -> underscores), argument meaning and return shapes, every number coming
from the CUDA library.  R is not in this image, so this Python layer plays the
role r/ccgp.R plays for an R user (INTEGRATION.md shows both).

Scalar, per-call work that the reference does in R around the linear algebra
(the parameter transforms' Jacobian, the hard-coded log-priors, the Halton /
inverse-gamma candidate grid, optim's control flow) stays host-side here too.

Citations: [A] 2D ... Anisotropic Public.R, [I] 2D ... Isotropic Public.R,
[V] 2D ... Isotropic Advanced.R, [M] Batch Sequential ME Design.R,
[H] Combined GP Heat Exchanger.R, [G] Combined GP Ground Vibrations.R.
"""
from __future__ import annotations

import math
import numpy as np

from .engine import (Engine, GAUSS_ISO, GAUSS_ANISO_LAMBDA, GAUSS_ISO_RAW2, NATURAL, LOGSCALE,
                     MEAN_GLS_BETA, MEAN_ZERO_PLUS_TAU2)

_engines = {}


def default_engine(device: int = 0) -> Engine:
    if device not in _engines:
        _engines[device] = Engine(device)
    return _engines[device]


# ---- priors / Jacobian: the scalar tail of logpost -------------------------------
def log_jacobian(theta, family, d):
    """[A]:459, [I]:452: -phi - 2 log(1+exp(-phi)) + sum(psi) [+ zeta]; vectorised over rows."""
    th = np.atleast_2d(theta)
    if family == GAUSS_ANISO_LAMBDA:
        psi, phi, zeta = th[:, :d], th[:, d], th[:, d + 1]
        return -phi - 2.0 * np.log1p(np.exp(-phi)) + psi.sum(axis=1) + zeta
    psi1, psi2, phi = th[:, 0], th[:, 1], th[:, 2]
    return -phi - 2.0 * np.log1p(np.exp(-phi)) + psi1 + psi2


def log_prior(theta, script, prior_pars=None):
    """Per-script log-prior: 'A' [A]:462, 'I' [I]:453 (also [M]:450), 'G' [G]:450,
    'V'/'H' parametrised inverse-gamma [V]:467 / [H]:462 with prior_pars=(a1,b1,a2,b2)."""
    th = np.atleast_2d(theta)
    if script == "A":
        psi1, psi2, zeta = th[:, 0], th[:, 1], th[:, 3]
        return -psi1 - psi1 ** 2 / 2 - psi2 - psi2 ** 2 / 2 - 4 * zeta - 4 / np.exp(zeta)
    psi1, psi2 = th[:, 0], th[:, 1]
    t1, t2 = np.exp(psi1), np.exp(psi2)
    if script in ("I", "M", "D1", "D2"):   # [I]:453 = [M]:450 = [D1]:636 = [D2]:597
        return -4 * psi1 - 2 / t1 - 6 * psi2 - 16 / t2
    if script == "G":
        return -4 * psi1 - 1 / t1 - 6 * psi2 - 75 / t2
    if script in ("V", "H"):
        a1, b1, a2, b2 = prior_pars
        return -(a1 + 1) * psi1 - b1 / t1 - (a2 + 1) * psi2 - b2 / t2
    raise ValueError("unknown script %r" % (script,))


_SCRIPT_FAMILY = {"A": GAUSS_ANISO_LAMBDA, "I": GAUSS_ISO, "M": GAUSS_ISO, "G": GAUSS_ISO, "H": GAUSS_ISO, "D1": 3, "D2": 4,
                  "V": GAUSS_ISO_RAW2}


def transform(theta, family, d):
    """Real-line rows -> natural-scale rows (p, theta.., [lambda]) ([A]:435-442)."""
    th = np.atleast_2d(np.asarray(theta, dtype=np.float64))
    if family == GAUSS_ANISO_LAMBDA:
        return np.column_stack([1.0 / (1.0 + np.exp(-th[:, d])), np.exp(th[:, :d]), np.exp(th[:, d + 1])])
    return np.column_stack([1.0 / (1.0 + np.exp(-th[:, 2])), np.exp(th[:, 0]), np.exp(th[:, 1])])


# ---- L1: correlation blocks --------------------------------------------------------
def Mixed_corr_matrix(D_train, p, theta1, theta2, lam=None, engine=None):
    """[A]:399-406 (lam given: anisotropic, theta1/theta2 are the per-dimension scales of a
    2-D design) or [I]:400-407 (lam None: isotropic)."""
    eng = engine or default_engine()
    D = np.atleast_2d(D_train)
    if lam is None:
        return eng.mixed_corr([p, theta1, theta2], GAUSS_ISO, D)
    return eng.mixed_corr([p, theta1, theta2, lam], GAUSS_ANISO_LAMBDA, D)


def Mixed_corr_vec(x_new, D_train, p, theta1, theta2, lam=None, engine=None):
    """[A]:416-422 / [I]:417-423 -> r (length n)."""
    eng = engine or default_engine()
    x = np.asarray(x_new, dtype=np.float64).reshape(1, -1)
    if lam is None:
        return eng.mixed_corr([p, theta1, theta2], GAUSS_ISO, x, D_train).reshape(-1)
    return eng.mixed_corr([p, theta1, theta2, lam], GAUSS_ANISO_LAMBDA, x, D_train).reshape(-1)


def cross_corr_matrix(D_old, D_new, theta, engine=None):
    """[M]:835-848 -> n_new x n_old iso Gaussian cross-Gram."""
    eng = engine or default_engine()
    return eng.mixed_corr([1.0, theta, theta], GAUSS_ISO, D_new, D_old)


# ---- L2: likelihood -----------------------------------------------------------------
R_EPS = 2.220446049250313e-16      # .Machine$double.eps, solve()'s default tol


def logpost_batch(D_train, theta, y, sigma2, script="A", prior_pars=None, engine=None, want_rinv=False, na_rule="rcond"):
    """Batched `logpost`: theta is a (B x k) matrix of real-line rows.
    -> dict(val[B], beta[B], loglik[B], status[B][, R_Inv[B,n,n]]); NaN where R gives NA ([A]:448-449).
    na_rule "rcond" (default) reproduces `try(solve(R))`: NA when the factorisation breaks down (status 1) or
    rcond_1(R) < .Machine$double.eps (status 2; one extra ccgp_rcond_batch call); "pivot" keeps only status 1."""
    eng = engine or default_engine()
    family = _SCRIPT_FAMILY[script]
    D = np.atleast_2d(D_train)
    eng.set_design(D, y)
    th = np.atleast_2d(np.asarray(theta, dtype=np.float64))
    nll, beta, status = eng.nll_batch(th, family, sigma2, scale=LOGSCALE)
    if na_rule == "rcond":
        rc, _, _ = eng.rcond_batch(th, family, scale=LOGSCALE)
        na = (status == 0) & ~(rc >= R_EPS)
        status = status.copy()
        status[na] = 2
        nll = np.where(na, np.nan, nll)
        beta = np.where(na, np.nan, beta)
    val = -nll + log_jacobian(th, family, D.shape[1]) + log_prior(th, script, prior_pars)
    out = dict(val=val, beta=beta, loglik=-nll, status=status)
    if want_rinv:
        out["R_Inv"] = eng.rinv_batch(th, family, scale=LOGSCALE)[0]
    return out


def logpost(D_train, theta, y, sigma2, script="A", prior_pars=None, engine=None):
    """`logpost(D.train, theta, y, sigma2[, theta.pars, lambda.pars])` ([A]:433-467,
    [I]:433-457, [V]:447-471, [G]:429-454) -> list(val, beta, R.Inv)."""
    r = logpost_batch(D_train, np.asarray(theta, dtype=np.float64).reshape(1, -1), y, sigma2, script, prior_pars,
                      engine, want_rinv=True)
    return dict(val=float(r["val"][0]), beta=float(r["beta"][0]), R_Inv=r["R_Inv"][0], like=float(np.exp(r["loglik"][0])))


def halton_base2(N):
    """fOptions::runif.halton(N, 1) ([V]:557): base-2 radical inverse of 1..N."""
    i = np.arange(1, N + 1, dtype=np.uint64)
    out = np.zeros(N)
    f = 0.5
    while np.any(i > 0):
        out += f * (i & np.uint64(1))
        i >>= np.uint64(1)
        f *= 0.5
    return out


def qigamma(p, alpha, beta):
    """pscl::qigamma ([V]:558-559) = 1 / qgamma(1 - p, alpha, rate = beta)."""
    from scipy.stats import gamma
    return 1.0 / gamma.ppf(1.0 - np.asarray(p), a=alpha, scale=1.0 / beta)


def sweep_candidates(theta1_pars, theta2_pars, N):
    """[V]:557-560: pars <- cbind(p, theta1, theta2) from ONE Halton stream (quirk Q5)."""
    u = halton_base2(N)
    return np.column_stack([u, qigamma(u, *theta1_pars), qigamma(u, *theta2_pars)])


def likeli_hyperpars(D_train, y_train, theta1_pars, theta2_pars, sigma2, N=1728, tau=100.0, engine=None):
    """[V]:552-578 (N=1728, tau=100; [H]:549-575 uses N=1000, tau=50): mean over the N
    quasi-random candidates of exp(cond.like)."""
    eng = engine or default_engine()
    eng.set_design(D_train, y_train)
    pars = sweep_candidates(theta1_pars, theta2_pars, N)
    nll, _, _ = eng.nll_batch(pars, GAUSS_ISO, sigma2, mean_mode=MEAN_ZERO_PLUS_TAU2, tau=tau)
    return float(np.mean(np.exp(-nll)))


def choose_hyperpars(D_train, y_train, hyperpars_matrix, sigma2, N=1728, tau=100.0, take_log=False, engine=None):
    """[V]:588-599: all H x N candidates in ONE batched call, then which.max per [V]:598
    (raw means) or [H]:591 (take_log=True).  -> dict(pars, likelihoods)."""
    eng = engine or default_engine()
    eng.set_design(D_train, y_train)
    hp = np.atleast_2d(hyperpars_matrix)
    H = hp.shape[0]
    # qigamma(u, a, b) = b / qgamma(1 - u, a): the quantile function is needed once per distinct SHAPE only
    # (the grids hold a handful of shapes), the scale is a multiplication -- same numbers as sweep_candidates
    from scipy.stats import gamma
    u = halton_base2(N)
    qg = {a: gamma.ppf(1.0 - u, a=a) for a in np.unique(hp[:, [0, 2]])}
    cand = np.empty((H * N, 3))
    for i in range(H):
        blk = cand[i * N:(i + 1) * N]
        blk[:, 0] = u
        blk[:, 1] = 1.0 / (qg[hp[i, 0]] * (1.0 / hp[i, 1]))
        blk[:, 2] = 1.0 / (qg[hp[i, 2]] * (1.0 / hp[i, 3]))
    nll, _, _ = eng.nll_batch(cand, GAUSS_ISO, sigma2, mean_mode=MEAN_ZERO_PLUS_TAU2, tau=tau)
    likes = np.exp(-nll).reshape(H, N).mean(axis=1)
    if take_log:
        likes = np.log(likes)
    return dict(pars=hp[int(np.argmax(likes))], likelihoods=likes)


# ---- prediction -----------------------------------------------------------------------
def factors(R_Inv, beta, y_train):
    """[A]:550-559 on host arrays already returned by logpost (O(n^2), kept for API parity)."""
    y = np.asarray(y_train, dtype=np.float64).reshape(-1)
    mean_factor = R_Inv @ (y - beta)
    var_factor1 = R_Inv.sum(axis=0)
    return np.concatenate([mean_factor, var_factor1, [var_factor1.sum()]])


def predict_post_batch(D_new, D_train, y_train, pars, sigma2, script="A", engine=None):
    """The S x T table `prediction` builds by calling predict.post per (row, site)
    ([A]:637-654 -> [A]:604-623) in one call.  pars: S natural-scale rows
    (p, theta1, theta2[, lambda]).  -> (mean[T,S], var[T,S]).
    script 'V' reproduces quirk Q2: the matrix uses scale `lambda`, the vector
    theta1*(1+lambda) ([V]:672)."""
    eng = engine or default_engine()
    eng.set_design(D_train, y_train)
    pars = np.atleast_2d(np.asarray(pars, dtype=np.float64))
    family = _SCRIPT_FAMILY[script]
    if family == GAUSS_ISO_RAW2:
        pv = np.column_stack([pars[:, 0], pars[:, 1], pars[:, 1] * (1.0 + pars[:, 2])])
        mean, var, _ = eng.predict(pars, GAUSS_ISO_RAW2, D_new, sigma2, pars_vec=pv, vec_family=GAUSS_ISO)
    else:
        mean, var, _ = eng.predict(pars, family, D_new, sigma2)
    return mean, var


def factors_device(D_train, y_train, pars, script="A", engine=None):
    """`factors.frame` ([A]:572-592) without its wire format: factor the S posterior rows once and keep the Cholesky
    factors on the device, keyed by the row index (Engine.factors -> ccgp_factors_create).  The returned object's
    predict(D_new, sigma2) gives the tables of `predict_post_batch` for any number of site sets (bit-identical)."""
    eng = engine or default_engine()
    eng.set_design(D_train, y_train)
    pars = np.atleast_2d(np.asarray(pars, dtype=np.float64))
    family = _SCRIPT_FAMILY[script]
    if family == GAUSS_ISO_RAW2:                       # quirk Q2 ([V]:672)
        pv = np.column_stack([pars[:, 0], pars[:, 1], pars[:, 1] * (1.0 + pars[:, 2])])
        return eng.factors(pars, GAUSS_ISO_RAW2, pars_vec=pv, vec_family=GAUSS_ISO)
    return eng.factors(pars, family)


def predict_post_factors(ff, D_new, sigma2):
    """(mean[T,S], var[T,S]) at the sites D_new from the factors `factors_device` left on the device."""
    mean, var, _ = ff.predict(D_new, sigma2)
    return mean, var


def predict_post(x_new, D_train, y_train, pars, sigma2, script="A", engine=None):
    """`predict.post(x.new, D.train, pars, sigma2)` -> cbind(mean, var) (1 x 2)."""
    m, v = predict_post_batch(np.asarray(x_new, dtype=np.float64).reshape(1, -1), D_train, y_train,
                              np.asarray(pars, dtype=np.float64).reshape(1, -1), sigma2, script, engine)
    return np.array([[m[0, 0], v[0, 0]]])


# ---- ME design --------------------------------------------------------------------------
def Entropy(D, p, theta1, theta2, engine=None):
    """[M]:856-861: -det(Mixed.corr.matrix(D, p, theta1, theta2))."""
    eng = engine or default_engine()
    nd, _, _ = eng.me_schur_batch(None, np.asarray(D, dtype=np.float64)[None], [[p, theta1, theta2]])
    return float(nd[0, 0])


def Augmented_Mixed_Entropy(D_old, D_new, p, theta1, theta2, R_old_Inv=None, engine=None):
    """[M]:869-877.  R.old.Inv is accepted for signature parity and ignored: the kernel
    refactors R.old itself (it is part of the same Cholesky)."""
    eng = engine or default_engine()
    nd, _, _ = eng.me_schur_batch(D_old, np.asarray(D_new, dtype=np.float64)[None], [[p, theta1, theta2]])
    return float(nd[0, 0])


def entropy_batch(D_old, D_new_pool, params, engine=None):
    """Batched sibling: negdet[C, P] for a pool of candidate second batches x parameter rows."""
    eng = engine or default_engine()
    return eng.me_schur_batch(D_old, D_new_pool, params)[0]


def Batch_Entropy_optim(D_old, n_new, d, p, theta1, theta2, n_starts, rng=None, engine=None, maxiter=100):
    """[M]:920-948: multi-start L-BFGS-B on [-1,1]^(n_new*d); each finite-difference
    gradient is ONE batched GPU call over the 2*n_new*d stencil (optim's ndeps=1e-3
    central differences).  -> dict(Design, log_entropy) with log_entropy = -min.val
    (a determinant, quirk Q4).  Starts are random LHDs (lhs::optimumLHS stays caller-side).
    Stopping rule as the reference's optim call ([M]:936: maxit = 100, pgtol = 0, factr = 1e7) -- the criterion is a
    tiny determinant, scipy's default gtol = 1e-5 would stop at the start point -- and the difference stencil is
    clipped to the box like optim's (the step actually taken divides the difference)."""
    from scipy.optimize import minimize
    eng = engine or default_engine()
    rng = rng or np.random.default_rng()
    D_old = None if D_old is None else np.atleast_2d(D_old)
    m = n_new * d
    par_row = [[p, theta1, theta2]]

    def f_and_g(x):
        pts = np.repeat(x[None, :], 2 * m + 1, axis=0)
        h = 1e-3
        up = np.minimum(x + h, 1.0)
        dn = np.maximum(x - h, -1.0)
        for i in range(m):
            pts[1 + 2 * i, i] = up[i]
            pts[2 + 2 * i, i] = dn[i]
        designs = pts.reshape(-1, d, n_new).transpose(0, 2, 1)   # c(D) is column-major
        nd = eng.me_schur_batch(D_old, designs, par_row)[0][:, 0]
        g = (nd[1::2] - nd[2::2]) / (up - dn)
        return float(nd[0]), g

    vals, designs = [], []
    for _ in range(n_starts):
        lhd = (np.argsort(rng.random((n_new, d)), axis=0) + rng.random((n_new, d))) / n_new
        start = (-1.0 + 2.0 * lhd).reshape(-1, order="F")
        res = minimize(f_and_g, start, jac=True, method="L-BFGS-B", bounds=[(-1.0, 1.0)] * m,
                       options=dict(maxiter=maxiter, gtol=0.0, ftol=1e7 * np.finfo(float).eps))
        vals.append(res.fun)
        designs.append(res.x.reshape(n_new, d, order="F"))
    k = int(np.argmin(vals))
    return dict(Design=designs[k], log_entropy=-vals[k])


def Entropy_optim(n, d, p, theta1, theta2, n_starts, rng=None, engine=None, maxiter=100):
    """[M]:886-912: first-batch ME design (no D.old)."""
    return Batch_Entropy_optim(None, n, d, p, theta1, theta2, n_starts, rng, engine, maxiter)


def kmedoids_design(D_old, subdesigns, k, engine=None):
    """The clustering step behind `k-medoids ME Design.txt` (reference ReadMe.md:54-60): k-medoids
    (cluster::pam) over the points of all second-batch designs; the medoids, appended to D.old, are
    the next batch.  subdesigns: (C, n_new, d) or (C*n_new, d).  -> dict(Design, medoid_rows, cost)."""
    eng = engine or default_engine()
    P = np.asarray(subdesigns, dtype=np.float64)
    P = P.reshape(-1, P.shape[-1])
    med, cost, _ = eng.kmedoids_pam(P, k)
    D_old = np.atleast_2d(np.asarray(D_old, dtype=np.float64))
    return dict(Design=np.vstack([D_old, P[med]]), medoid_rows=med, cost=cost)


# ---- CGP comparator: the start sweep and the leave-one-out loop of `CGP` ([A]:60-237) ----------
def cgp_standardise(X):
    """[A]:70-72: per column (x - min) / (max - min) and the scales max - min."""
    X = np.atleast_2d(np.asarray(X, dtype=np.float64))
    lo = X.min(axis=0)
    scales = X.max(axis=0) - lo
    return (X - lo) / scales, scales


def cgp_bounds(X, nugget_l=0.001, theta_l=1e-4):
    """[A]:79-92: the box (lower, upper) of (lambda, theta_1..p, kappa, bw) the sweep and optim's L-BFGS-B use."""
    Xs, _ = cgp_standardise(X)
    n, p = Xs.shape
    iu = np.triu_indices(n, 1)
    d2 = ((Xs[:, None, :] - Xs[None, :, :]) ** 2).sum(axis=2)[iu]
    m = np.mean(1.0 / d2)
    alpha_l, kappa_u = np.log(10.0 ** 2) * m, np.log(10.0 ** 6) * m
    return (np.concatenate([[nugget_l], np.full(p, theta_l), [alpha_l, 0.0]]),
            np.concatenate([[1.0], np.full(p, alpha_l), [kappa_u, 1.0]]))


def cgp_start_candidates(X, num_starts=5, rng=None):
    """[A]:137-147: the (500 + num_starts) x (p + 3) random Latin hypercube of start candidates, scaled to the box."""
    rng = rng or np.random.default_rng()
    lower, upper = cgp_bounds(X)
    N, k = 500 + num_starts, len(lower)
    lhd = (np.column_stack([rng.permutation(N) + 1 for _ in range(k)]) - 0.5) / N
    return lhd * (upper - lower) + lower


def var_MLE_DK_batch(X, yobs, starts, engine=None):
    """`apply(starts, 1, var.MLE.DK)` ([A]:148) in one launch -> cand_obj."""
    eng = engine or default_engine()
    Xs, _ = cgp_standardise(X)
    return eng.cgp_objective_batch(Xs, np.asarray(yobs, dtype=np.float64), starts)[0]


def cgp_best_starts(X, yobs, starts, num_starts=5, engine=None):
    """[A]:148-150: the rows whose objective ranks among the num_starts smallest (rank(.., ties = "min") <= num_starts)."""
    obj = var_MLE_DK_batch(X, yobs, starts, engine)
    rank = np.array([1 + np.sum(obj < v) for v in obj])
    return np.asarray(starts)[rank <= num_starts], obj


def cgp_jackknife(X, yobs, par, engine=None):
    """Yp_jackknife and rmscv ([A]:166-201) for the fitted row par = op$par (lambda, Stand_theta, kappa, bw)."""
    eng = engine or default_engine()
    Xs, _ = cgp_standardise(X)
    yobs = np.asarray(yobs, dtype=np.float64)
    yp = eng.cgp_jackknife(Xs, yobs, par)[0]
    return dict(Yp_jackknife=yp, rmscv=float(np.sqrt(np.sum((yobs - yp) ** 2) / len(yobs))))
