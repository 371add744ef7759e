"""CPU oracle for the Combined-GP hot path -- TEST INFRASTRUCTURE ONLY.

This module is a NumPy/SciPy (LAPACK) restatement of the numerical core of the
reference R scripts (oharari/Convex-Combination-of-Gaussian-Processes).  Only
`tests/`, `__graft_entry__.smoke()` and `bench.py`'s cpu_baseline / `--impl
reference` legs may import it; the product path (the CUDA library) never does.

PARITY UNPINNED: the reference ships no tests, golden vectors or stored outputs
of this path, and neither R nor Rscript exists in the build container, so the
oracle could not be checked against the reference running here.  It restates
the R functions line by line, using the same LAPACK routines base R calls
(dgesv+dgecon for `solve`, dgetrf for `det`, dpotrf/dpotri for `chol`/
`chol2inv`), and is cross-checked three ways in tests/: reference-faithful path
vs minimal (Cholesky) path vs a 50-digit mpmath truth.
Two pieces ARE pinned by shipped files: `pam_kmedoids` reproduces rows 15-21 of
`k-medoids ME Design.txt` from `All_Subdesigns.txt` exactly (tests/test_kmedoids.py); and the
likelihood + predictor, driven through the Metropolis loop, reproduce the reference's only
stored output (`Ground Vibrations Emulator/Results/Size 50 Results 1.txt`) statistically:
RMSPE and the predictions themselves (tests/test_end_to_end_gv.py; their chain is unseeded).

File aliases (all under /root/reference):
  [A] 2D Codes and Designs/2D Combined GP Anisotropic Public.R
  [I] 2D Codes and Designs/2D Combined GP Isotropic Public.R
  [V] 2D Codes and Designs/2D Combined GP Isotropic Advanced.R
  [M] Batch Sequential ME Designs/Batch Sequential ME Design.R
  [H] Heat Exchanger Emulator/Combined GP Heat Exchanger.R
  [G] Ground Vibrations Emulator/Combined GP Ground Vibrations.R

Third-party arithmetic on the path that is NOT in /root/reference (un-pinned,
no DESCRIPTION/renv.lock): `mnormt::dmnorm` (call sites [A]:455 [V]:465,573).
Two published implementations are restated: `dmnorm_pdsolve` (mnormt >= 1.5,
pd.solve = chol + chol2inv) and `dmnorm_legacy` (mnormt 1.4, solve + qr).
"""
from __future__ import annotations

import math
import numpy as np
from scipy.linalg import lapack

LOG2PI = math.log(2.0 * math.pi)
EPS = np.finfo(np.float64).eps

FAMILY_ISO = 0           # [I]:400-407  params (p, theta1, theta2)
FAMILY_ANISO_LAMBDA = 1  # [A]:399-406  params (p, theta_1..theta_d, lambda)
FAMILY_ISO_RAW2 = 2      # [V]:414-421  params (p, theta1, lambda): R2 uses `lambda` as its scale
FAMILY_MATERN1D = 3      # [D1] "1D Combined GP Public.R":577-584          params (p, theta1, theta2), both Matern(nu)
FAMILY_MATERN_SPLINE1D = 4  # [D2] "1D ... Two Families Public.R":454-462  Matern(nu, theta1) + cubic spline(theta2)
MATERN_NU = 5.0          # [D1]:1080, [D2]:1027


# --------------------------------------------------------------------------
# L1: correlation assembly (expanded-square form, exactly as the R code does)
# --------------------------------------------------------------------------
def corr_matrix(X, theta):
    """[A]:351-360 `corr.matrix(X, theta1, theta2)` generalised to a d-vector
    theta ([H]:328-337 has the same general-vector form).
    Dist = U + t(U) + V with U_ij = sum_k x_ik^2 theta_k, V = -2 X Theta X'."""
    X = np.asarray(X, dtype=np.float64)
    theta = np.asarray(theta, dtype=np.float64).reshape(-1)
    n, d = X.shape
    Theta = np.diag(theta)
    u = ((X ** 2) @ Theta).sum(axis=1)          # apply(X^2 %*% Theta, 1, sum)
    U = np.repeat(u[:, None], n, axis=1)        # matrix(u, n, n, byrow=F)
    V = -2.0 * (X @ Theta) @ X.T
    Dist = U + U.T + V
    return np.exp(-Dist)


def corr_matrix_ISO(X, theta):
    """[I]:350-359 `corr.matrix.ISO(X, theta)`: Theta = theta * I."""
    X = np.asarray(X, dtype=np.float64)
    return corr_matrix(X, np.full(X.shape[1], float(theta)))


def corr_vec(x, X, theta):
    """[A]:369-377 `corr.vec`: exp(-(theta'x^2 - 2 X Theta x + rowsum(X^2 Theta)))."""
    X = np.asarray(X, dtype=np.float64)
    x = np.asarray(x, dtype=np.float64).reshape(-1)
    theta = np.asarray(theta, dtype=np.float64).reshape(-1)
    Theta = np.diag(theta)
    t0 = float(theta @ (x ** 2))
    return np.exp(-(t0 - 2.0 * (X @ Theta) @ x + ((X ** 2) @ Theta).sum(axis=1)))


def corr_vec_ISO(x, X, theta):
    """[I]:369-377."""
    X = np.asarray(X, dtype=np.float64)
    return corr_vec(x, X, np.full(X.shape[1], float(theta)))


def Matern_corr_func(nu, h, theta):
    """[D1]:348-351: (2 sqrt(nu)|h|/theta)^nu K_nu(.) / (Gamma(nu) 2^(nu-1)), 1 at h == 0 (vectorised)."""
    from scipy.special import kv, gamma
    h = np.abs(np.asarray(h, dtype=np.float64))
    t = 2.0 * math.sqrt(nu) * h / theta
    with np.errstate(invalid="ignore", divide="ignore", over="ignore"):
        v = t ** nu * kv(nu, t) / (gamma(nu) * 2.0 ** (nu - 1.0))
    return np.where(h == 0, 1.0, v)


def spline_corr_func(theta, h):
    """[D2]:346-357: nonnegative cubic spline with support theta (vectorised)."""
    u = np.abs(np.asarray(h, dtype=np.float64)) / theta
    return np.where(u <= 0.5, 1 - 6 * u ** 2 + 6 * u ** 3, np.where(u <= 1.0, 2 * (1 - u) ** 3, 0.0))


def _mixed_1d(U, family, params, nu, normalise=True):
    p, t1, t2 = np.asarray(params, dtype=np.float64).reshape(-1)[:3]
    R1 = Matern_corr_func(nu, U, t1)
    R2 = Matern_corr_func(nu, U, t2) if family == FAMILY_MATERN1D else spline_corr_func(t2, U)
    out = p ** 2 * R1 + (1 - p) ** 2 * R2
    return out / (p ** 2 + (1 - p) ** 2) if normalise else out


def component_scales(family, params, d):
    """Map one natural-scale parameter row to (p, theta_comp1[d], theta_comp2[d]).

    ISO          (p, theta1, theta2)            [I]:400-407
    ANISO_LAMBDA (p, theta_1..theta_d, lambda)  [A]:399-406 -> comp2 = (1+lambda)*theta
    ISO_RAW2     (p, theta1, lambda)            [V]:414-421 -> comp2 scale = lambda
    """
    params = np.asarray(params, dtype=np.float64).reshape(-1)
    p = params[0]
    if family in (FAMILY_ISO, FAMILY_ISO_RAW2):
        return p, np.full(d, params[1]), np.full(d, params[2])
    if family == FAMILY_ANISO_LAMBDA:
        th = params[1:1 + d]
        lam = params[1 + d]
        return p, th.copy(), (1.0 + lam) * th
    raise ValueError("unknown family %r" % (family,))


def Mixed_corr_matrix(D, family, params):
    """[A]:399-406 / [I]:400-407 / [V]:414-421:
    R = (p^2 R1 + (1-p)^2 R2) / (p^2 + (1-p)^2)."""
    D = np.asarray(D, dtype=np.float64)
    if family in (FAMILY_MATERN1D, FAMILY_MATERN_SPLINE1D):
        # [D1]:368-374 corr.matrix / [D2]:454-462 corr.matrix.combined: U = |A - t(A)|
        x = D.reshape(-1)
        return _mixed_1d(np.abs(x[None, :] - x[:, None]), family, params, MATERN_NU)
    p, t1, t2 = component_scales(family, params, D.shape[1])
    R1 = corr_matrix(D, t1)
    R2 = corr_matrix(D, t2)
    return (p ** 2 * R1 + (1 - p) ** 2 * R2) / (p ** 2 + (1 - p) ** 2)


def Mixed_corr_vec(x_new, D, family, params):
    """[A]:416-422 / [I]:417-423."""
    D = np.asarray(D, dtype=np.float64)
    if family in (FAMILY_MATERN1D, FAMILY_MATERN_SPLINE1D):
        # [D1]:383-389,... Mixed.corr.vec; [D2]:472-480 corr.vec.combined RETURNS BEFORE DIVIDING (quirk Q3)
        U = np.abs(float(np.asarray(x_new).reshape(-1)[0]) - D.reshape(-1))
        return _mixed_1d(U, family, params, MATERN_NU, normalise=(family == FAMILY_MATERN1D))
    p, t1, t2 = component_scales(family, params, D.shape[1])
    c1 = corr_vec(x_new, D, t1)
    c2 = corr_vec(x_new, D, t2)
    return (p ** 2 * c1 + (1 - p) ** 2 * c2) / (p ** 2 + (1 - p) ** 2)


def cross_corr_matrix(D_old, D_new, theta):
    """[M]:835-848 iso cross-Gram (n_new x n_old), Dist = U + V + W."""
    D_old = np.asarray(D_old, dtype=np.float64)
    D_new = np.asarray(D_new, dtype=np.float64)
    n1, d = D_new.shape
    n2 = D_old.shape[0]
    Theta = np.diag(np.full(d, float(theta)))
    U = np.repeat(((D_new ** 2) @ Theta).sum(axis=1)[:, None], n2, axis=1)
    V = -2.0 * (D_new @ Theta) @ D_old.T
    W = np.repeat(((D_old ** 2) @ Theta).sum(axis=1)[None, :], n1, axis=0)
    return np.exp(-(U + V + W))


# --------------------------------------------------------------------------
# base-R linear algebra, restated on the same LAPACK routines
# --------------------------------------------------------------------------
def r_solve(A, tol=EPS):
    """base R `solve(A)`: dgesv(A, I) then dgecon('1'); error if rcond < tol
    ([A]:448 wraps it in try() -> NA).  Returns None where R would error."""
    A = np.asarray(A, dtype=np.float64)
    n = A.shape[0]
    if not np.all(np.isfinite(A)):
        return None
    anorm = np.abs(A).sum(axis=0).max()
    lu, piv, x, info = lapack.dgesv(A, np.eye(n))
    if info != 0:
        return None
    rcond, info2 = lapack.dgecon(lu, anorm, norm="1")
    if info2 != 0 or rcond < tol:
        return None
    return x


def r_det(A):
    """base R `det`: dgetrf, sign * exp(sum(log(abs(u_ii)))) ([M]:860,874)."""
    A = np.asarray(A, dtype=np.float64)
    lu, piv, info = lapack.dgetrf(A)
    if info < 0:
        raise ValueError("dgetrf")
    dg = np.diag(lu)
    if info > 0 or np.any(dg == 0.0):
        return 0.0
    sign = 1.0
    for i, pv in enumerate(piv):
        if pv != i:
            sign = -sign
    sign *= np.prod(np.sign(dg))
    return float(sign * math.exp(np.log(np.abs(dg)).sum()))


def r_logabsdet(A):
    lu, piv, info = lapack.dgetrf(np.asarray(A, dtype=np.float64))
    return float(np.log(np.abs(np.diag(lu))).sum())


def beta_MLE(R_inv, y):
    """[A]:385-389: 1'R^-1 y / sum(R^-1)."""
    y = np.asarray(y, dtype=np.float64).reshape(-1)
    return float(np.ones(len(y)) @ R_inv @ y / R_inv.sum())


def dmnorm_pdsolve(y, mean, varcov):
    """mnormt::dmnorm(log=TRUE), mnormt >= 1.5: pd.solve = symmetrise, chol
    (dpotrf 'U'), chol2inv (dpotri), log.det = 2*sum(log(diag(U))).
    Returns (logpdf, symmetric_ok) -- pd.solve *stops* when
    max|x - t(x)| > .Machine$double.eps; we report that instead of raising."""
    y = np.asarray(y, dtype=np.float64).reshape(-1)
    d = len(y)
    Xc = y - mean
    sym_ok = bool(np.max(np.abs(varcov - varcov.T)) <= EPS)
    S = (varcov + varcov.T) / 2.0
    u, info = lapack.dpotrf(S, lower=0)
    if info != 0:
        return float("nan"), sym_ok
    inv, info = lapack.dpotri(u, lower=0)
    inv = np.triu(inv) + np.triu(inv, 1).T
    log_det = 2.0 * np.log(np.diag(u)).sum()
    Q = float((inv @ Xc) @ Xc)
    return -(Q + d * LOG2PI + log_det) / 2.0, sym_ok


def dmnorm_legacy(y, mean, varcov):
    """mnormt 1.4 (paper era): Q via solve(varcov), logDet = sum(log|diag(qr)|)."""
    y = np.asarray(y, dtype=np.float64).reshape(-1)
    d = len(y)
    Xc = y - mean
    inv = r_solve(varcov)
    if inv is None:
        return float("nan")
    Q = float((inv @ Xc) @ Xc)
    qr_r = np.linalg.qr(varcov, mode="r")
    logdet = float(np.log(np.abs(np.diag(qr_r))).sum())
    return -(Q + d * LOG2PI + logdet) / 2.0


# --------------------------------------------------------------------------
# L2: per-candidate numerical core
# --------------------------------------------------------------------------
def loglik_reference(D, y, sigma2, family, params, legacy_dmnorm=False):
    """Reference-faithful log-likelihood of one candidate, i.e. the body of
    `logpost` up to `log.like` ([A]:444-455, [I]:441-451, [V]:456-465):
    R -> solve(R) -> beta.MLE -> dmnorm(y, beta, (p^2+(1-p)^2) sigma2 R, log=1).
    Returns dict(loglik, beta, R_inv, status); status 2 where R would give NA."""
    y = np.asarray(y, dtype=np.float64).reshape(-1)
    p = float(np.asarray(params).reshape(-1)[0])
    R = Mixed_corr_matrix(D, family, params)
    R_inv = r_solve(R)
    if R_inv is None:
        return dict(loglik=float("nan"), beta=float("nan"), R_inv=None, status=2)
    beta = beta_MLE(R_inv, y)
    varcov = (p ** 2 + (1 - p) ** 2) * sigma2 * R
    if legacy_dmnorm:
        ll = dmnorm_legacy(y, beta, varcov)
    else:
        ll, _ = dmnorm_pdsolve(y, beta, varcov)
    st = 0 if math.isfinite(ll) else 1
    return dict(loglik=ll, beta=beta, R_inv=R_inv, status=st)


def transform_theta(family, theta, d):
    """logpost's re-parametrisation ([A]:435-442, [I]:435-440, [V]:449-454):
    real-line vector (psi_1.., phi[, zeta]) -> natural (p, theta.., [lambda])
    ISO / ISO_RAW2: theta = (psi1, psi2, phi);  ANISO: (psi_1..psi_d, phi, zeta)."""
    theta = np.asarray(theta, dtype=np.float64).reshape(-1)
    if family != FAMILY_ANISO_LAMBDA:
        psi1, psi2, phi = theta[:3]
        return np.array([1.0 / (1.0 + math.exp(-phi)), math.exp(psi1), math.exp(psi2)])
    psi = theta[:d]
    phi = theta[d]
    zeta = theta[d + 1]
    return np.concatenate([[1.0 / (1.0 + math.exp(-phi))], np.exp(psi), [math.exp(zeta)]])


def log_jacobian(family, theta, d):
    """[A]:459 / [I]:452: -phi - 2 log(1+e^-phi) + sum(psi) [+ zeta]."""
    theta = np.asarray(theta, dtype=np.float64).reshape(-1)
    if family != FAMILY_ANISO_LAMBDA:
        psi1, psi2, phi = theta[:3]
        return -phi - 2.0 * math.log(1.0 + math.exp(-phi)) + psi1 + psi2
    psi = theta[:d]
    phi = theta[d]
    zeta = theta[d + 1]
    return -phi - 2.0 * math.log(1.0 + math.exp(-phi)) + float(psi.sum()) + zeta


def log_prior(script, theta, prior_pars=None):
    """The per-script hard-coded log-prior: [A]:462, [I]:453, [G]:450, or the
    parametrised inverse-gamma form of [V]:467 / [H]:462."""
    theta = np.asarray(theta, dtype=np.float64).reshape(-1)
    if script == "A":
        psi1, psi2, phi, zeta = theta[:4]
        return -psi1 - psi1 ** 2 / 2 - psi2 - psi2 ** 2 / 2 - 4 * zeta - 4 / math.exp(zeta)
    psi1, psi2 = theta[0], theta[1]
    t1, t2 = math.exp(psi1), math.exp(psi2)
    if script in ("I", "D1", "D2"):      # [I]:453 = [D1]:636 = [D2]:597
        return -4 * psi1 - 2 / t1 - 6 * psi2 - 16 / t2
    if script == "G":
        return -4 * psi1 - 1 / t1 - 6 * psi2 - 75 / t2
    if script in ("V", "H"):
        a1, b1, a2, b2 = prior_pars
        return -(a1 + 1) * psi1 - b1 / t1 - (a2 + 1) * psi2 - b2 / t2
    raise ValueError(script)


def logpost(D, theta, y, sigma2, family, script, prior_pars=None):
    """Full `logpost` ([A]:433-467 etc.): list(val, beta, R.Inv)."""
    D = np.asarray(D, dtype=np.float64)
    nat = transform_theta(family, theta, D.shape[1])
    r = loglik_reference(D, y, sigma2, family, nat)
    val = r["loglik"] + log_jacobian(family, theta, D.shape[1]) + log_prior(script, theta, prior_pars)
    return dict(val=val, beta=r["beta"], R_inv=r["R_inv"], loglik=r["loglik"], status=r["status"])


def loglik_minimal(D, y, sigma2, family, params, mean_mode="gls", tau=0.0):
    """Minimal algorithm = what the CUDA kernel computes:
    direct-difference Gram, Cholesky, two forward solves, dots, log-det."""
    D = np.asarray(D, dtype=np.float64)
    y = np.asarray(y, dtype=np.float64).reshape(-1)
    n, d = D.shape
    p = float(np.asarray(params).reshape(-1)[0])
    w = p ** 2 + (1 - p) ** 2
    R = Mixed_corr_matrix_direct(D, family, params)
    L, info = lapack.dpotrf(R, lower=1)
    if info != 0:
        return dict(loglik=float("nan"), beta=float("nan"), status=1)
    L = np.tril(L)
    zy = lapack.dtrtrs(L, y, lower=1)[0]
    z1 = lapack.dtrtrs(L, np.ones(n), lower=1)[0]
    s11 = float(z1 @ z1)
    s1y = float(z1 @ zy)
    c = w * sigma2
    logdetR = 2.0 * float(np.log(np.diag(L)).sum())
    beta = s1y / s11
    resid = zy - beta * z1
    QR = float(resid @ resid)
    if mean_mode == "gls":
        ll = -0.5 * (QR / c + n * LOG2PI + n * math.log(c) + logdetR)
        return dict(loglik=ll, beta=beta, status=0)
    g = 1.0 + tau ** 2 * s11 / c
    quad = QR / c + s1y ** 2 / (c * s11 * g)
    ll = -0.5 * (quad + n * LOG2PI + n * math.log(c) + logdetR + math.log(g))
    return dict(loglik=ll, beta=beta, status=0)


def cond_loglike_reference(D, y, sigma2, family, params, tau):
    """`cond.like` before the exp() ([V]:564-575, [H]:561-572):
    dmnorm(y, 0, sigma2 (p^2+(1-p)^2) R + tau^2 11', log=1)."""
    y = np.asarray(y, dtype=np.float64).reshape(-1)
    n = len(y)
    p = float(np.asarray(params).reshape(-1)[0])
    R = Mixed_corr_matrix(D, family, params)
    varcov = sigma2 * (p ** 2 + (1 - p) ** 2) * R + tau ** 2 * np.ones((n, n))
    ll, _ = dmnorm_pdsolve(y, 0.0, varcov)
    return ll


def halton_base2(N):
    """fOptions::runif.halton(N, 1): van der Corput radical inverse, base 2,
    indices 1..N ([V]:557)."""
    out = np.empty(N)
    for i in range(1, N + 1):
        f, r, k = 0.5, 0.0, i
        while k > 0:
            r += f * (k & 1)
            k >>= 1
            f *= 0.5
        out[i - 1] = r
    return out


def qigamma(p, alpha, beta):
    """pscl::qigamma(p, alpha, beta) = 1 / qgamma(1 - p, alpha, rate = beta)."""
    from scipy.stats import gamma
    return 1.0 / gamma.ppf(1.0 - np.asarray(p), a=alpha, scale=1.0 / beta)


def sweep_candidates(theta1_pars, theta2_pars, N):
    """[V]:557-560: one Halton stream u -> (p, theta1, theta2) = (u, qIG(u;a1,b1), qIG(u;a2,b2))."""
    u = halton_base2(N)
    return np.column_stack([u, qigamma(u, *theta1_pars), qigamma(u, *theta2_pars)])


def likeli_hyperpars(D, y, theta1_pars, theta2_pars, sigma2, N=1728, tau=100.0, family=FAMILY_ISO):
    """[V]:552-578 (N=1728, tau=100) / [H]:549-575 (N=1000, tau=50): mean of exp(cond.like)."""
    pars = sweep_candidates(theta1_pars, theta2_pars, N)
    ll = np.array([cond_loglike_reference(D, y, sigma2, family, row, tau) for row in pars])
    return float(np.mean(np.exp(ll)))


def choose_hyperpars(D, y, hyperpars_matrix, sigma2, N=1728, tau=100.0, family=FAMILY_ISO):
    """[V]:588-599: which.max over the rows of the hyper-prior grid (first max)."""
    likes = np.array([likeli_hyperpars(D, y, h[0:2], h[2:4], sigma2, N, tau, family)
                      for h in np.asarray(hyperpars_matrix)])
    return dict(pars=np.asarray(hyperpars_matrix)[int(np.argmax(likes))], likelihoods=likes)


def factors(R_inv, beta, y):
    """[A]:550-559: mean.factor = R^-1 (y - beta 1), var.factor1 = colSums(R^-1),
    var.factor2 = sum(var.factor1)."""
    y = np.asarray(y, dtype=np.float64).reshape(-1)
    mean_factor = R_inv @ (y - beta)
    var_factor1 = R_inv.sum(axis=0)
    return mean_factor, var_factor1, float(var_factor1.sum())


def predict_post(x_new, D, family_vec, params_vec, beta, mean_factor, var_factor1, var_factor2, R_inv, sigma2):
    """[A]:604-623: mean = beta + mean.factor'r;
    var = sigma2 (1 - r'R^-1 r + (1 - var.factor1'r)^2 / var.factor2)   (quirk Q1: sigma2, not c)."""
    r = Mixed_corr_vec(x_new, D, family_vec, params_vec)
    var = sigma2 * (1.0 - r @ R_inv @ r + (1.0 - var_factor1 @ r) ** 2 / var_factor2)
    mean = beta + mean_factor @ r
    return float(mean), float(var)


def predict_table(D, y, sigma2, family, pars, X_new, family_vec=None, pars_vec=None):
    """`prediction`'s inner table ([A]:639): for S posterior rows x T sites,
    (mean, var) from predict.post with factors() recomputed from logpost's
    R.Inv/beta.  family_vec/pars_vec allow quirk Q2 ([V]:672: the correlation
    *vector* uses theta1*(1+lambda) although the matrix used `lambda`)."""
    pars = np.atleast_2d(pars)
    X_new = np.atleast_2d(X_new)
    S, T = pars.shape[0], X_new.shape[0]
    mean = np.full((T, S), np.nan)
    var = np.full((T, S), np.nan)
    for s in range(S):
        r = loglik_reference(D, y, sigma2, family, pars[s])
        if r["R_inv"] is None:
            continue
        mf, vf1, vf2 = factors(r["R_inv"], r["beta"], y)
        fv = family if family_vec is None else family_vec
        pv = pars[s] if pars_vec is None else np.atleast_2d(pars_vec)[s]
        for t in range(T):
            mean[t, s], var[t, s] = predict_post(X_new[t], D, fv, pv, r["beta"], mf, vf1, vf2, r["R_inv"], sigma2)
    return mean, var


def Entropy(D, p, theta1, theta2):
    """[M]:856-861: -det(Mixed.corr.matrix(D, p, theta1, theta2)) (iso)."""
    return -r_det(Mixed_corr_matrix(D, FAMILY_ISO, [p, theta1, theta2]))


def mixed_R_old_inv(D_old, p, theta1, theta2):
    """[M]:924-925: R.old and solve(R.old, tol=1e-16)."""
    R_old = (p ** 2 * corr_matrix_ISO(D_old, theta1) + (1 - p) ** 2 * corr_matrix_ISO(D_old, theta2)) / (p ** 2 + (1 - p) ** 2)
    return r_solve(R_old, tol=1e-16)


def Augmented_Mixed_Entropy(D_old, D_new, p, theta1, theta2, R_old_inv):
    """[M]:869-877: -det(R.new - R.cross R.old.Inv R.cross')."""
    w = p ** 2 + (1 - p) ** 2
    R_cross = (p ** 2 * cross_corr_matrix(D_old, D_new, theta1) + (1 - p) ** 2 * cross_corr_matrix(D_old, D_new, theta2)) / w
    R_new = Mixed_corr_matrix(D_new, FAMILY_ISO, [p, theta1, theta2])
    return -r_det(R_new - R_cross @ R_old_inv @ R_cross.T)


def me_schur_negdet_batch(D_old, D_new_pool, params):
    """Oracle for the batched ME criterion: out[c, q] = Augmented.Mixed.Entropy
    of pool design c under parameter row q (R.old.Inv recomputed per row as in
    Batch.Entropy.optim [M]:924-925).  which.min over c = first index of min."""
    D_new_pool = np.asarray(D_new_pool, dtype=np.float64)
    params = np.atleast_2d(params)
    out = np.empty((D_new_pool.shape[0], params.shape[0]))
    for q, (p, t1, t2) in enumerate(params):
        Rinv = mixed_R_old_inv(D_old, p, t1, t2)
        for c in range(D_new_pool.shape[0]):
            out[c, q] = Augmented_Mixed_Entropy(D_old, D_new_pool[c], p, t1, t2, Rinv)
    return out


def subset_logdet(pool, idx, family, params):
    """log det R[S,S] for one index subset (SURVEY 8d ME-B(2)); minimal path."""
    sub = np.asarray(pool)[np.asarray(idx)]
    R = Mixed_corr_matrix_direct(sub, family, params)
    L, info = lapack.dpotrf(R, lower=1)
    if info != 0:
        return float("nan")
    return 2.0 * float(np.log(np.diag(L)).sum())


def Mixed_corr_matrix_direct(D, family, params):
    """Direct-difference form of the mixed Gram (exact symmetry, unit diagonal)."""
    D = np.asarray(D, dtype=np.float64)
    if family in (FAMILY_MATERN1D, FAMILY_MATERN_SPLINE1D):
        return Mixed_corr_matrix(D, family, params)
    p, t1, t2 = component_scales(family, params, D.shape[1])
    w = p ** 2 + (1 - p) ** 2
    diff = D[:, None, :] - D[None, :, :]
    s1 = (diff ** 2 * t1).sum(axis=2)
    s2 = (diff ** 2 * t2).sum(axis=2)
    R = (p ** 2 / w) * np.exp(-s1) + ((1 - p) ** 2 / w) * np.exp(-s2)
    np.fill_diagonal(R, 1.0)
    return R


# --------------------------------------------------------------------------
# extended-precision truth (mpmath, 50 digits) for the conditioning study
# --------------------------------------------------------------------------
def loglik_truth(D, y, sigma2, family, params, mean_mode="gls", tau=0.0, dps=50):
    """50-digit evaluation of the same likelihood (GLS-beta or zero-mean+tau^2)."""
    import mpmath as mp
    mp.mp.dps = dps
    D = np.asarray(D, dtype=np.float64)
    y = np.asarray(y, dtype=np.float64).reshape(-1)
    n, d = D.shape
    p, t1, t2 = component_scales(family, params, d)
    p = mp.mpf(float(p))
    w = p ** 2 + (1 - p) ** 2
    R = mp.matrix(n, n)
    for i in range(n):
        for j in range(i + 1):
            s1 = mp.mpf(0)
            s2 = mp.mpf(0)
            for k in range(d):
                df = mp.mpf(float(D[i, k])) - mp.mpf(float(D[j, k]))
                s1 += mp.mpf(float(t1[k])) * df * df
                s2 += mp.mpf(float(t2[k])) * df * df
            v = (p ** 2 * mp.e ** (-s1) + (1 - p) ** 2 * mp.e ** (-s2)) / w
            R[i, j] = v
            R[j, i] = v
    c = w * mp.mpf(float(sigma2))
    yv = mp.matrix([mp.mpf(float(v)) for v in y])
    one = mp.matrix([mp.mpf(1)] * n)
    L = mp.cholesky(R)
    zy = mp.lu_solve(L, yv)
    z1 = mp.lu_solve(L, one)
    s11 = sum(z1[i] ** 2 for i in range(n))
    s1y = sum(z1[i] * zy[i] for i in range(n))
    syy = sum(zy[i] ** 2 for i in range(n))
    logdetR = 2 * sum(mp.log(L[i, i]) for i in range(n))
    log2pi = mp.log(2 * mp.pi)
    if mean_mode == "gls":
        beta = s1y / s11
        QR = syy - s1y ** 2 / s11
        ll = -(QR / c + n * log2pi + n * mp.log(c) + logdetR) / 2
        return float(ll), float(beta)
    g = 1 + mp.mpf(float(tau)) ** 2 * s11 / c
    quad = syy / c - (mp.mpf(float(tau)) ** 2 / c ** 2) * s1y ** 2 / g
    ll = -(quad + n * log2pi + n * mp.log(c) + logdetR + mp.log(g)) / 2
    return float(ll), float(s1y / s11)


def cond1(R):
    """1-norm condition number estimate via dgecon (what base R's rcond uses)."""
    anorm = np.abs(R).sum(axis=0).max()
    lu, piv, info = lapack.dgetrf(R)
    rcond, _ = lapack.dgecon(lu, anorm, norm="1")
    return 1.0 / rcond if rcond > 0 else float("inf")


# --------------------------------------------------------------------------
# harness helpers (the simulators of the scripts, restated for synthetic y)
# --------------------------------------------------------------------------
def test_function(code, x, y):
    """[A]:330-341 bivariate simulators; code 4 is the script default ([A]:848)."""
    if code == 1:
        return np.exp(-1.4 * x) * np.cos(7 * np.pi * x * y / 2) + np.log(x + y + 0.1)
    if code == 2:
        return ((x - 0.2) ** 2 - (y - 0.7) ** 2) * np.exp(-5 * ((x - 0.8) ** 2 + (y - 0.1) ** 2)) * np.cos(10 * (x - 0.5) * y)
    if code == 3:
        return ((x - 0.5) ** 2 + 4 * (y - 0.8) ** 2) * (np.cos(np.pi * (x - 0.1)) + np.cos(np.pi * (y - 0.5)))
    if code == 4:
        return (np.sin(2 * x) + np.cos(4 * x)) * (np.sin(8 * y) + np.cos(4 * y))
    if code == 5:
        return np.sin(9 * x - 4.5) / (9 * x - 4.5) * np.sin(12 * y - 6) / (12 * y - 6)
    raise ValueError(code)


# ---------------------------------------------------------------------------------------------
# k-medoids (PAM) -- the step between All_Subdesigns.txt and `k-medoids ME Design.txt`
# (reference ReadMe.md:54-60; the script that ran it is not shipped).  R's cluster::pam is an
# un-vendored CRAN dependency with no pinned version; this restates its published algorithm
# (Kaufman & Rousseeuw 1990, ch. 2: BUILD, then SWAP by steepest descent) with Euclidean
# distances.  Pinned by the shipped files: on the 7000 points of All_Subdesigns.txt it ends at
# exactly rows 15-21 of `k-medoids ME Design.txt` (tests/test_kmedoids.py).
def pam_kmedoids(P, k, max_swaps=1000, trace=None):
    P = np.asarray(P, dtype=np.float64)
    n = P.shape[0]
    D = np.zeros((n, n))
    for kk in range(P.shape[1]):                       # sum over coordinates in order, no FMA (as the device)
        df = P[:, kk][:, None] - P[:, kk][None, :]
        D += df * df
    np.sqrt(D, out=D)
    med = [int(np.argmin(D.sum(axis=0)))]
    near = D[:, med[0]].copy()
    while len(med) < k:
        gain = np.zeros(n)
        for lo in range(0, n, 1024):                   # blocked: bounds the temporary
            gain[lo:lo + 1024] = np.maximum(near[:, None] - D[:, lo:lo + 1024], 0.0).sum(axis=0)
        gain[med] = -np.inf
        h = int(np.argmax(gain))
        med.append(h)
        near = np.minimum(near, D[:, h])
    if trace is not None:
        trace.append(("build", list(med)))
    swaps = 0
    while swaps < max_swaps and k < n:
        Dm = D[:, med]
        order = np.argsort(Dm, axis=1, kind="stable")
        nearest = order[:, 0]
        d1 = Dm[np.arange(n), nearest]
        d2 = Dm[np.arange(n), order[:, 1]] if k > 1 else np.full(n, np.inf)
        cur = d1.sum()
        best = (0.0, None, None)
        for i in range(k):
            base = np.where(nearest == i, d2, d1)
            newcost = np.empty(n)
            for lo in range(0, n, 1024):
                newcost[lo:lo + 1024] = np.minimum(base[:, None], D[:, lo:lo + 1024]).sum(axis=0)
            newcost[med] = np.inf
            h = int(np.argmin(newcost))
            delta = newcost[h] - cur
            if delta < best[0] and delta < -1e-12 * cur:
                best = (delta, i, h)
        if best[1] is None:
            break
        med[best[1]] = best[2]
        swaps += 1
        if trace is not None:
            trace.append(("swap", best[1], best[2]))
    cost = D[:, med].min(axis=1).sum()
    return np.array(med, dtype=np.int32), float(cost), swaps


# ------------------------------------------------------------------------------------------------------------
# CGP comparator (Ba & Joseph's composite GP as re-stated in every script): the objective of
# its 505-candidate start sweep and the leave-one-out loop.  Literal restatement of [A]:93-200
# ("2D Combined GP Anisotropic Public.R"; the same text sits in all eight scripts).
def cgp_standardise(X):
    """[A]:70-72: Stand_DD = (x - min) / (max - min) per column; scales = max - min."""
    X = np.atleast_2d(np.asarray(X, dtype=np.float64))
    lo = X.min(axis=0)
    scales = X.max(axis=0) - lo
    return (X - lo) / scales, scales


def cgp_psi(Xs, theta):
    """PSI / Stand_PSI [A]:93-104: exp(-dist(X diag(sqrt(theta)))^2) -- `dist` is the Euclidean distance, squared again."""
    A = Xs * np.sqrt(np.asarray(theta, dtype=np.float64))[None, :]
    D = np.sqrt(((A[:, None, :] - A[None, :, :]) ** 2).sum(axis=2))
    return np.exp(-D ** 2)


def cgp_bounds(Xs, nugget_l=0.001, theta_l=1e-4):
    """[A]:79-92: lower / upper box of (lambda, theta_1..theta_p, kappa, bw)."""
    n, p = Xs.shape
    iu = np.triu_indices(n, 1)
    d = np.sqrt(((Xs[:, None, :] - Xs[None, :, :]) ** 2).sum(axis=2))[iu]
    m = np.mean(1.0 / d ** 2)
    alpha_l = np.log(10.0 ** 2) * m
    kappa_u = np.log(10.0 ** 6) * m
    lower = np.concatenate([[nugget_l], np.full(p, theta_l), [alpha_l, 0.0]])
    upper = np.concatenate([[1.0], np.full(p, alpha_l), [kappa_u, 1.0]])
    return lower, upper


def _cgp_iterate(G, L, Gbw, y, lam, reps=4):
    """The four re-weighting passes shared by var.MLE.DK ([A]:113-124), the jackknife ([A]:172-183) and the final
    fit ([A]:206-217).  Returns (Sig diagonal, Sig2 of the last pass, e of the last pass)."""
    n = len(y)
    one = np.ones(n)
    sig = np.ones(n)
    sig2 = 1.0
    e = np.zeros(n)
    for _ in range(reps):
        Q = G + lam * (np.sqrt(sig)[:, None] * L * np.sqrt(sig)[None, :])
        invQ = r_solve(Q, tol=EPS)
        beta = (one @ invQ @ y) / (one @ invQ @ one)
        temp = invQ @ (y - beta * one)
        gip = beta * one + G @ temp
        e = y - gip
        sig = (Gbw @ e ** 2) / (Gbw @ one)
        sig2 = float(np.mean(sig))
        sig = sig / sig2
    return sig, sig2, e


def cgp_var_mle_dk(Xs, y, ww):
    """var.MLE.DK(ww) [A]:104-135 on the standardised design: log(det(Q)) + n log(tau2), 1e6 when not finite
    (det underflows to 0 -> log = -Inf -> 1e6: reproduced by going through det itself)."""
    Xs = np.atleast_2d(Xs)
    y = np.asarray(y, dtype=np.float64)
    n, p = Xs.shape
    ww = np.asarray(ww, dtype=np.float64)
    lam, th, kappa, bw = ww[0], ww[1:p + 1], ww[p + 1], ww[p + 2]
    G, L, Gbw = cgp_psi(Xs, th), cgp_psi(Xs, kappa + th), cgp_psi(Xs, th * bw)
    one = np.ones(n)
    try:
        sig, _, _ = _cgp_iterate(G, L, Gbw, y, lam)
        Q = G + lam * (np.sqrt(sig)[:, None] * L * np.sqrt(sig)[None, :])
        invQ = r_solve(Q, tol=EPS)
        beta = (one @ invQ @ y) / (one @ invQ @ one)
        tau2 = (y - beta * one) @ invQ @ (y - beta * one) / n
        with np.errstate(divide="ignore", invalid="ignore"):
            val = float(np.log(r_det(Q)) + n * np.log(tau2))
    except Exception:  # noqa: BLE001  (solve() refusing a singular Q stops the R function; the sweep never sees it in practice)
        val = float("nan")
    return val if np.isfinite(val) else 1e6


def cgp_jackknife(X, y, lam, theta, alpha, bw):
    """Leave-one-out predictions Yp_jackknife and rmscv ([A]:166-201) on the ORIGINAL coordinates with
    theta = Stand_theta / scales^2, alpha = Stand_alpha / scales^2 ([A]:164-165)."""
    X = np.atleast_2d(np.asarray(X, dtype=np.float64))
    y = np.asarray(y, dtype=np.float64)
    n = X.shape[0]
    Gf, Lf, Gbwf = cgp_psi(X, theta), cgp_psi(X, alpha), cgp_psi(X, np.asarray(theta) * bw)
    out = np.zeros(n)
    for jf in range(n):
        keep = np.arange(n) != jf
        G, L, Gbw, yk = Gf[np.ix_(keep, keep)], Lf[np.ix_(keep, keep)], Gbwf[np.ix_(keep, keep)], y[keep]
        onem = np.ones(n - 1)
        sig, sig2, e = _cgp_iterate(G, L, Gbw, yk, lam)
        Q = G + lam * (np.sqrt(sig)[:, None] * L * np.sqrt(sig)[None, :])
        invQ = r_solve(Q, tol=EPS)
        beta = (onem @ invQ @ yk) / (onem @ invQ @ onem)
        temp = invQ @ (yk - beta * onem)
        g, l, gbw = Gf[jf, keep], Lf[jf, keep], Gbwf[jf, keep]
        vjf = ((gbw @ e ** 2) / (gbw @ onem)) / sig2
        q = g + lam * np.sqrt(vjf) * (np.sqrt(sig) * l)
        out[jf] = beta + q @ temp
    return out, float(np.sqrt(np.sum((y - out) ** 2) / n))
