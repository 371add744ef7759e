"""Lock-step drivers for the reference's sequential callers of the hot path.

`Metro` ([A]:484-539) evaluates ONE `logpost` per proposal, `LearnBayes::laplace` ([A]:494) one per
Nelder-Mead vertex: the GPU is idle between calls.  Here C independent chains / simplices advance
in lock-step and every step is ONE batched `logpost` over all of them, so the kernel's throughput
becomes end-to-end fit throughput (BASELINE config 3: "multi-start hyperparameter optimisation
across training/test sets").

Everything here is host-side control flow.  The numerics come from the callable
`logpost_fn(theta[B, k]) -> dict(val[B], beta[B], ...)` -- in production
`reference_api.logpost_batch` (CUDA), in the CPU tests the oracle.

Reference behaviour kept, with the quirks named:
  * proposal `rmnorm(1, theta.old, sqrt(2) * v)`: the third argument of mnormt::rmnorm is the
    COVARIANCE, so the proposal covariance is sqrt(2) v ([A]:512);
  * accept iff `l.cand - l.old > log(u)` with u drawn BEFORE the proposal ([A]:511-516); a candidate
    whose likelihood is NA makes R's `if` fail -- here it is rejected and counted (`n_na`);
  * the Geweke test runs on `samp[(k-samp.size):(k-1)]`: a matrix indexed by ONE vector, i.e. on the
    FIRST parameter's column only ([A]:530); p-value = 2(1 - pnorm(|z|));
  * the chain stops when pv >= alpha (checked every batch.size accepts once samp.size are stored) or
    after N accepted samples; the last samp.size samples are returned ([A]:538).
"""
from __future__ import annotations

import math
import numpy as np


# ------------------------------------------------------------------ coda::geweke.diag
def _ar_yule_walker_aic(x):
    """stats::ar(x, aic=TRUE, method='yule-walker'): -> (coefficients, var.pred).
    order.max = min(n-1, floor(10 log10 n)); AIC_k = n log(v_k) + 2k; var.pred scaled by n/(n-(order+1))."""
    x = np.asarray(x, dtype=np.float64)
    n = x.size
    order_max = int(min(n - 1, math.floor(10.0 * math.log10(n))))
    xc = x - x.mean()
    r = np.array([np.dot(xc[: n - k], xc[k:]) / n for k in range(order_max + 1)])
    if not (r[0] > 0):
        raise FloatingPointError("zero-variance series")
    # Levinson-Durbin: coefs[k] = AR(k) coefficients, v[k] = innovations variance
    v = np.empty(order_max + 1)
    v[0] = r[0]
    coefs = [np.zeros(0)]
    phi = np.zeros(0)
    for k in range(1, order_max + 1):
        acc = r[k] - np.dot(phi, r[k - 1:0:-1]) if k > 1 else r[1]
        refl = acc / v[k - 1]
        phi = np.concatenate([phi - refl * phi[::-1], [refl]])
        v[k] = v[k - 1] * (1.0 - refl * refl)
        coefs.append(phi.copy())
    aic = n * np.log(v) + 2.0 * np.arange(order_max + 1)
    order = int(np.argmin(aic))
    var_pred = v[order] * n / (n - (order + 1))
    return coefs[order], var_pred


def spectrum0_ar(x):
    """coda::spectrum0.ar: spectral density at frequency zero from the AIC-selected AR fit."""
    ar, var_pred = _ar_yule_walker_aic(x)
    return var_pred / (1.0 - ar.sum()) ** 2


def geweke_z(x, frac1=0.1, frac2=0.5):
    """coda::geweke.diag(mcmc(x))$z for one series (windows: first 10 %, last 50 %)."""
    x = np.asarray(x, dtype=np.float64)
    n = x.size
    end1 = int(math.ceil(1 + frac1 * (n - 1)))          # 1-based, inclusive
    start2 = int(math.floor(n - frac2 * (n - 1)))
    a, b = x[:end1], x[start2 - 1:]
    va, vb = spectrum0_ar(a), spectrum0_ar(b)
    return (a.mean() - b.mean()) / math.sqrt(va / a.size + vb / b.size)


def geweke_pvalue(x):
    """`min(2*(1-pnorm(abs(geweke.diag(mcmc(x))$z))))` with the reference's try() -> 0 ([A]:530-532)."""
    try:
        z = geweke_z(x)
    except (FloatingPointError, ZeroDivisionError, ValueError):
        return 0.0
    if not np.isfinite(z):
        return 0.0
    return float(math.erfc(abs(z) / math.sqrt(2.0)))     # = 2 (1 - pnorm(|z|))


# ------------------------------------------------------------------ LearnBayes::laplace, batched
def _neg_inv(H):
    return -np.linalg.inv(H)


def laplace_batch(logpost_fn, starts, max_iter=500, reltol=1.4901161193847656e-08, hess_step=1e-3):
    """`LearnBayes::laplace(logpost, start)` for C starts at once: Nelder-Mead (optim's constants:
    reflection 1, contraction 0.5, expansion 2; initial simplex step 0.1 * max|start| as in
    optim's nmmin; stop when f_high - f_low <= reltol (|f_low| + reltol)) on -logpost, every simplex
    advanced in lock-step: ONE batched call evaluates the reflection, expansion and both
    contraction points of all simplices (4C rows), shrinks add one more call.  The Hessian at
    each mode is the central-difference stencil of optimHess (ndeps = hess_step), again one call.
    -> dict(mode[C,k], var[C,k,k] = -H^-1, int[C], converge[C], evals)."""
    starts = np.atleast_2d(np.asarray(starts, dtype=np.float64))
    C, k = starts.shape
    evals = 0

    def f(th):                                   # minimise -logpost; NA -> +inf (never chosen)
        nonlocal evals
        evals += th.shape[0]
        v = -np.asarray(logpost_fn(th)["val"], dtype=np.float64)
        return np.where(np.isfinite(v), v, np.inf)

    step = 0.1 * np.maximum(np.abs(starts).max(axis=1), 1e-300)
    step = np.where(np.abs(starts).max(axis=1) == 0, 0.1, step)
    simplex = np.repeat(starts[:, None, :], k + 1, axis=1)          # C x (k+1) x k
    for j in range(k):
        simplex[:, j + 1, j] += step
    fv = f(simplex.reshape(-1, k)).reshape(C, k + 1)
    active = np.ones(C, dtype=bool)
    it = 0
    while active.any() and it < max_iter:
        it += 1
        order = np.argsort(fv, axis=1, kind="stable")
        simplex = np.take_along_axis(simplex, order[:, :, None], axis=1)
        fv = np.take_along_axis(fv, order, axis=1)
        lo, hi = fv[:, 0], fv[:, -1]
        done = np.abs(hi - lo) <= reltol * (np.abs(lo) + reltol)
        active &= ~done
        if not active.any():
            break
        idx = np.nonzero(active)[0]
        S, F = simplex[idx], fv[idx]
        cen = S[:, :-1, :].mean(axis=1)
        worst = S[:, -1, :]
        xr = 2.0 * cen - worst
        xe = 3.0 * cen - 2.0 * worst                                  # cen + 2 (xr - cen)
        xoc = 1.5 * cen - 0.5 * worst                                 # outside contraction
        xic = 0.5 * cen + 0.5 * worst                                 # inside contraction
        fall = f(np.concatenate([xr, xe, xoc, xic])).reshape(4, -1)
        fr, fe, foc, fic = fall
        new_pt = worst.copy()
        new_f = F[:, -1].copy()
        shrink = np.zeros(idx.size, dtype=bool)
        for i in range(idx.size):
            if fr[i] < F[i, 0]:
                if fe[i] < fr[i]:
                    new_pt[i], new_f[i] = xe[i], fe[i]
                else:
                    new_pt[i], new_f[i] = xr[i], fr[i]
            elif fr[i] < F[i, -2]:
                new_pt[i], new_f[i] = xr[i], fr[i]
            elif fr[i] < F[i, -1]:
                if foc[i] <= fr[i]:
                    new_pt[i], new_f[i] = xoc[i], foc[i]
                else:
                    shrink[i] = True
            else:
                if fic[i] < F[i, -1]:
                    new_pt[i], new_f[i] = xic[i], fic[i]
                else:
                    shrink[i] = True
        S[:, -1, :] = new_pt
        F[:, -1] = new_f
        if shrink.any():
            sh = np.nonzero(shrink)[0]
            S[sh, 1:, :] = 0.5 * (S[sh, 1:, :] + S[sh, :1, :])
            F[sh, 1:] = f(S[sh, 1:, :].reshape(-1, k)).reshape(sh.size, k)
        simplex[idx], fv[idx] = S, F
    best = np.argmin(fv, axis=1)
    mode = simplex[np.arange(C), best]
    fmode = fv[np.arange(C), best]
    # optimHess: central differences of the (numerical, central) gradient == 4-point second differences
    h = hess_step
    pts = [mode]
    for a in range(k):
        for b in range(a, k):
            for sa, sb in ((1, 1), (1, -1), (-1, 1), (-1, -1)):
                p = mode.copy()
                p[:, a] += sa * h
                p[:, b] += sb * h
                pts.append(p)
    vals = -f(np.concatenate(pts)).reshape(len(pts), C)              # logpost values
    H = np.zeros((C, k, k))
    q = 1
    for a in range(k):
        for b in range(a, k):
            pp, pm, mp, mm = vals[q], vals[q + 1], vals[q + 2], vals[q + 3]
            q += 4
            H[:, a, b] = H[:, b, a] = (pp - pm - mp + mm) / (4.0 * h * h)
    var = np.stack([_neg_inv(H[c]) for c in range(C)])
    sign, logdet = np.linalg.slogdet(var)
    integral = 0.5 * k * math.log(2.0 * math.pi) + 0.5 * logdet - fmode
    return dict(mode=mode, var=var, int=integral, converge=~active, evals=evals, iterations=it)


# ------------------------------------------------------------------ Metro, C chains in lock-step
def Metro_multichain(mu, v, N, samp_size, batch_size, alpha, logpost_fn, rng, max_proposals=None):
    """C independent copies of the reference's `Metro` loop ([A]:484-539) after its Laplace step.
    mu[C,k]: chain starts (pars$mu), v[C,k,k] or [k,k]: pars$v.  One `logpost_fn` call per
    lock-step iteration over the still-running chains.
    -> list of C dicts(sample[samp_size,k], beta[samp_size], n_accept, n_proposals, pv, n_na).
    `rng` supplies, per iteration and in this order, u ~ U(0,1)[C] then z ~ N(0,1)[C,k]; chain c uses
    row c (so a sequential re-run with the same rows reproduces each chain exactly)."""
    mu = np.atleast_2d(np.asarray(mu, dtype=np.float64))
    C, k = mu.shape
    v = np.asarray(v, dtype=np.float64)
    if v.ndim == 2:
        v = np.repeat(v[None], C, axis=0)
    chol = np.stack([np.linalg.cholesky(math.sqrt(2.0) * v[c]) for c in range(C)])
    samp = np.zeros((C, N, k))
    beta = np.zeros((C, N))
    first = logpost_fn(mu)
    l_old = np.asarray(first["val"], dtype=np.float64).copy()
    theta_old = mu.copy()
    kacc = np.zeros(C, dtype=np.int64)             # accepted samples stored (R's k-1)
    nprop = np.zeros(C, dtype=np.int64)
    n_na = np.zeros(C, dtype=np.int64)
    pv = np.zeros(C)
    running = np.ones(C, dtype=bool)
    tested = np.full(C, -1, dtype=np.int64)
    limit = max_proposals if max_proposals is not None else 1000 * N
    while running.any() and nprop.max() < limit:
        u = rng.random(C)
        z = rng.standard_normal((C, k))
        idx = np.nonzero(running)[0]
        cand = theta_old[idx] + np.einsum("cij,cj->ci", chol[idx], z[idx])
        res = logpost_fn(cand)
        l_cand = np.asarray(res["val"], dtype=np.float64)
        b_cand = np.asarray(res["beta"], dtype=np.float64)
        nprop[idx] += 1
        R = l_cand - l_old[idx]
        na = ~np.isfinite(l_cand)
        n_na[idx] += na
        acc = (~na) & (R > np.log(u[idx]))
        for j in np.nonzero(acc)[0]:
            c = idx[j]
            samp[c, kacc[c]] = cand[j]
            beta[c, kacc[c]] = b_cand[j]
            theta_old[c] = cand[j]
            l_old[c] = l_cand[j]
            kacc[c] += 1
        # stationarity test: every iteration of the reference's loop re-evaluates this condition,
        # so a chain resting on a multiple of batch.size repeats the (identical) test -- same result
        # (the result only changes when a sample was accepted, so it is computed once per count)
        for c in idx:
            ka = kacc[c]
            if ka >= samp_size and ka % batch_size == 0 and tested[c] != ka:
                # R: samp[(k-samp.size):(k-1)] with k-1 = ka accepted -> 0-based rows ka-samp_size .. ka-1, column 1
                pv[c] = geweke_pvalue(samp[c, ka - samp_size:ka, 0])
                tested[c] = ka
            if kacc[c] >= N or pv[c] >= alpha:
                running[c] = False
    out = []
    for c in range(C):
        ka = int(kacc[c])
        lo = max(ka - samp_size, 0)
        out.append(dict(sample=samp[c, lo:ka].copy(), beta=beta[c, lo:ka].copy(), n_accept=ka,
                        n_proposals=int(nprop[c]), pv=float(pv[c]), n_na=int(n_na[c])))
    return out
