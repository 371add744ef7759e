"""Lock-step generator of the second-batch ME designs (first half).

`All_Subdesigns.txt` is "1000 size-7 second batch designs, each corresponding to a single triplet
of parameters drawn from the posterior" (reference ReadMe.md:54-56): `Batch.Entropy.optim`
([M]:920-948) run once per posterior draw -- 25 L-BFGS-B starts each, every objective evaluation a
7x7 Schur determinant ([M]:869-877), every gradient a 2*14-point finite-difference stencil.  The
reference does this one `optim` call at a time (and the driving script is not shipped).

Here all (draw, start) problems advance together: one iteration = ONE paired batch of Schur
determinants for the gradients (K problems x 28 stencil points generated on the device,
`ccgp_me_schur_stencil`) plus one per line-search trial (`ccgp_me_schur_paired`).  The optimiser is a bound-projected L-BFGS (two-loop recursion, memory 5,
Armijo backtracking on the projected step, variables pinned at a bound whose gradient points
outward are frozen): same problem class and stopping rule as optim's L-BFGS-B (factr = 1e7,
pgtol = 0, maxit = 100, central differences with ndeps = 1e-3), not a transcription of lbfgsb.c --
the reference's own starts are unseeded LHDs, so the comparison is on the criterion reached.
"""
from __future__ import annotations

import numpy as np


def lockstep_minimize(fun_batch, X0, lower, upper, maxit=100, factr=1e7, ndeps=1e-3, memory=5, max_backtrack=12,
                      stencil_batch=None):
    """Minimise K independent box-constrained problems in lock-step.
    fun_batch(X[B, m], owner[B]) -> f[B]: objective of problem owner[b] at X[b] (owner is sorted, every problem
    appears the same number of times in one call).  stencil_batch(X[K, m]) -> vals[K, 2m+1], optional: the
    objective at the central-difference points of every row (column 0: the row itself, 1+2i / 2+2i: coordinate i
    plus / minus ndeps, clipped to the box), produced without materialising them on the host.
    -> dict(x[K,m], f[K], iterations, evals, converged[K])."""
    X = np.clip(np.asarray(X0, dtype=np.float64), lower, upper)
    K, m = X.shape
    ids = np.arange(K)
    evals = 0

    def F(Xb, owner):
        nonlocal evals
        evals += Xb.shape[0]
        v = np.asarray(fun_batch(Xb, owner), dtype=np.float64)
        return np.where(np.isfinite(v), v, np.inf)

    def grad(Xc):
        nonlocal evals
        if stencil_batch is not None:
            vals = np.asarray(stencil_batch(Xc), dtype=np.float64)
            evals += vals.size
            vals = np.where(np.isfinite(vals), vals, np.inf)
            eps = np.minimum(Xc + ndeps, upper) - np.maximum(Xc - ndeps, lower)
            return (vals[:, 1::2] - vals[:, 2::2]) / eps
        # central differences, the step shrunk at a bound exactly as optim does not: R's fmingr evaluates
        # outside-the-box points clipped to the bound with the epsilon adjusted -- we do the same
        Xp = np.repeat(Xc[:, None, :], m, axis=1)
        Xm = Xp.copy()
        ar = np.arange(m)
        Xp[:, ar, ar] = np.minimum(Xc + ndeps, upper)
        Xm[:, ar, ar] = np.maximum(Xc - ndeps, lower)
        both = np.concatenate([Xp, Xm], axis=1).reshape(K * 2 * m, m)
        vals = F(both, np.repeat(ids, 2 * m)).reshape(K, 2 * m)
        eps = Xp[:, ar, ar] - Xm[:, ar, ar]
        return (vals[:, :m] - vals[:, m:]) / eps

    f = F(X, ids)
    g = grad(X)
    S = np.zeros((memory, K, m)); Y = np.zeros((memory, K, m)); rho = np.zeros((memory, K))
    nhist = 0
    active = np.ones(K, dtype=bool)
    tol = factr * np.finfo(np.float64).eps
    it = 0
    while active.any() and it < maxit:
        it += 1
        # variables sitting on a bound with the gradient pushing outward stay there this iteration
        pinned = ((X <= lower) & (g > 0)) | ((X >= upper) & (g < 0))
        q = np.where(pinned, 0.0, g)
        alphas = []
        for h in range(nhist - 1, -1, -1):
            a = rho[h] * np.einsum("km,km->k", S[h], q)
            alphas.append(a)
            q = q - a[:, None] * Y[h]
        if nhist > 0:
            yy = np.einsum("km,km->k", Y[nhist - 1], Y[nhist - 1])
            sy = np.einsum("km,km->k", S[nhist - 1], Y[nhist - 1])
            gamma = np.where(yy > 0, sy / np.where(yy > 0, yy, 1.0), 1.0)
        else:
            gn = np.linalg.norm(q, axis=1)
            gamma = 1.0 / np.maximum(gn, 1e-300)           # first step of unit length
        r = gamma[:, None] * q
        for h, a in zip(range(nhist), reversed(alphas)):
            b = rho[h] * np.einsum("km,km->k", Y[h], r)
            r = r + (a - b)[:, None] * S[h]
        p = np.where(pinned, 0.0, -r)
        slope = np.einsum("km,km->k", g, p)
        bad = ~(slope < 0)                                  # not a descent direction: fall back to steepest descent
        if bad.any():
            p[bad] = np.where(pinned[bad], 0.0, -g[bad])
            slope[bad] = np.einsum("km,km->k", g[bad], p[bad])
        step = np.ones(K)
        Xn, fn = X.copy(), f.copy()
        todo = active & (slope < 0)
        for _ in range(max_backtrack):
            if not todo.any():
                break
            trial = np.clip(X + step[:, None] * p, lower, upper)
            ft = F(trial, ids)                              # every problem is evaluated (constant group size); masks pick
            dec = np.einsum("km,km->k", g, trial - X)
            ok = todo & (ft <= f + 1e-4 * dec) & np.isfinite(ft)
            Xn[ok], fn[ok] = trial[ok], ft[ok]
            todo &= ~ok
            step[todo] *= 0.5
        moved = active & (fn < f)
        gn_ = grad(Xn)
        s_new, y_new = Xn - X, gn_ - g
        sy = np.einsum("km,km->k", s_new, y_new)
        good = moved & (sy > 1e-10 * np.einsum("km,km->k", y_new, y_new))
        if nhist == memory:
            S[:-1], Y[:-1], rho[:-1] = S[1:].copy(), Y[1:].copy(), rho[1:].copy()
            nhist -= 1
        S[nhist] = np.where(good[:, None], s_new, 0.0)
        Y[nhist] = np.where(good[:, None], y_new, 0.0)
        rho[nhist] = np.where(good, 1.0 / np.where(good, sy, 1.0), 0.0)
        nhist += 1
        # optim's L-BFGS-B stop: (f_k - f_{k+1}) / max(|f_k|, |f_{k+1}|, 1) <= factr * epsmch; no move = converged
        rel = (f - fn) / np.maximum(np.maximum(np.abs(f), np.abs(fn)), 1.0)
        done = active & (~moved | (rel <= tol))
        X = np.where(active[:, None], Xn, X)
        g = np.where(active[:, None], gn_, g)
        f = np.where(active, fn, f)
        active &= ~done
    return dict(x=X, f=f, iterations=it, evals=evals, converged=~active)


def random_lhd_starts(rng, count, n_new, d):
    """`-1 + 2 * lhs::optimumLHS(n.new, d)` stand-in ([M]:933): random Latin hypercube designs on [-1, 1]^d,
    returned as count x (n_new*d) rows in c(D) (column-major) order."""
    out = np.empty((count, n_new * d))
    for c in range(count):
        lhd = (np.argsort(rng.random((n_new, d)), axis=0) + rng.random((n_new, d))) / n_new
        out[c] = (-1.0 + 2.0 * lhd).reshape(-1, order="F")
    return out


def all_subdesigns(D_old, params, n_new, d, n_starts, rng, engine, maxit=100, starts=None, device_stencil=True):
    """One `Batch.Entropy.optim` ([M]:920-948) per parameter row (p, theta1, theta2), all rows and all starts in
    lock-step.  -> dict(designs[P, n_new, d] (the All_Subdesigns array), values[P] (= -det, the minimised
    criterion), all_values[P, n_starts], iterations, evals)."""
    params = np.atleast_2d(np.asarray(params, dtype=np.float64))
    P = params.shape[0]
    K = P * n_starts
    X0 = random_lhd_starts(rng, K, n_new, d) if starts is None else np.asarray(starts, dtype=np.float64).reshape(K, n_new * d)
    D_old = None if D_old is None else np.atleast_2d(D_old)

    def fun(Xb, owner):
        # owner is sorted with a constant count per problem, problems are draw-major: designs of draw q are contiguous
        group = Xb.shape[0] // P
        designs = Xb.reshape(-1, d, n_new).transpose(0, 2, 1)
        return engine.me_schur_paired(D_old, designs, params, group)[0]

    def stencil(Xc):                                        # the 2m+1 points of every problem, generated on the device
        return engine.me_schur_stencil(D_old, Xc, n_new, d, params, n_starts, h=1e-3, lower=-1.0, upper=1.0)[0]

    res = lockstep_minimize(fun, X0, -1.0, 1.0, maxit=maxit, stencil_batch=stencil if device_stencil else None)
    fv = res["f"].reshape(P, n_starts)
    best = np.argmin(fv, axis=1)                             # which.min over the starts ([M]:944)
    xs = res["x"].reshape(P, n_starts, n_new * d)[np.arange(P), best]
    designs = xs.reshape(P, d, n_new).transpose(0, 2, 1)
    return dict(designs=designs, values=fv[np.arange(P), best], all_values=fv, iterations=res["iterations"], evals=res["evals"])
