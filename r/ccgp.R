# ccgp.R -- drop-in definitions of the reference's hot-path closures on top of libccgp.so.
#
#   source("<reference script's function section>")     # everything before "End of Functions!!!"
#   dyn.load("ccgp_shim.so"); source("ccgp.R")
#   ccgp.init(device = 0)            # or ccgp.init(n.gpus = 0): every GPU of the box behind one context
#   ccgp.use.script("A")             # which script's logpost / kernel family / prior the names below stand for
#
# The names defined here shadow the pure-R versions of the SAME script; the script's drivers (Metro,
# factors.frame, factors, predict.post, prediction, compare.GP, Combined.GP.fit) are NOT redefined and run
# unchanged on top of them: `predict.post` keeps reading R.Inv / mean.factor / var.factor out of the pars row
# ([A]:604-623) that `logpost` -> `factors.frame` fill, exactly as before.  The batched siblings
# (logpost.batch, predict.post.batch, prediction.table, choose.hyperpars' single sweep, Metro.multichain,
# Batch.Entropy.optim's device stencil) are what turn the kernel speed into wall-clock time.
#
# This file has never been executed in the build image (no R there): tests/test_r_boundary.py checks its
# delimiters, that every .Call target is a registered shim routine with the matching argument count, and that
# every reference closure of SURVEY 8b is defined; tests/test_r_crosscheck.py is the hook that runs real R.
#
# Script aliases: A 2D Anisotropic Public, I 2D Isotropic Public, V 2D Isotropic Advanced, M Batch Sequential
# ME Design, G Ground Vibrations, H Heat Exchanger, D1 1D Public (Matern + Matern), D2 1D Two Families.
# Families (include/ccgp.h): 0 GAUSS_ISO (p, theta1, theta2); 1 GAUSS_ANISO_LAMBDA (p, theta_1..theta_d, lambda);
# 2 GAUSS_ISO_RAW2 ([V]: second component scale = lambda); 3 MATERN1D; 4 MATERN_SPLINE1D.

.ccgp <- new.env()
.ccgp$eps <- .Machine$double.eps            # solve()'s default tol: rcond below it -> error -> NA ([A]:448-449)

ccgp.init <- function(device = 0L, n.gpus = NULL) {
  .ccgp$ctx <- if (is.null(n.gpus)) .Call("ccgp_R_create", as.integer(device)) else .Call("ccgp_R_create_multi", as.integer(n.gpus))
  .ccgp$design <- NULL
  if (is.null(.ccgp$cfg)) ccgp.use.script("A")
  invisible(.Call("ccgp_R_num_gpus", .ccgp$ctx))
}

# per-script constants: kernel family, prior line of logpost, sweep constants of likeli.hyperpars
ccgp.use.script <- function(script) {
  cfg <- switch(script,
    A  = list(family = 1L, prior = function(th, ...) { lam <- exp(th[, 4]); -th[, 1] - th[, 1]^2 / 2 - th[, 2] - th[, 2]^2 / 2 - 4 * th[, 4] - 4 / lam }),  # [A]:462
    I  = list(family = 0L, prior = function(th, ...) -4 * th[, 1] - 2 / exp(th[, 1]) - 6 * th[, 2] - 16 / exp(th[, 2])),                                   # [I]:453
    M  = list(family = 0L, prior = function(th, ...) -4 * th[, 1] - 2 / exp(th[, 1]) - 6 * th[, 2] - 16 / exp(th[, 2])),                                   # [M]:450
    G  = list(family = 0L, prior = function(th, ...) -4 * th[, 1] - 1 / exp(th[, 1]) - 6 * th[, 2] - 75 / exp(th[, 2])),                                   # [G]:450
    V  = list(family = 2L, N = 1728L, tau = 100, log.argmax = FALSE,                                                                                        # [V]:467, 555-556, 598
              prior = function(th, a, b) -(a[1] + 1) * th[, 1] - a[2] / exp(th[, 1]) - (b[1] + 1) * th[, 2] - b[2] / exp(th[, 2])),
    H  = list(family = 0L, N = 1000L, tau = 50, log.argmax = TRUE,                                                                                          # [H]:462, 552-553, 591
              prior = function(th, a, b) -(a[1] + 1) * th[, 1] - a[2] / exp(th[, 1]) - (b[1] + 1) * th[, 2] - b[2] / exp(th[, 2])),
    D1 = list(family = 3L, prior = function(th, ...) -4 * th[, 1] - 2 / exp(th[, 1]) - 6 * th[, 2] - 16 / exp(th[, 2])),                                   # [D1]:636
    D2 = list(family = 4L, prior = function(th, ...) -4 * th[, 1] - 2 / exp(th[, 1]) - 6 * th[, 2] - 16 / exp(th[, 2])),                                   # [D2]:598
    stop("unknown script alias: ", script))
  cfg$script <- script
  .ccgp$cfg <- cfg
  invisible(cfg)
}

.ccgp.design <- function(D.train, y) {
  key <- list(D.train, y)
  if (!identical(.ccgp$design, key)) {
    .Call("ccgp_R_set_design", .ccgp$ctx, matrix(as.double(as.matrix(D.train)), nrow = nrow(as.matrix(D.train))), as.double(c(y)))
    .ccgp$design <- key
  }
}

.ccgp.nu <- function(nu) if (!is.null(nu) && !identical(.ccgp$nu, nu)) { .Call("ccgp_R_set_matern_nu", .ccgp$ctx, as.double(nu)); .ccgp$nu <- nu }

# log-Jacobian of logpost's re-parametrisation ([A]:459, [I]:452): rows of Theta are (psi.., phi[, zeta])
.ccgp.jacobian <- function(Theta, family, d) {
  if (family == 1L) { phi <- Theta[, d + 1]; -phi - 2 * log(1 + exp(-phi)) + rowSums(Theta[, 1:d, drop = FALSE]) + Theta[, d + 2] }
  else { phi <- Theta[, 3]; -phi - 2 * log(1 + exp(-phi)) + Theta[, 1] + Theta[, 2] }
}

# ---- batched likelihood ----------------------------------------------------------------------------------------
# Theta: B x k matrix of real-line rows.  Returns list(val, beta, loglik, status); NA where the reference returns NA:
# the factorisation breaks down (status 1) or rcond_1(R) < .Machine$double.eps (status 2), `try(solve(R))` at [A]:448.
logpost.batch <- function(D.train, Theta, y, sigma2, ..., nu = NULL, na.rule = "rcond") {
  cfg <- .ccgp$cfg
  .ccgp.design(D.train, y); .ccgp.nu(nu)
  Theta <- matrix(as.double(Theta), nrow = nrow(Theta))
  r <- .Call("ccgp_R_nll_batch", .ccgp$ctx, cfg$family, 1L, Theta, as.double(sigma2), 0L, 0)
  status <- r[[3]]
  if (na.rule == "rcond") {
    rc <- .Call("ccgp_R_rcond_batch", .ccgp$ctx, cfg$family, 1L, Theta)
    status[status == 0L & !(rc[[1]] >= .ccgp$eps)] <- 2L
  }
  loglik <- -r[[1]]; loglik[status != 0L] <- NA
  beta <- r[[2]]; beta[status != 0L] <- NA
  list(val = loglik + .ccgp.jacobian(Theta, cfg$family, ncol(as.matrix(D.train))) + cfg$prior(Theta, ...),
       beta = beta, loglik = loglik, status = status)
}

# ---- reference signatures ---------------------------------------------------------------------------------------
# logpost(D.train, theta, y, sigma2)                           [A]:433-467  [I]:433-457  [M]:430-454  [G]:429-454
# logpost(D.train, theta, y, sigma2, theta.pars, lambda.pars)  [V]:447-471  (also returns like = exp(log.like))
# logpost(D.train, theta, y, sigma2, theta1.pars, theta2.pars) [H]:441-466
# logpost(D.train, theta, y, sigma2, nu)                       [D1]:609-641 [D2]:576-602
# -> list(val, beta, R.Inv): R.Inv is what factors.frame stores per accepted sample ([A]:521, 587).
logpost <- function(D.train, theta, y, sigma2, ...) {
  cfg <- .ccgp$cfg
  extra <- list(...)
  nu <- if (cfg$family >= 3L) { if (!is.null(extra$nu)) extra$nu else extra[[1]] } else NULL
  Th <- matrix(as.double(theta), nrow = 1)
  b <- if (cfg$family >= 3L) logpost.batch(D.train, Th, y, sigma2, nu = nu) else logpost.batch(D.train, Th, y, sigma2, ...)
  n <- nrow(as.matrix(D.train))
  out <- if (b$status[1] == 0L) {
    ri <- .Call("ccgp_R_rinv_batch", .ccgp$ctx, cfg$family, 1L, Th, as.integer(n))
    list(val = b$val[1], beta = b$beta[1], R.Inv = matrix(ri[[1]], n, n))
  } else list(val = NA, beta = NA, R.Inv = NA)
  if (cfg$script == "V") out$like <- exp(b$loglik[1])          # [V]:470
  out
}

# Mixed.corr.matrix(D.train, p, theta1, theta2)          [I]:400-407 [M]:397-404 [G]:396 [H]:408
# Mixed.corr.matrix(D.train, p, theta1, theta2, lambda)  [A]:399-406 (theta1, theta2 = the two coordinate scales)
# Mixed.corr.matrix(D.train, p, theta1, lambda)          [V]:414-421 (4th argument is the raw second scale)
# Mixed.corr.matrix(D.train, p, theta1, theta2, nu)      [D1]:577-584, [D2] (corr.matrix.combined, [D2]:454-462)
.ccgp.params <- function(p, theta1, theta2, fifth) {
  fam <- .ccgp$cfg$family
  if (fam == 1L) c(p, theta1, theta2, fifth) else { if (fam >= 3L) .ccgp.nu(fifth); c(p, theta1, theta2) }
}
Mixed.corr.matrix <- function(D.train, p, theta1, theta2, lambda = NULL, nu = lambda) {
  D <- matrix(as.double(as.matrix(D.train)), nrow = nrow(as.matrix(D.train)))
  .Call("ccgp_R_mixed_corr", .ccgp$ctx, .ccgp$cfg$family, as.double(.ccgp.params(p, theta1, theta2, nu)), D, NULL)
}

# Mixed.corr.vec(x.new, D.train, p, theta1, theta2[, lambda | nu])   [A]:416-422 [I]:417-423 [V]:431-437 [D1]:593-600
# [V]'s predict.post calls it with theta1 * (1 + lambda) as the second scale ([V]:672, quirk Q2): that is the
# plain iso form, so family 2 (RAW2) and family 0 coincide here -- both take the 5th slot as a raw scale.
# [D2]'s corr.vec.combined returns before normalising ([D2]:479, quirk Q3); the device kernel reproduces that
# inside ccgp_predict; this single-site helper returns the normalised vector, as [D1]'s does.
Mixed.corr.vec <- function(x.new, D.train, p, theta1, theta2, lambda = NULL, nu = lambda) {
  D <- matrix(as.double(as.matrix(D.train)), nrow = nrow(as.matrix(D.train))); x <- matrix(as.double(x.new), nrow = 1)
  fam <- .ccgp$cfg$family
  if (fam == 2L) fam <- 0L
  c(.Call("ccgp_R_mixed_corr", .ccgp$ctx, fam, as.double(.ccgp.params(p, theta1, theta2, nu)), x, D))
}

# [M]:835-848
cross.corr.matrix <- function(D.old, D.new, theta)
  .Call("ccgp_R_mixed_corr", .ccgp$ctx, 0L, c(1, theta, theta), matrix(as.double(D.new), nrow = nrow(D.new)),
        matrix(as.double(D.old), nrow = nrow(D.old)))

# ---- empirical-Bayes sweep ([V]:552-599, [H]:549-595) -----------------------------------------------------------
.ccgp.sweep <- function(theta1.pars, theta2.pars, N) {
  p <- runif.halton(N, 1)                                     # one Halton stream for p, theta1 and theta2 (quirk Q5)
  cbind(p, sapply(p, qigamma, alpha = theta1.pars[1], beta = theta1.pars[2]),
        sapply(p, qigamma, alpha = theta2.pars[1], beta = theta2.pars[2]))
}
likeli.hyperpars <- function(D.train, y.train, theta1.pars, theta2.pars, sigma2) {
  cfg <- .ccgp$cfg
  .ccgp.design(D.train, y.train)
  r <- .Call("ccgp_R_nll_batch", .ccgp$ctx, 0L, 0L, .ccgp.sweep(theta1.pars, theta2.pars, cfg$N), as.double(sigma2), 1L, as.double(cfg$tau))
  mean(exp(-r[[1]]))
}
# all H x N candidates in ONE call (103 680 at [V]'s sizes, 624 000 at [H]'s)
choose.hyperpars <- function(D.train, y.train, hyperpars.matrix, sigma2) {
  cfg <- .ccgp$cfg
  hp <- as.matrix(hyperpars.matrix); H <- nrow(hp)
  .ccgp.design(D.train, y.train)
  cand <- do.call(rbind, lapply(seq_len(H), function(i) .ccgp.sweep(hp[i, 1:2], hp[i, 3:4], cfg$N)))
  r <- .Call("ccgp_R_nll_batch", .ccgp$ctx, 0L, 0L, cand, as.double(sigma2), 1L, as.double(cfg$tau))
  likes <- colMeans(matrix(exp(-r[[1]]), nrow = cfg$N))
  log.likes <- if (cfg$log.argmax) log(likes) else likes       # [H]:591 logs first, [V]:598 does not
  list(pars = hyperpars.matrix[which.max(log.likes), ], likelihoods = log.likes)
}

# best row of a batch of natural-scale candidates by likelihood alone -- the which.max of a grid / multi-start
# maximum-likelihood search (BASELINE configs[3]) without shipping B values back: tau = NULL is logpost's GLS
# likelihood ([A]:452-455), tau > 0 the zero-mean + tau^2 form of cond.like ([V]:564-575).  On a multi-GPU context
# the (min, index) pair is reduced with NCCL inside the library.  Returns list(loglik, index) (1-based, first on ties).
loglik.argmax <- function(D.train, y, cand, sigma2, tau = NULL) {
  .ccgp.design(D.train, y)
  r <- .Call("ccgp_R_nll_argmin", .ccgp$ctx, .ccgp$cfg$family, 0L, matrix(as.double(cand), nrow = nrow(cand)), as.double(sigma2),
             if (is.null(tau)) 0L else 1L, if (is.null(tau)) 0 else as.double(tau))
  list(loglik = -r[[1]], index = r[[2]])
}

# ---- prediction ---------------------------------------------------------------------------------------------------
# pars: S x k natural-scale rows in the script's own order -- (p, theta1, theta2, lambda) [A]; (p, theta1, theta2)
# [I][M][G][H][D1][D2]; (p, theta1, lambda) [V] -- i.e. the first columns of factors.frame's output.
# Returns list(mean, var), each T x S: entry [t, s] = predict.post(D.new[t, ], D.train, pars.frame[s, ], sigma2).
predict.post.batch <- function(D.new, D.train, y.train, pars, sigma2, nu = NULL) {
  cfg <- .ccgp$cfg
  .ccgp.design(D.train, y.train); .ccgp.nu(nu)
  k <- if (cfg$family == 1L) ncol(as.matrix(D.train)) + 2L else 3L
  P <- matrix(as.double(as.matrix(pars)[, 1:k, drop = FALSE]), ncol = k)
  pv <- NULL; vf <- -1L
  if (cfg$script == "V") { pv <- cbind(P[, 1], P[, 2], P[, 2] * (1 + P[, 3])); vf <- 0L }   # [V]:672
  r <- .Call("ccgp_R_predict", .ccgp$ctx, cfg$family, P, vf, pv, matrix(as.double(as.matrix(D.new)), nrow = nrow(as.matrix(D.new))), as.double(sigma2))
  list(mean = r[[1]], var = r[[2]])
}

# factors.frame ([A]:572-592) without its wire format: factor every posterior row ONCE, keep the factors on the device
# (HBM, keyed by the row index), predict at any number of site sets from them.  `pars` as in predict.post.batch.
#   ff <- factors.device(D.train, y.train, pars.frame)
#   tab <- predict.post.factors(ff, D.new, sigma2)     # list(mean, var), each T x S, identical to predict.post.batch
factors.device <- function(D.train, y.train, pars, nu = NULL) {
  cfg <- .ccgp$cfg
  .ccgp.design(D.train, y.train); .ccgp.nu(nu)
  k <- if (cfg$family == 1L) ncol(as.matrix(D.train)) + 2L else 3L
  P <- matrix(as.double(as.matrix(pars)[, 1:k, drop = FALSE]), ncol = k)
  pv <- NULL; vf <- -1L
  if (cfg$script == "V") { pv <- cbind(P[, 1], P[, 2], P[, 2] * (1 + P[, 3])); vf <- 0L }   # [V]:672
  .Call("ccgp_R_factors_create", .ccgp$ctx, cfg$family, P, vf, pv)
}
predict.post.factors <- function(ff, D.new, sigma2) {
  r <- .Call("ccgp_R_factors_predict", .ccgp$ctx, ff, matrix(as.double(as.matrix(D.new)), nrow = nrow(as.matrix(D.new))), as.double(sigma2))
  list(mean = r[[1]], var = r[[2]])
}
factors.release <- function(ff) invisible(.Call("ccgp_R_factors_release", ff))

# t(apply_pb(D.new, 1, prediction, alpha, code, pars.frame, D.train, sigma2)) ([A]:685) from ONE predictive table:
# per site the same statements as `prediction` ([A]:644-653): mean of means, rnorm draws, quantiles.
prediction.table <- function(D.new, alpha, pars.frame, D.train, y.train, sigma2, nu = NULL, drop.negative.var = FALSE) {
  tab <- predict.post.batch(D.new, D.train, y.train, pars.frame, sigma2, nu)
  out <- t(sapply(seq_len(nrow(tab$mean)), function(t) {
    m <- tab$mean[t, ]; v <- tab$var[t, ]
    if (drop.negative.var) { keep <- v >= 0; m <- m[keep]; v <- v[keep] }     # [I]:625, [V]:696, [M]:626 drop such rows
    y.hat <- mean(m)
    pp <- rnorm(length(m), m, sqrt(v))
    q <- quantile(pp, c(alpha / 2, 1 - alpha / 2))
    c(y.hat, mean(y.hat <= pp), q[1], q[2])
  }))
  colnames(out) <- c("y.hat.Combined", "Quant.Combined", "LL.Combined", "UL.Combined")
  out
}

# ---- maximum-entropy design ([M]:835-948) ---------------------------------------------------------------------------
entropy.batch <- function(D.old, D.new.list, params) {
  # D.new.list: list of n.new x d matrices; params: P x 3 (p, theta1, theta2); returns the C x P matrix of -det
  n.new <- nrow(D.new.list[[1]]); d <- ncol(D.new.list[[1]])
  Dn <- sapply(D.new.list, function(D) as.double(c(D)))
  .Call("ccgp_R_me_schur_batch", .ccgp$ctx, if (is.null(D.old)) NULL else matrix(as.double(D.old), nrow = nrow(D.old)),
        Dn, as.integer(n.new), as.integer(d), matrix(as.double(params), ncol = 3))[[1]]
}
# which.min over a pool of candidate designs for each parameter row ([M]:944-945): list(value, index)
entropy.argmin <- function(D.old, D.new.list, params) {
  n.new <- nrow(D.new.list[[1]]); d <- ncol(D.new.list[[1]])
  Dn <- sapply(D.new.list, function(D) as.double(c(D)))
  r <- .Call("ccgp_R_me_argmin", .ccgp$ctx, if (is.null(D.old)) NULL else matrix(as.double(D.old), nrow = nrow(D.old)),
             Dn, as.integer(n.new), as.integer(d), matrix(as.double(params), ncol = 3))
  list(value = r[[1]], index = r[[2]])
}
Entropy <- function(D, p, theta1, theta2) entropy.batch(NULL, list(as.matrix(D)), matrix(c(p, theta1, theta2), 1))[1, 1]   # [M]:856-861
# [M]:869-877 (R.old.Inv is accepted for signature parity; the kernel refactors R.old itself)
Augmented.Mixed.Entropy <- function(D.old, D.new, p, theta1, theta2, R.old.Inv = NULL)
  entropy.batch(D.old, list(as.matrix(D.new)), matrix(c(p, theta1, theta2), 1))[1, 1]

# the 2m+1 central-difference points of optim's L-BFGS-B gradient ([M]:936; ndeps = 1e-3, clipped to the box as
# optim's fmingr does) for one or many base designs in one launch: X is (n.new * d) x K
.ccgp.stencil <- function(D.old, X, n.new, d, params, group, h = 1e-3, lo = -1, hi = 1)
  .Call("ccgp_R_me_schur_stencil", .ccgp$ctx, matrix(as.double(D.old), nrow = nrow(D.old)), matrix(as.double(X), nrow = n.new * d),
        as.integer(n.new), as.integer(d), matrix(as.double(params), ncol = 3), as.integer(group), as.double(h), as.double(lo), as.double(hi))

# Batch.Entropy.optim(D.old, n.new, d, p, theta1, theta2, n.starts) [M]:920-948: same starts (optimumLHS), same
# optimiser (optim, L-BFGS-B, box [-1, 1]); the criterion AND its finite-difference gradient come from one
# device launch per optim iteration instead of 2 m + 1 R-level calls.  Returns list(Design, log.entropy) with
# log.entropy = -min.val, a determinant, as the reference does (quirk Q4).
Batch.Entropy.optim <- function(D.old, n.new, d, p, theta1, theta2, n.starts) {
  prm <- matrix(c(p, theta1, theta2), 1); m <- n.new * d; h <- 1e-3
  fn <- function(x) .ccgp.stencil(D.old, matrix(x, ncol = 1), n.new, d, prm, 1L)[1, 1]
  gr <- function(x) {
    v <- .ccgp.stencil(D.old, matrix(x, ncol = 1), n.new, d, prm, 1L)[, 1]
    (v[2 * seq_len(m)] - v[2 * seq_len(m) + 1]) / (pmin(x + h, 1) - pmax(x - h, -1))
  }
  vals <- rep(0, n.starts); designs <- list()
  for (k in 1:n.starts) {
    start <- c(-1 + 2 * optimumLHS(n.new, d))
    optimum <- optim(start, fn, gr, method = "L-BFGS-B", lower = rep(-1, m), upper = rep(1, m))
    designs[[k]] <- matrix(optimum$par, ncol = d); vals[k] <- optimum$value
  }
  list(Design = designs[[which.min(vals)]], log.entropy = -vals[which.min(vals)])
}
# Entropy.optim(n, d, p, theta1, theta2, n.starts) [M]:886-912: first batch, criterion -det R(D)
Entropy.optim <- function(n, d, p, theta1, theta2, n.starts) {
  fn <- function(x) Entropy(matrix(x, ncol = d), p, theta1, theta2)
  vals <- rep(0, n.starts); designs <- list()
  for (k in 1:n.starts) {
    optimum <- optim(c(-1 + 2 * optimumLHS(n, d)), fn, method = "L-BFGS-B", lower = rep(-1, n * d), upper = rep(1, n * d))
    designs[[k]] <- matrix(optimum$par, ncol = d); vals[k] <- optimum$value
  }
  list(Design = designs[[which.min(vals)]], log.entropy = -vals[which.min(vals)])
}
# paired criterion for lock-step optimisers over (posterior draw, start) pairs: column c of the
# (n.new*d) x (P*group) matrix of c(D.new) vectors is evaluated against params[c %/% group + 1, ] only
entropy.paired <- function(D.old, D.new.cols, n.new, d, params, group)
  .Call("ccgp_R_me_schur_paired", .ccgp$ctx, D.old, D.new.cols, as.integer(n.new), as.integer(d), params, as.integer(group))[[1]]

# log det R[S, S] for the rows of an integer index matrix (1-based, C x m) into a pool of points (SURVEY 8d ME-B)
subset.logdet.batch <- function(pool, idx, params, family = .ccgp$cfg$family) {
  storage.mode(idx) <- "integer"
  .Call("ccgp_R_subset_logdet_batch", .ccgp$ctx, matrix(as.double(pool), nrow = nrow(pool)), idx, as.integer(family), as.double(params))[[1]]
}

# ---- lock-step Metro (SURVEY 8f rank 1) -----------------------------------------------------------------------------
# C independent copies of the reference's Metro loop ([A]:484-539, after its laplace() step),
# advanced together: one logpost.batch() call per iteration over the chains still running.
# pars: list(mu = C x k matrix of chain starts, v = k x k proposal matrix, as laplace() returns it).
# Same proposal (rmnorm(1, theta.old, sqrt(2) * v)), same acceptance rule, same Geweke stop on the
# first column; an NA likelihood rejects the candidate instead of aborting the script.
# R.Inv of the retained samples is recomputed on demand (logpost on the kept rows) instead of
# being carried through the loop.  `...` goes to the script's prior ([V]: theta.pars, lambda.pars; [H] likewise).
Metro.multichain <- function(pars, N, samp.size, batch.size, alpha, D.train, sigma2, y, ...) {
  mu <- as.matrix(pars$mu); C <- nrow(mu); k <- ncol(mu)
  samp <- array(0, c(C, N, k)); beta <- matrix(0, C, N)
  theta.old <- mu
  l.old <- logpost.batch(D.train, theta.old, y, sigma2, ...)$val
  kk <- rep(1L, C); pv <- rep(0, C); running <- rep(TRUE, C)
  while (any(running)) {
    idx <- which(running)
    u <- runif(length(idx))
    cand <- t(sapply(idx, function(c) rmnorm(1, as.vector(theta.old[c, ]), sqrt(2) * pars$v)))
    l.cand <- logpost.batch(D.train, cand, y, sigma2, ...)
    R <- l.cand$val - l.old[idx]
    acc <- !is.na(R) & R > log(u)
    for (j in which(acc)) {
      c <- idx[j]
      samp[c, kk[c], ] <- cand[j, ]; beta[c, kk[c]] <- l.cand$beta[j]
      theta.old[c, ] <- cand[j, ]; l.old[c] <- l.cand$val[j]; kk[c] <- kk[c] + 1L
    }
    for (c in idx) {
      if ((kk[c] - 1) >= samp.size && (kk[c] - 1) %% batch.size == 0) {
        p1 <- try(min(2 * (1 - pnorm(abs(geweke.diag(mcmc(samp[c, (kk[c] - samp.size):(kk[c] - 1), 1]))$z)))), silent = TRUE)
        pv[c] <- if (inherits(p1, "try-error") || is.na(p1)) 0 else p1
      }
      if (kk[c] > N || pv[c] >= alpha) running[c] <- FALSE
    }
  }
  lapply(seq_len(C), function(c) {
    rows <- (kk[c] - samp.size):(kk[c] - 1)
    list(sample = data.frame(samp[c, rows, ]), beta = beta[c, rows])
  })
}

# ---- k-medoids design (reference ReadMe.md:54-60) ---------------------------------------------------------------------
# 7-medoids (cluster::pam) over the points of all second-batch designs -> the next batch.
# subdesigns: (C * n.new) x d matrix (the layout of All_Subdesigns.txt).
kmedoids.design <- function(D.old, subdesigns, k) {
  P <- matrix(as.double(as.matrix(subdesigns)), nrow = nrow(subdesigns))
  r <- .Call("ccgp_R_kmedoids_pam", .ccgp$ctx, P, as.integer(k), 1000L)   # list(medoid rows (1-based), cost, swaps)
  list(Design = rbind(as.matrix(D.old), P[r[[1]], , drop = FALSE]), medoid.rows = r[[1]], cost = r[[2]])
}

# ---- CGP comparator ([A]:60-237): the 505-candidate start sweep and the leave-one-out loop on the device ----------
# Inside the reference's CGP() the two lines to swap are
#   cand_obj <- apply(starts, 1, var.MLE.DK)                        ([A]:148)  ->  var.MLE.DK.batch(DD, yobs, starts)
#   for (jf in 1:n) { ... Yjfp[jf] <- beta + t(q) %*% temp }        ([A]:166-199) ->  Yjfp <- cgp.jackknife(DD, yobs, op$par)
# (optim's own calls of var.MLE.DK, [A]:151-156, stay scalar R: they are sequential line searches.)
.ccgp.standardise <- function(X) apply(as.matrix(X), 2, function(x) (x - min(x)) / max(x - min(x)))     # [A]:70
var.MLE.DK.batch <- function(X, yobs, starts)
  .Call("ccgp_R_cgp_objective_batch", .ccgp$ctx, matrix(as.double(.ccgp.standardise(X)), nrow = nrow(X)), as.double(yobs),
        matrix(as.double(starts), nrow = nrow(starts)))
cgp.jackknife <- function(X, yobs, par)
  .Call("ccgp_R_cgp_jackknife", .ccgp$ctx, matrix(as.double(.ccgp.standardise(X)), nrow = nrow(X)), as.double(yobs), as.double(par))
