# ccgp.R -- drop-in definitions of the reference's hot-path closures on top of libccgp.so.
# Source this AFTER the function section of a reference script (everything before its
# "End of Functions!!!" marker): the names below then shadow the pure-R versions, and the
# script's drivers (Metro, factors.frame, prediction, Batch.Entropy.optim, ...) run unchanged.
#
#   dyn.load("ccgp_shim.so"); source("ccgp.R"); ccgp.init(device = 0)
#
# Families: 0 GAUSS_ISO (p, theta1, theta2); 1 GAUSS_ANISO_LAMBDA (p, theta_1..theta_d, lambda);
# 2 GAUSS_ISO_RAW2 ([V]: second component scale = lambda).  Scalar work the reference does in R
# around the linear algebra (Jacobian, hard-coded priors, Halton grid) stays in R, unchanged.

.ccgp <- new.env()

ccgp.init <- function(device = 0L) {
  .ccgp$ctx <- .Call("ccgp_R_create", as.integer(device))
  .ccgp$design <- NULL
  invisible(TRUE)
}

.ccgp.design <- function(D.train, y) {
  key <- list(D.train, y)
  if (!identical(.ccgp$design, key)) {
    .Call("ccgp_R_set_design", .ccgp$ctx, matrix(as.double(D.train), nrow = nrow(D.train)), as.double(c(y)))
    .ccgp$design <- key
  }
}

.na.status <- function(v, status) { v[status != 0L] <- NA; v }

# ---- batched siblings (what the kept drivers should call once per batch) -------------------
logpost.batch <- function(D.train, Theta, y, sigma2, family, log.prior) {
  # Theta: B x k matrix of real-line rows; log.prior: function(Theta) -> length-B vector
  .ccgp.design(D.train, y)
  Theta <- matrix(as.double(Theta), nrow = nrow(Theta))
  r <- .Call("ccgp_R_nll_batch", .ccgp$ctx, as.integer(family), 1L, Theta, as.double(sigma2), 0L, 0)
  d <- ncol(D.train)
  if (family == 1L) {
    phi <- Theta[, d + 1]; jac <- -phi - 2 * log(1 + exp(-phi)) + rowSums(Theta[, 1:d, drop = FALSE]) + Theta[, d + 2]
  } else {
    phi <- Theta[, 3]; jac <- -phi - 2 * log(1 + exp(-phi)) + Theta[, 1] + Theta[, 2]
  }
  list(val = .na.status(-r[[1]], r[[3]]) + jac + log.prior(Theta), beta = .na.status(r[[2]], r[[3]]))
}

predict.post.batch <- function(D.new, D.train, y.train, pars, sigma2, family, pars.vec = NULL, vec.family = -1L) {
  # pars: S x k natural-scale rows (p, theta1, theta2[, lambda]); returns list(mean, var), each T x S
  .ccgp.design(D.train, y.train)
  r <- .Call("ccgp_R_predict", .ccgp$ctx, as.integer(family), matrix(as.double(as.matrix(pars)), nrow = nrow(pars)),
             as.integer(vec.family), if (is.null(pars.vec)) NULL else matrix(as.double(pars.vec), nrow = nrow(pars.vec)),
             matrix(as.double(as.matrix(D.new)), nrow = nrow(D.new)), as.double(sigma2))
  list(mean = r[[1]], var = r[[2]])
}

entropy.batch <- function(D.old, D.new.list, params) {
  # D.new.list: list of n.new x d matrices; params: P x 3 (p, theta1, theta2); returns C x P matrix of -det
  n.new <- nrow(D.new.list[[1]]); d <- ncol(D.new.list[[1]])
  Dn <- sapply(D.new.list, function(D) as.double(c(D)))
  .Call("ccgp_R_me_schur_batch", .ccgp$ctx, if (is.null(D.old)) NULL else matrix(as.double(D.old), nrow = nrow(D.old)),
        Dn, as.integer(n.new), as.integer(d), matrix(as.double(params), ncol = 3))[[1]]
}

# ---- reference signatures ---------------------------------------------------------------------
# [I]:400-407 / [A]:399-406
Mixed.corr.matrix <- function(D.train, p, theta1, theta2, lambda = NULL) {
  D <- matrix(as.double(D.train), nrow = nrow(D.train))
  if (is.null(lambda)) .Call("ccgp_R_mixed_corr", .ccgp$ctx, 0L, c(p, theta1, theta2), D, NULL)
  else .Call("ccgp_R_mixed_corr", .ccgp$ctx, 1L, c(p, theta1, theta2, lambda), D, NULL)
}

# [I]:417-423 / [A]:416-422
Mixed.corr.vec <- function(x.new, D.train, p, theta1, theta2, lambda = NULL) {
  D <- matrix(as.double(D.train), nrow = nrow(D.train)); x <- matrix(as.double(x.new), nrow = 1)
  if (is.null(lambda)) c(.Call("ccgp_R_mixed_corr", .ccgp$ctx, 0L, c(p, theta1, theta2), x, D))
  else c(.Call("ccgp_R_mixed_corr", .ccgp$ctx, 1L, c(p, theta1, theta2, lambda), x, D))
}

# [M]:835-848
cross.corr.matrix <- function(D.old, D.new, theta)
  .Call("ccgp_R_mixed_corr", .ccgp$ctx, 0L, c(1, theta, theta), matrix(as.double(D.new), nrow = nrow(D.new)),
        matrix(as.double(D.old), nrow = nrow(D.old)))

# [A]:433-467 (anisotropic script; the isotropic scripts differ only in family and prior line)
logpost <- function(D.train, theta, y, sigma2) {
  .ccgp.design(D.train, y)
  Th <- matrix(as.double(theta), nrow = 1)
  r <- .Call("ccgp_R_nll_batch", .ccgp$ctx, 1L, 1L, Th, as.double(sigma2), 0L, 0)
  ri <- .Call("ccgp_R_rinv_batch", .ccgp$ctx, 1L, 1L, Th, nrow(D.train))
  psi1 <- theta[1]; psi2 <- theta[2]; phi <- theta[3]; zeta <- theta[4]; lambda <- exp(zeta)
  log.jacob <- -phi - 2 * log(1 + exp(-phi)) + psi1 + psi2 + zeta
  log.prior <- -psi1 - psi1^2 / 2 - psi2 - psi2^2 / 2 - 4 * zeta - 4 / lambda
  ok <- r[[3]][1] == 0L
  n <- nrow(D.train)
  list(val = if (ok) -r[[1]][1] + log.jacob + log.prior else NA, beta = if (ok) r[[2]][1] else NA,
       R.Inv = if (ok) matrix(ri[[1]], n, n) else NA)
}

# [V]:552-578: N = 1728 Halton candidates, tau = 100, one batched call
likeli.hyperpars <- function(D.train, y.train, theta1.pars, theta2.pars, sigma2, N = 1728, tau = 100) {
  .ccgp.design(D.train, y.train)
  p <- runif.halton(N, 1)
  theta1 <- sapply(p, qigamma, alpha = theta1.pars[1], beta = theta1.pars[2])
  theta2 <- sapply(p, qigamma, alpha = theta2.pars[1], beta = theta2.pars[2])
  r <- .Call("ccgp_R_nll_batch", .ccgp$ctx, 0L, 0L, cbind(p, theta1, theta2), as.double(sigma2), 1L, as.double(tau))
  mean(exp(-r[[1]]))
}

# [A]:604-623 (one site, one posterior row); the table version is predict.post.batch
predict.post <- function(x.new, D.train, pars, sigma2) {
  stop("predict.post needs y.train with the CUDA path: call predict.post.batch(D.new, D.train, y.train, pars, sigma2, family)")
}

# [M]:856-861
Entropy <- function(D, p, theta1, theta2) entropy.batch(NULL, list(as.matrix(D)), matrix(c(p, theta1, theta2), 1))[1, 1]

# [M]:869-877 (R.old.Inv is accepted for signature parity; the kernel refactors R.old itself)
Augmented.Mixed.Entropy <- function(D.old, D.new, p, theta1, theta2, R.old.Inv = NULL)
  entropy.batch(D.old, list(as.matrix(D.new)), matrix(c(p, theta1, theta2), 1))[1, 1]

# ---- lock-step Metro (SURVEY 8f rank 1) --------------------------------------------------------
# C independent copies of the reference's Metro loop ([A]:484-539, after its laplace() step),
# advanced together: one logpost.batch() call per iteration over the chains still running.
# pars: list(mu = C x k matrix of chain starts, v = k x k proposal matrix, as laplace() returns it).
# Same proposal (rmnorm(1, theta.old, sqrt(2) * v)), same acceptance rule, same Geweke stop on the
# first column; an NA likelihood rejects the candidate instead of aborting the script.
# R.Inv of the retained samples is recomputed on demand (one ccgp_R_rinv_batch call) instead of
# being carried through the loop.
Metro.multichain <- function(pars, N, samp.size, batch.size, alpha, D.train, sigma2, y, family, log.prior) {
  mu <- as.matrix(pars$mu); C <- nrow(mu); k <- ncol(mu)
  samp <- array(0, c(C, N, k)); beta <- matrix(0, C, N)
  theta.old <- mu
  l.old <- logpost.batch(D.train, theta.old, y, sigma2, family, log.prior)$val
  kk <- rep(1L, C); pv <- rep(0, C); running <- rep(TRUE, C)
  while (any(running)) {
    idx <- which(running)
    u <- runif(length(idx))
    cand <- t(sapply(idx, function(c) rmnorm(1, as.vector(theta.old[c, ]), sqrt(2) * pars$v)))
    l.cand <- logpost.batch(D.train, cand, y, sigma2, family, log.prior)
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

# ---- k-medoids design (reference ReadMe.md:54-60) ------------------------------------------------
# 7-medoids (cluster::pam) over the points of all second-batch designs -> the next batch.
# subdesigns: (C * n.new) x d matrix (the layout of All_Subdesigns.txt).
kmedoids.design <- function(D.old, subdesigns, k) {
  P <- matrix(as.double(as.matrix(subdesigns)), nrow = nrow(subdesigns))
  r <- .Call("ccgp_R_kmedoids_pam", .ccgp$ctx, P, as.integer(k), 1000L)   # list(medoid rows (1-based), cost, swaps)
  list(Design = rbind(as.matrix(D.old), P[r[[1]], , drop = FALSE]), medoid.rows = r[[1]], cost = r[[2]])
}

# paired criterion for lock-step optimisers over (posterior draw, start) pairs: column c of the
# (n.new*d) x (P*group) matrix of c(D.new) vectors is evaluated against params[c %/% group + 1, ] only
entropy.paired <- function(D.old, D.new.cols, n.new, d, params, group)
  .Call("ccgp_R_me_schur_paired", .ccgp$ctx, D.old, D.new.cols, as.integer(n.new), as.integer(d), params, as.integer(group))[[1]]
