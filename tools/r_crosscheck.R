# r_crosscheck.R -- run the REFERENCE's own closures on the committed golden inputs (SURVEY 8c fallback).
#
#   Rscript tools/r_crosscheck.R <reference_root> <input_dir> <output_dir> [time]
#
# R is not in the build image, so this script has never been executed there; tests/test_r_crosscheck.py
# and bench.py call it only when an `Rscript` is on PATH and the reference scripts are reachable
# (CCGP_REFERENCE_ROOT, default /root/reference) and skip with a message otherwise.
#
# For each script the FUNCTION SECTION (everything before its "End of Functions!!!" marker: [A]:834,
# [I]:840, [V]:918, [M]:958, [H]:733, [G]:683) is evaluated in a fresh environment with the library()
# lines removed; the few third-party functions the hot path needs are supplied when their package is
# missing (mnormt::dmnorm in its pd.solve form, fOptions::runif.halton base 2, pscl::qigamma).
# Inputs are the CSV files tests/test_r_crosscheck.py exports from tests/golden/*.npz; outputs are CSV
# files the same test diffs against the oracle's golden vectors.

args <- commandArgs(trailingOnly = TRUE)
ref.root <- args[1]; in.dir <- args[2]; out.dir <- args[3]
do.time <- length(args) >= 4 && args[4] == "time"

rd <- function(name) as.matrix(read.csv(file.path(in.dir, paste0(name, ".csv")), header = FALSE))
wr <- function(x, name) write.table(format(as.matrix(x), digits = 17), file.path(out.dir, paste0(name, ".csv")),
                                    sep = ",", quote = FALSE, row.names = FALSE, col.names = FALSE)

fallbacks <- function(env) {
  if (!requireNamespace("mnormt", quietly = TRUE)) {
    # mnormt >= 1.5: pd.solve(varcov, log.det = TRUE) -> chol, chol2inv; see oracle dmnorm_pdsolve
    assign("dmnorm", function(x, mean = rep(0, d), varcov, log = FALSE) {
      d <- if (is.matrix(varcov)) ncol(varcov) else 1
      X <- t(matrix(x, ncol = d)) - mean
      S <- (varcov + t(varcov)) / 2
      u <- chol(S)
      conc <- chol2inv(u)
      Q <- colSums((conc %*% X) * X)
      logPDF <- as.vector(Q + d * logb(2 * pi) + 2 * sum(log(diag(u)))) / (-2)
      if (log) logPDF else exp(logPDF)
    }, envir = env)
  } else assign("dmnorm", mnormt::dmnorm, envir = env)
  if (!requireNamespace("fOptions", quietly = TRUE)) {
    assign("runif.halton", function(n, dimension) {
      sapply(seq_len(n), function(i) { f <- 1; r <- 0; k <- i
        while (k > 0) { f <- f / 2; r <- r + f * (k %% 2); k <- k %/% 2 }
        r })
    }, envir = env)
  } else assign("runif.halton", fOptions::runif.halton, envir = env)
  if (!requireNamespace("pscl", quietly = TRUE)) {
    assign("qigamma", function(p, alpha, beta) 1 / qgamma(1 - p, alpha, rate = beta), envir = env)
  } else assign("qigamma", pscl::qigamma, envir = env)
  # apply_pb draws a progress bar; plain apply has the same value
  assign("apply_pb", function(X, MARGIN, FUN, ...) apply(X, MARGIN, FUN, ...), envir = env)
}

load.functions <- function(rel.path) {
  lines <- readLines(file.path(ref.root, rel.path), warn = FALSE)
  end <- grep("End of Functions!!!", lines, fixed = TRUE)[1]
  lines <- lines[seq_len(end - 1)]
  lines <- lines[!grepl("^\\s*(library|require)\\(", lines)]
  env <- new.env(parent = globalenv())
  eval(parse(text = lines), envir = env)
  fallbacks(env)          # after the script so apply_pb is the plain one
  env
}

A <- load.functions("2D Codes and Designs/2D Combined GP Anisotropic Public.R")
I <- load.functions("2D Codes and Designs/2D Combined GP Isotropic Public.R")
V <- load.functions("2D Codes and Designs/2D Combined GP Isotropic Advanced.R")
M <- load.functions("Batch Sequential ME Designs/Batch Sequential ME Design.R")

# ---- [A] logpost on the n = 100 headline rows ------------------------------------------------------
X100 <- rd("maximin100"); y100 <- c(rd("c1n100_y")); th <- rd("c1n100_theta")
res <- t(apply(th, 1, function(t) { r <- A$logpost(X100, t, y100, 1.0)
  psi1 <- t[1]; psi2 <- t[2]; phi <- t[3]; zeta <- t[4]
  jac <- -phi - 2 * log(1 + exp(-phi)) + psi1 + psi2 + zeta
  pri <- -psi1 - psi1^2 / 2 - psi2 - psi2^2 / 2 - 4 * zeta - 4 / exp(zeta)
  c(r$val, r$beta, r$val - jac - pri) }))
wr(res, "A_logpost_c1n100")            # columns: val, beta, log.like

# ---- [I] logpost's log.like on the n = 14 iso rows --------------------------------------------------
X14 <- rd("maximin14"); y14 <- c(rd("c1n14_y")); nat <- rd("c1n14_nat")
res <- t(apply(nat, 1, function(q) { t <- c(log(q[2]), log(q[3]), log(q[1] / (1 - q[1])))
  r <- I$logpost(X14, t, y14, 0.7)
  jac <- -t[3] - 2 * log(1 + exp(-t[3])) + t[1] + t[2]
  pri <- -4 * t[1] - 2 / q[2] - 6 * t[2] - 16 / q[3]
  c(r$val - jac - pri, r$beta) }))
wr(res, "I_loglike_c1n14gls")          # columns: log.like, beta

# ---- [V] likeli.hyperpars (N = 1728, tau = 100 are hard-coded at [V]:555-556) ---------------------------
hp <- rd("hyperpars_2d"); rows <- c(rd("likeli_rows")) + 1
res <- sapply(rows, function(i) V$likeli.hyperpars(X14, y14, hp[i, 1:2], hp[i, 3:4], 0.7))
wr(matrix(res, ncol = 1), "V_likeli_hyperpars")

# ---- [A] predict.post over the pred14 table (pars rows built with the script's own logpost + factors) ---
pars <- rd("pred14_pars"); Xnew <- rd("pred14_Xnew"); y14p <- c(rd("pred14_y"))
tab.mean <- matrix(0, nrow(Xnew), nrow(pars)); tab.var <- tab.mean
for (s in seq_len(nrow(pars))) {
  q <- pars[s, ]
  t <- c(log(q[2]), log(q[3]), log(q[1] / (1 - q[1])), log(q[4]))
  r <- A$logpost(X14, t, y14p, 0.9)
  f <- A$factors(c(c(r$R.Inv), r$beta), 14, y14p)
  row <- c(q, r$beta, f, c(r$R.Inv))
  for (i in seq_len(nrow(Xnew))) {
    mv <- A$predict.post(Xnew[i, ], X14, row, 0.9)
    tab.mean[i, s] <- mv[1]; tab.var[i, s] <- mv[2]
  }
}
wr(tab.mean, "A_pred14_mean"); wr(tab.var, "A_pred14_var")

# ---- [M] Augmented.Mixed.Entropy over pool[1:200] x params; Entropy of the initial design -------------
D.old <- rd("me_initial14"); pool <- rd("me_pool200"); prm <- rd("me_params")
nd <- matrix(0, 200, nrow(prm))
for (q in seq_len(nrow(prm))) {
  p <- prm[q, 1]; t1 <- prm[q, 2]; t2 <- prm[q, 3]
  R.old <- (p^2 * M$corr.matrix.ISO(D.old, t1) + (1 - p)^2 * M$corr.matrix.ISO(D.old, t2)) / (p^2 + (1 - p)^2)
  R.old.Inv <- solve(R.old, tol = 1e-16)
  for (c in 1:200) nd[c, q] <- M$Augmented.Mixed.Entropy(D.old, matrix(pool[c, ], ncol = 2), p, t1, t2, R.old.Inv)
}
wr(nd, "M_negdet_200")
wr(matrix(apply(prm, 1, function(q) M$Entropy(D.old, q[1], q[2], q[3])), ncol = 1), "M_entropy_initial14")

# ---- optional: time the reference logpost at n = 100 (bench.py's cpu_baseline kind "reference") --------
if (do.time) {
  reps <- min(nrow(th), 48)
  t0 <- proc.time()[["elapsed"]]
  for (b in seq_len(reps)) A$logpost(X100, th[b, ], y100, 1.0)
  dt <- proc.time()[["elapsed"]] - t0
  si <- sessionInfo()
  writeLines(c(sprintf("evals,%d", reps), sprintf("seconds,%.6f", dt), sprintf("BLAS,%s", si$BLAS), sprintf("R,%s", R.version.string)),
             file.path(out.dir, "timing.csv"))
}
cat("r_crosscheck: done\n")
