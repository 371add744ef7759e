/* ccgp_shim.c -- .Call entry points that bind the reference R scripts to libccgp.so.
 *
 * Build (where R exists):
 *     R CMD SHLIB ccgp_shim.c -I../include -L../convex-combination-of-gaussian-processes_b200/lib -lccgp
 * R is NOT in the build container: there the CPU test suite compiles this file with
 * `gcc -fsyntax-only -Wall -Wextra -Werror` against the stand-in headers under tests/r_stub/
 * (tests/test_r_boundary.py), which pins every libccgp call to the prototypes of include/ccgp.h.
 * Every routine is registered (R_init_ccgp_shim at the end of this file), so `.Call("ccgp_R_...")`
 * resolves through the registration table, with the argument count checked by R.
 *
 * Conventions: REALSXP matrices are column-major doubles and are handed to libccgp as they
 * are (no copy); INTSXP is int32.  The context handle is an external pointer with a C
 * finalizer.  libccgp never calls back into R and returns an int status, so Rf_error() is
 * only raised here, after every temporary has been PROTECTed/released (Rf_error longjmps).
 * Per-candidate numerical failure is not an error: the value comes back NaN with status 1
 * and r/ccgp.R turns it into NA, reproducing `try(solve(R))` -> NA at [A]:448-449.
 */
#include <R.h>
#include <Rinternals.h>
#include <R_ext/Rdynload.h>
#include <stdint.h>
#include "ccgp.h"

static void ctx_finalizer(SEXP ptr) {
    ccgp_ctx* ctx = (ccgp_ctx*)R_ExternalPtrAddr(ptr);
    if (ctx) { ccgp_destroy(ctx); R_ClearExternalPtr(ptr); }
}

static ccgp_ctx* get_ctx(SEXP ptr) {
    ccgp_ctx* ctx = (ccgp_ctx*)R_ExternalPtrAddr(ptr);
    if (!ctx) Rf_error("ccgp: context already destroyed");
    return ctx;
}

static void check(ccgp_ctx* ctx, int rc, const char* what) {
    if (rc != CCGP_OK) Rf_error("ccgp %s failed (%d): %s", what, rc, ccgp_last_error(ctx));
}

static SEXP wrap_ctx(ccgp_ctx* ctx) {
    SEXP ptr = PROTECT(R_MakeExternalPtr(ctx, R_NilValue, R_NilValue));
    R_RegisterCFinalizerEx(ptr, ctx_finalizer, TRUE);
    UNPROTECT(1);
    return ptr;
}

SEXP ccgp_R_create(SEXP device) {
    ccgp_ctx* ctx = NULL;
    int rc = ccgp_create(&ctx, Rf_asInteger(device));
    if (rc != CCGP_OK) Rf_error("ccgp_create failed (%d): %s", rc, ccgp_last_error(NULL));
    return wrap_ctx(ctx);
}

/* all (n_gpus <= 0) or the first n_gpus GPUs of the box behind one context */
SEXP ccgp_R_create_multi(SEXP n_gpus) {
    ccgp_ctx* ctx = NULL;
    int rc = ccgp_create_multi(&ctx, Rf_asInteger(n_gpus));
    if (rc != CCGP_OK) Rf_error("ccgp_create_multi failed (%d): %s", rc, ccgp_last_error(NULL));
    return wrap_ctx(ctx);
}

SEXP ccgp_R_num_gpus(SEXP ptr) {
    SEXP out = PROTECT(Rf_allocVector(INTSXP, 1));
    INTEGER(out)[0] = ccgp_num_gpus(get_ctx(ptr));
    UNPROTECT(1);
    return out;
}

/* Matern smoothness nu of the 1-D families ([D1]:1080) */
SEXP ccgp_R_set_matern_nu(SEXP ptr, SEXP nu) {
    ccgp_ctx* ctx = get_ctx(ptr);
    check(ctx, ccgp_set_matern_nu(ctx, Rf_asReal(nu)), "set_matern_nu");
    return R_NilValue;
}

/* D.train (n x d), y (n) */
SEXP ccgp_R_set_design(SEXP ptr, SEXP X, SEXP y) {
    ccgp_ctx* ctx = get_ctx(ptr);
    if (!Rf_isMatrix(X) || !Rf_isReal(X) || !Rf_isReal(y)) Rf_error("ccgp: D.train must be a double matrix, y a double vector");
    int n = Rf_nrows(X), d = Rf_ncols(X);
    if (XLENGTH(y) != n) Rf_error("ccgp: length(y) != nrow(D.train)");
    check(ctx, ccgp_set_design(ctx, REAL(X), n, d, REAL(y)), "set_design");
    return R_NilValue;
}

/* cand: B x k matrix.  Returns list(nll, beta, status). */
SEXP ccgp_R_nll_batch(SEXP ptr, SEXP family, SEXP scale, SEXP cand, SEXP sigma2, SEXP mean_mode, SEXP tau) {
    ccgp_ctx* ctx = get_ctx(ptr);
    if (!Rf_isMatrix(cand) || !Rf_isReal(cand)) Rf_error("ccgp: candidates must be a double matrix");
    int64_t B = Rf_nrows(cand);
    SEXP nll = PROTECT(Rf_allocVector(REALSXP, B));
    SEXP beta = PROTECT(Rf_allocVector(REALSXP, B));
    SEXP status = PROTECT(Rf_allocVector(INTSXP, B));
    int rc = ccgp_nll_batch(ctx, Rf_asInteger(family), Rf_asInteger(scale), REAL(cand), B, B, Rf_asReal(sigma2),
                            Rf_asInteger(mean_mode), Rf_asReal(tau), REAL(nll), REAL(beta), (int32_t*)INTEGER(status));
    SEXP out = PROTECT(Rf_allocVector(VECSXP, 3));
    SET_VECTOR_ELT(out, 0, nll); SET_VECTOR_ELT(out, 1, beta); SET_VECTOR_ELT(out, 2, status);
    UNPROTECT(4);
    check(ctx, rc, "nll_batch");
    return out;
}

/* Returns list(Rinv (n x n x B array), beta, status). */
SEXP ccgp_R_rinv_batch(SEXP ptr, SEXP family, SEXP scale, SEXP cand, SEXP n_) {
    ccgp_ctx* ctx = get_ctx(ptr);
    int64_t B = Rf_nrows(cand);
    int n = Rf_asInteger(n_);
    SEXP rinv = PROTECT(Rf_allocVector(REALSXP, (R_xlen_t)n * n * B));
    SEXP beta = PROTECT(Rf_allocVector(REALSXP, B));
    SEXP status = PROTECT(Rf_allocVector(INTSXP, B));
    int rc = ccgp_rinv_batch(ctx, Rf_asInteger(family), Rf_asInteger(scale), REAL(cand), B, B, REAL(rinv), REAL(beta),
                             (int32_t*)INTEGER(status));
    SEXP out = PROTECT(Rf_allocVector(VECSXP, 3));
    SET_VECTOR_ELT(out, 0, rinv); SET_VECTOR_ELT(out, 1, beta); SET_VECTOR_ELT(out, 2, status);
    UNPROTECT(4);
    check(ctx, rc, "rinv_batch");
    return out;
}

/* pars: S x k, pars_vec: S x k' or NULL, Xnew: T x d.  Returns list(mean (T x S), var (T x S), status). */
SEXP ccgp_R_predict(SEXP ptr, SEXP family, SEXP pars, SEXP vec_family, SEXP pars_vec, SEXP Xnew, SEXP sigma2) {
    ccgp_ctx* ctx = get_ctx(ptr);
    int64_t S = Rf_nrows(pars), T = Rf_nrows(Xnew);
    SEXP mean = PROTECT(Rf_allocMatrix(REALSXP, (int)T, (int)S));
    SEXP var = PROTECT(Rf_allocMatrix(REALSXP, (int)T, (int)S));
    SEXP status = PROTECT(Rf_allocVector(INTSXP, S));
    const double* pv = Rf_isNull(pars_vec) ? NULL : REAL(pars_vec);
    int rc = ccgp_predict(ctx, Rf_asInteger(family), REAL(pars), S, S, Rf_asInteger(vec_family), pv, S, REAL(Xnew), T,
                          Rf_asReal(sigma2), REAL(mean), REAL(var), (int32_t*)INTEGER(status));
    SEXP out = PROTECT(Rf_allocVector(VECSXP, 3));
    SET_VECTOR_ELT(out, 0, mean); SET_VECTOR_ELT(out, 1, var); SET_VECTOR_ELT(out, 2, status);
    UNPROTECT(4);
    check(ctx, rc, "predict");
    return out;
}

/* ---- factors kept on the device: factors.frame [A]:572-592 without its wire format ----
 * The handle is an external pointer that keeps the context's external pointer alive in its `prot` slot, so the
 * context cannot be finalized first; its own finalizer releases the HBM. */
static void factors_finalizer(SEXP ptr) {
    ccgp_factors* f = (ccgp_factors*)R_ExternalPtrAddr(ptr);
    SEXP cptr = R_ExternalPtrProtected(ptr);
    if (f && cptr != R_NilValue) {
        ccgp_ctx* ctx = (ccgp_ctx*)R_ExternalPtrAddr(cptr);
        if (ctx) ccgp_factors_destroy(ctx, f);
    }
    R_ClearExternalPtr(ptr);
}

/* pars: S x k, pars_vec: S x k' or NULL.  Returns the factors handle. */
SEXP ccgp_R_factors_create(SEXP ptr, SEXP family, SEXP pars, SEXP vec_family, SEXP pars_vec) {
    ccgp_ctx* ctx = get_ctx(ptr);
    int64_t S = Rf_nrows(pars);
    const double* pv = Rf_isNull(pars_vec) ? NULL : REAL(pars_vec);
    ccgp_factors* f = NULL;
    int rc = ccgp_factors_create(ctx, Rf_asInteger(family), REAL(pars), S, S, Rf_asInteger(vec_family), pv, S, &f);
    check(ctx, rc, "factors_create");
    SEXP out = PROTECT(R_MakeExternalPtr(f, R_NilValue, ptr));
    R_RegisterCFinalizerEx(out, factors_finalizer, TRUE);
    UNPROTECT(1);
    return out;
}

/* Xnew: T x d.  Returns list(mean (T x S), var (T x S), status) from the stored factors. */
SEXP ccgp_R_factors_predict(SEXP ptr, SEXP fptr, SEXP Xnew, SEXP sigma2) {
    ccgp_ctx* ctx = get_ctx(ptr);
    ccgp_factors* f = (ccgp_factors*)R_ExternalPtrAddr(fptr);
    if (!f) Rf_error("ccgp: factors already released");
    int64_t S = 0, T = Rf_nrows(Xnew);
    check(ctx, ccgp_factors_info(ctx, f, &S, NULL, NULL), "factors_info");
    SEXP mean = PROTECT(Rf_allocMatrix(REALSXP, (int)T, (int)S));
    SEXP var = PROTECT(Rf_allocMatrix(REALSXP, (int)T, (int)S));
    SEXP status = PROTECT(Rf_allocVector(INTSXP, S));
    int rc = ccgp_factors_predict(ctx, f, REAL(Xnew), T, Rf_asReal(sigma2), REAL(mean), REAL(var), (int32_t*)INTEGER(status));
    SEXP out = PROTECT(Rf_allocVector(VECSXP, 3));
    SET_VECTOR_ELT(out, 0, mean); SET_VECTOR_ELT(out, 1, var); SET_VECTOR_ELT(out, 2, status);
    UNPROTECT(4);
    check(ctx, rc, "factors_predict");
    return out;
}

SEXP ccgp_R_factors_release(SEXP fptr) {
    factors_finalizer(fptr);
    return R_NilValue;
}

/* D.old: n_old x d (or NULL), D.new: (n_new*d) x C matrix whose columns are c(D.new) vectors,
 * params: P x 3.  Returns list(negdet (C x P), logdet (C x P), status (C x P)). */
SEXP ccgp_R_me_schur_batch(SEXP ptr, SEXP D_old, SEXP D_new, SEXP n_new_, SEXP d_, SEXP params) {
    ccgp_ctx* ctx = get_ctx(ptr);
    int n_new = Rf_asInteger(n_new_), d = Rf_asInteger(d_);
    int n_old = Rf_isNull(D_old) ? 0 : Rf_nrows(D_old);
    int64_t C = Rf_ncols(D_new), P = Rf_nrows(params);
    SEXP negdet = PROTECT(Rf_allocMatrix(REALSXP, (int)C, (int)P));
    SEXP logdet = PROTECT(Rf_allocMatrix(REALSXP, (int)C, (int)P));
    SEXP status = PROTECT(Rf_allocMatrix(INTSXP, (int)C, (int)P));
    int rc = ccgp_me_schur_batch(ctx, n_old ? REAL(D_old) : NULL, n_old, d, REAL(D_new), n_new, C, REAL(params), P, P,
                                 REAL(negdet), REAL(logdet), (int32_t*)INTEGER(status));
    SEXP out = PROTECT(Rf_allocVector(VECSXP, 3));
    SET_VECTOR_ELT(out, 0, negdet); SET_VECTOR_ELT(out, 1, logdet); SET_VECTOR_ELT(out, 2, status);
    UNPROTECT(4);
    check(ctx, rc, "me_schur_batch");
    return out;
}

/* Mixed correlation block between the rows of A and of B (B = NULL: A with itself). */
SEXP ccgp_R_mixed_corr(SEXP ptr, SEXP family, SEXP params, SEXP A, SEXP B) {
    ccgp_ctx* ctx = get_ctx(ptr);
    int na = Rf_nrows(A), d = Rf_ncols(A);
    int nb = Rf_isNull(B) ? na : Rf_nrows(B);
    SEXP out = PROTECT(Rf_allocMatrix(REALSXP, na, nb));
    int rc = ccgp_mixed_corr(ctx, Rf_asInteger(family), REAL(params), REAL(A), na, Rf_isNull(B) ? NULL : REAL(B), nb, d, REAL(out));
    UNPROTECT(1);
    check(ctx, rc, "mixed_corr");
    return out;
}

/* PAM k-medoids of the rows of P (n x d).  Returns list(medoid rows (1-based), cost, swaps). */
SEXP ccgp_R_kmedoids_pam(SEXP ptr, SEXP P, SEXP k_, SEXP max_swaps) {
    ccgp_ctx* ctx = get_ctx(ptr);
    int k = Rf_asInteger(k_);
    SEXP med = PROTECT(Rf_allocVector(INTSXP, k));
    SEXP cost = PROTECT(Rf_allocVector(REALSXP, 1));
    SEXP swaps = PROTECT(Rf_allocVector(INTSXP, 1));
    int rc = ccgp_kmedoids_pam(ctx, REAL(P), (int64_t)Rf_nrows(P), Rf_ncols(P), k, Rf_asInteger(max_swaps),
                               (int32_t*)INTEGER(med), REAL(cost), (int32_t*)INTEGER(swaps));
    for (int i = 0; i < k; ++i) INTEGER(med)[i] += 1;
    SEXP out = PROTECT(Rf_allocVector(VECSXP, 3));
    SET_VECTOR_ELT(out, 0, med); SET_VECTOR_ELT(out, 1, cost); SET_VECTOR_ELT(out, 2, swaps);
    UNPROTECT(4);
    check(ctx, rc, "kmedoids_pam");
    return out;
}

/* Paired ME criterion: D.new is (n_new*d) x (P*group), column c belongs to parameter row c %/% group.
 * Returns list(negdet (P*group), status). */
SEXP ccgp_R_me_schur_paired(SEXP ptr, SEXP D_old, SEXP D_new, SEXP n_new_, SEXP d_, SEXP params, SEXP group_) {
    ccgp_ctx* ctx = get_ctx(ptr);
    int n_new = Rf_asInteger(n_new_), d = Rf_asInteger(d_);
    int n_old = Rf_isNull(D_old) ? 0 : Rf_nrows(D_old);
    int64_t C = Rf_ncols(D_new), P = Rf_nrows(params), group = Rf_asInteger(group_);
    if (C != P * group) Rf_error("ccgp: need nrow(params) * group designs");
    SEXP negdet = PROTECT(Rf_allocVector(REALSXP, (R_xlen_t)C));
    SEXP status = PROTECT(Rf_allocVector(INTSXP, (R_xlen_t)C));
    int rc = ccgp_me_schur_paired(ctx, n_old ? REAL(D_old) : NULL, n_old, d, REAL(D_new), n_new, group, REAL(params), P, P,
                                  REAL(negdet), (int32_t*)INTEGER(status));
    SEXP out = PROTECT(Rf_allocVector(VECSXP, 2));
    SET_VECTOR_ELT(out, 0, negdet); SET_VECTOR_ELT(out, 1, status);
    UNPROTECT(3);
    check(ctx, rc, "me_schur_paired");
    return out;
}

/* Central-difference stencil of the ME criterion, generated on the device: X is (n_new*d) x (P*group) (columns =
 * c(D.new) of each problem).  Returns the (2m+1) x (P*group) matrix of criterion values (row 1: base point). */
SEXP ccgp_R_me_schur_stencil(SEXP ptr, SEXP D_old, SEXP X, SEXP n_new_, SEXP d_, SEXP params, SEXP group_, SEXP h_, SEXP lo_, SEXP hi_) {
    ccgp_ctx* ctx = get_ctx(ptr);
    int n_new = Rf_asInteger(n_new_), d = Rf_asInteger(d_);
    int n_old = Rf_isNull(D_old) ? 0 : Rf_nrows(D_old);
    int64_t K = Rf_ncols(X), P = Rf_nrows(params), group = Rf_asInteger(group_);
    if (K != P * group) Rf_error("ccgp: need nrow(params) * group problems");
    SEXP vals = PROTECT(Rf_allocMatrix(REALSXP, 2 * n_new * d + 1, (int)K));
    int rc = ccgp_me_schur_stencil(ctx, n_old ? REAL(D_old) : NULL, n_old, d, REAL(X), n_new, group, REAL(params), P, P,
                                   Rf_asReal(h_), Rf_asReal(lo_), Rf_asReal(hi_), REAL(vals), NULL);
    UNPROTECT(1);
    check(ctx, rc, "me_schur_stencil");
    return vals;
}

/* which.min of the batched likelihood (choose.hyperpars' argmax over a sweep is the argmin of -loglik):
 * returns list(best nll, 1-based index).  On a multi-GPU context the reduction is the NCCL all-reduce. */
SEXP ccgp_R_nll_argmin(SEXP ptr, SEXP family, SEXP scale, SEXP cand, SEXP sigma2, SEXP mean_mode, SEXP tau) {
    ccgp_ctx* ctx = get_ctx(ptr);
    if (!Rf_isMatrix(cand) || !Rf_isReal(cand)) Rf_error("ccgp: candidates must be a double matrix");
    int64_t B = Rf_nrows(cand), idx = -1;
    double best = 0.0;
    int rc = ccgp_nll_argmin(ctx, Rf_asInteger(family), Rf_asInteger(scale), REAL(cand), B, B, Rf_asReal(sigma2),
                             Rf_asInteger(mean_mode), Rf_asReal(tau), &best, &idx);
    check(ctx, rc, "nll_argmin");
    SEXP v = PROTECT(Rf_allocVector(REALSXP, 1));
    SEXP i = PROTECT(Rf_allocVector(REALSXP, 1));        /* double: B may exceed 2^31 */
    REAL(v)[0] = best; REAL(i)[0] = (double)(idx + 1);
    SEXP out = PROTECT(Rf_allocVector(VECSXP, 2));
    SET_VECTOR_ELT(out, 0, v); SET_VECTOR_ELT(out, 1, i);
    UNPROTECT(3);
    return out;
}

/* rcond_1(R) per candidate: the number solve(R) tests against .Machine$double.eps ([A]:448-449).
 * Returns list(rcond, beta, status). */
SEXP ccgp_R_rcond_batch(SEXP ptr, SEXP family, SEXP scale, SEXP cand) {
    ccgp_ctx* ctx = get_ctx(ptr);
    if (!Rf_isMatrix(cand) || !Rf_isReal(cand)) Rf_error("ccgp: candidates must be a double matrix");
    int64_t B = Rf_nrows(cand);
    SEXP rcond = PROTECT(Rf_allocVector(REALSXP, B));
    SEXP beta = PROTECT(Rf_allocVector(REALSXP, B));
    SEXP status = PROTECT(Rf_allocVector(INTSXP, B));
    int rc = ccgp_rcond_batch(ctx, Rf_asInteger(family), Rf_asInteger(scale), REAL(cand), B, B, REAL(rcond), REAL(beta),
                              (int32_t*)INTEGER(status));
    SEXP out = PROTECT(Rf_allocVector(VECSXP, 3));
    SET_VECTOR_ELT(out, 0, rcond); SET_VECTOR_ELT(out, 1, beta); SET_VECTOR_ELT(out, 2, status);
    UNPROTECT(4);
    check(ctx, rc, "rcond_batch");
    return out;
}

/* which.min over the candidate designs of each parameter row ([M]:944-945): returns list(best -det (P), 1-based index (P)). */
SEXP ccgp_R_me_argmin(SEXP ptr, SEXP D_old, SEXP D_new, SEXP n_new_, SEXP d_, SEXP params) {
    ccgp_ctx* ctx = get_ctx(ptr);
    int n_new = Rf_asInteger(n_new_), d = Rf_asInteger(d_);
    int n_old = Rf_isNull(D_old) ? 0 : Rf_nrows(D_old);
    int64_t C = Rf_ncols(D_new), P = Rf_nrows(params);
    SEXP val = PROTECT(Rf_allocVector(REALSXP, P));
    SEXP idx = PROTECT(Rf_allocVector(REALSXP, P));
    int64_t* tmp = (int64_t*)R_alloc((size_t)(P > 0 ? P : 1), sizeof(int64_t));
    int rc = ccgp_me_argmin(ctx, n_old ? REAL(D_old) : NULL, n_old, d, REAL(D_new), n_new, C, REAL(params), P, P, REAL(val), tmp);
    if (rc == CCGP_OK) for (int64_t q = 0; q < P; ++q) REAL(idx)[q] = (double)(tmp[q] + 1);
    SEXP out = PROTECT(Rf_allocVector(VECSXP, 2));
    SET_VECTOR_ELT(out, 0, val); SET_VECTOR_ELT(out, 1, idx);
    UNPROTECT(3);
    check(ctx, rc, "me_argmin");
    return out;
}

/* log det R[S,S] for C index subsets (idx: C x m INTEGER matrix, 1-based) of a pool (N x d): returns list(logdet, status). */
SEXP ccgp_R_subset_logdet_batch(SEXP ptr, SEXP pool, SEXP idx, SEXP family, SEXP params) {
    ccgp_ctx* ctx = get_ctx(ptr);
    if (!Rf_isMatrix(pool) || !Rf_isReal(pool) || !Rf_isMatrix(idx) || !Rf_isInteger(idx)) Rf_error("ccgp: pool must be a double matrix, idx an integer matrix");
    int64_t N = Rf_nrows(pool), C = Rf_nrows(idx);
    int d = Rf_ncols(pool), m = Rf_ncols(idx);
    int32_t* zero = (int32_t*)R_alloc((size_t)(C * m > 0 ? C * m : 1), sizeof(int32_t));
    for (int64_t e = 0; e < C * m; ++e) zero[e] = INTEGER(idx)[e] - 1;
    SEXP logdet = PROTECT(Rf_allocVector(REALSXP, C));
    SEXP status = PROTECT(Rf_allocVector(INTSXP, C));
    int rc = ccgp_subset_logdet_batch(ctx, REAL(pool), N, d, zero, m, C, C, Rf_asInteger(family), REAL(params), REAL(logdet),
                                      (int32_t*)INTEGER(status));
    SEXP out = PROTECT(Rf_allocVector(VECSXP, 2));
    SET_VECTOR_ELT(out, 0, logdet); SET_VECTOR_ELT(out, 1, status);
    UNPROTECT(3);
    check(ctx, rc, "subset_logdet_batch");
    return out;
}

/* ---- registration --------------------------------------------------------------------------------------- */
#define CALLDEF(name, n) {#name, (DL_FUNC)&name, n}
/* CGP comparator ([A]:104-151, 166-199): objective of the start sweep / leave-one-out predictions */
SEXP ccgp_R_cgp_objective_batch(SEXP ptr, SEXP Xs, SEXP y, SEXP W) {
    ccgp_ctx* ctx = get_ctx(ptr);
    int n = Rf_nrows(Xs), p = Rf_ncols(Xs);
    int64_t B = Rf_nrows(W);
    if (Rf_ncols(W) != p + 3 || XLENGTH(y) != (R_xlen_t)n) Rf_error("ccgp: W needs ncol(Xs) + 3 columns and y nrow(Xs) entries");
    SEXP out = PROTECT(Rf_allocVector(REALSXP, B));
    int rc = ccgp_cgp_objective_batch(ctx, REAL(Xs), REAL(y), n, p, REAL(W), B, B, REAL(out), NULL);
    UNPROTECT(1);
    check(ctx, rc, "cgp_objective_batch");
    return out;
}
SEXP ccgp_R_cgp_jackknife(SEXP ptr, SEXP Xs, SEXP y, SEXP w) {
    ccgp_ctx* ctx = get_ctx(ptr);
    int n = Rf_nrows(Xs), p = Rf_ncols(Xs);
    if (XLENGTH(w) != (R_xlen_t)(p + 3) || XLENGTH(y) != (R_xlen_t)n) Rf_error("ccgp: w needs ncol(Xs) + 3 entries and y nrow(Xs) entries");
    SEXP out = PROTECT(Rf_allocVector(REALSXP, n));
    int rc = ccgp_cgp_jackknife(ctx, REAL(Xs), REAL(y), n, p, REAL(w), REAL(out), NULL);
    UNPROTECT(1);
    check(ctx, rc, "cgp_jackknife");
    return out;
}

static const R_CallMethodDef ccgp_call_methods[] = {
    CALLDEF(ccgp_R_create, 1),
    CALLDEF(ccgp_R_create_multi, 1),
    CALLDEF(ccgp_R_num_gpus, 1),
    CALLDEF(ccgp_R_set_matern_nu, 2),
    CALLDEF(ccgp_R_set_design, 3),
    CALLDEF(ccgp_R_nll_batch, 7),
    CALLDEF(ccgp_R_nll_argmin, 7),
    CALLDEF(ccgp_R_rinv_batch, 5),
    CALLDEF(ccgp_R_rcond_batch, 4),
    CALLDEF(ccgp_R_predict, 7),
    CALLDEF(ccgp_R_factors_create, 5),
    CALLDEF(ccgp_R_factors_predict, 4),
    CALLDEF(ccgp_R_factors_release, 1),
    CALLDEF(ccgp_R_me_schur_batch, 6),
    CALLDEF(ccgp_R_me_argmin, 6),
    CALLDEF(ccgp_R_me_schur_paired, 7),
    CALLDEF(ccgp_R_me_schur_stencil, 10),
    CALLDEF(ccgp_R_subset_logdet_batch, 5),
    CALLDEF(ccgp_R_mixed_corr, 5),
    CALLDEF(ccgp_R_kmedoids_pam, 4),
    CALLDEF(ccgp_R_cgp_objective_batch, 4),
    CALLDEF(ccgp_R_cgp_jackknife, 4),
    {NULL, NULL, 0}
};

void R_init_ccgp_shim(DllInfo* dll) {
    R_registerRoutines(dll, NULL, ccgp_call_methods, NULL, NULL);
    R_useDynamicSymbols(dll, FALSE);
}
