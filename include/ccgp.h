/* ccgp.h -- plain C ABI of libccgp.so, the B200 (sm_100a) engine for the
 * data-parallel core of the Combined-GP reference R scripts
 * (oharari/Convex-Combination-of-Gaussian-Processes).
 *
 * The reference has no FFI of its own: its hot path is ordinary R closures.
 * Each entry point below therefore replaces one (batched) reference closure;
 * the citation names the closure a maintainer would re-point at it through
 * `.Call` (shim: r/ccgp_shim.c, wrappers: r/ccgp.R).
 * File aliases (under the reference root):
 *   [A] 2D Codes and Designs/2D Combined GP Anisotropic Public.R
 *   [I] 2D Codes and Designs/2D Combined GP Isotropic Public.R
 *   [V] 2D Codes and Designs/2D Combined GP Isotropic Advanced.R
 *   [M] Batch Sequential ME Designs/Batch Sequential ME Design.R
 *   [H] Heat Exchanger Emulator/Combined GP Heat Exchanger.R
 *
 * Conventions: every matrix is COLUMN-MAJOR double (R's layout), every
 * pointer is caller-owned; `*_dev` variants take device pointers on the
 * context's device and enqueue on the context's stream without synchronising
 * (call ccgp_sync), the plain variants take HOST pointers, copy in/out and
 * return when the results are in the caller's buffers.  All functions return
 * 0 or a negative ccgp_status; ccgp_last_error() gives the message.  A context
 * is bound to one GPU and is not thread-safe; use one context per thread/GPU.
 * There is no CPU fallback: without a usable CUDA device ccgp_create fails.
 * Per-candidate numerical failure is NOT an error: the value is NaN and the
 * candidate's entry in out_status is 1 (matrix not positive definite; the R
 * wrapper maps it to NA like `try(solve(R))` at [A]:448-449).
 */
#ifndef CCGP_H
#define CCGP_H
#include <stdint.h>
#ifdef __cplusplus
extern "C" {
#endif

typedef struct ccgp_ctx ccgp_ctx;

enum ccgp_status {
    CCGP_OK = 0,
    CCGP_ERR_ARG = -1,         /* bad argument (sizes, NULLs, unknown enum) */
    CCGP_ERR_CUDA = -2,        /* CUDA runtime error (message has the detail) */
    CCGP_ERR_UNSUPPORTED = -3, /* valid request this build does not cover */
    CCGP_ERR_STATE = -4        /* call order (no design set, ...) */
};

/* Component-kernel family = which Mixed.corr.matrix the candidate row feeds.
 * Natural-scale candidate rows (ccgp_scale NATURAL):
 *   GAUSS_ISO          (p, theta1, theta2)            [I]:400-407, [M]:397-404
 *   GAUSS_ANISO_LAMBDA (p, theta_1..theta_d, lambda)  [A]:399-406 (2nd component (1+lambda)*theta)
 *   GAUSS_ISO_RAW2     (p, theta1, lambda)            [V]:414-421 (2nd component scale = lambda)
 * LOGSCALE rows are logpost's real-line vectors ([A]:435-442, [I]:435-440):
 *   iso (psi1, psi2, phi), aniso (psi_1..psi_d, phi, zeta); the kernel applies
 *   theta=exp(psi), p=1/(1+exp(-phi)), lambda=exp(zeta).
 *   MATERN1D           (p, theta1, theta2)  1-D only: both components Matern(nu)
 *                      "1D Combined GP Public.R":348-374, 577-584 (Matern.corr.func, Mixed.corr.matrix)
 *   MATERN_SPLINE1D    (p, theta1, theta2)  1-D only: Matern(nu, theta1) + cubic spline (support theta2)
 *                      "1D Combined GP Two Families Public.R":346-369, 454-462 (corr.matrix.combined)
 *   (real-line rows as for GAUSS_ISO; nu via ccgp_set_matern_nu, default 5 as at [D1]:1080) */
enum ccgp_family { CCGP_GAUSS_ISO = 0, CCGP_GAUSS_ANISO_LAMBDA = 1, CCGP_GAUSS_ISO_RAW2 = 2,
                   CCGP_MATERN1D = 3, CCGP_MATERN_SPLINE1D = 4 };
enum ccgp_scale { CCGP_NATURAL = 0, CCGP_LOGSCALE = 1 };
/* GLS_BETA: dmnorm(y, beta.MLE, c R) as in logpost [A]:452-455.
 * ZERO_PLUS_TAU2: dmnorm(y, 0, c R + tau^2 11') as in cond.like [V]:564-575. */
enum ccgp_mean_mode { CCGP_MEAN_GLS_BETA = 0, CCGP_MEAN_ZERO_PLUS_TAU2 = 1 };

/* number of columns of a candidate row for (family, d) */
int ccgp_num_params(int family, int d);

/* ---- context ------------------------------------------------------------ */
int ccgp_create(ccgp_ctx** ctx, int device);
int ccgp_destroy(ccgp_ctx* ctx);
/* All GPUs of the box behind one context, for a single-threaded caller such as R's .Call:
 * n_gpus <= 0 takes every visible device.  ccgp_set_design broadcasts the design; ccgp_nll_batch,
 * ccgp_nll_argmin, ccgp_predict, ccgp_me_schur_batch and ccgp_me_argmin split their batch into contiguous
 * slices (candidates / posterior rows / parameter rows or designs), one GPU and one host thread per slice,
 * each GPU writing its slice of the caller's output; per-candidate results are bit-identical for every GPU
 * count.  which.min ([V]:598, [M]:944-945) is reduced with two NCCL all-reduces (MIN over the values, then MIN
 * over the indices of the winners: lowest index on ties) on communicators from ncclCommInitAll; libnccl.so.2
 * is dlopen()ed here (override the path with CCGP_NCCL_LIB), single-GPU contexts never touch it.  Every other
 * entry point runs on GPU 0 of the set.  The `_dev` variants are single-GPU only. */
int ccgp_create_multi(ccgp_ctx** ctx, int n_gpus);
int ccgp_num_gpus(const ccgp_ctx* ctx);
/* NCCL collectives issued so far by a multi-GPU context (0 for a single-GPU one) */
int64_t ccgp_collective_count(const ccgp_ctx* ctx);
const char* ccgp_last_error(const ccgp_ctx* ctx); /* ctx may be NULL: last create error */
int ccgp_sync(ccgp_ctx* ctx);
/* Matern smoothness nu of the 1-D families (integer or half-integer, 0 < nu <= 50; default 5) */
int ccgp_set_matern_nu(ccgp_ctx* ctx, double nu);
/* run on the caller's CUDA stream (a cudaStream_t; NULL = the legacy default stream)
 * so the caller's events bracket the kernels; ccgp_use_own_stream goes back to the
 * context's private non-blocking stream */
int ccgp_set_stream(ccgp_ctx* ctx, void* stream);
int ccgp_use_own_stream(ccgp_ctx* ctx);
int ccgp_device(const ccgp_ctx* ctx);
/* number of kernels this context has launched so far (bench.py's gpu_launches) */
int64_t ccgp_launch_count(const ccgp_ctx* ctx);
/* last kernel configuration chosen by ccgp_nll_batch (threads per candidate,
 * dynamic shared bytes, resident CTAs per SM, variant id) -- for reports */
int ccgp_last_nll_config(const ccgp_ctx* ctx, int* team, int* smem_bytes, int* ctas_per_sm, int* variant);
/* measured FP64 FMA throughput of this GPU, FLOP/s, from a dependent-free DFMA
 * loop on every SM (the roofline denominator bench.py reports against) */
int ccgp_measure_fp64_peak(ccgp_ctx* ctx, double* flops_per_s);
/* the same pipe driven in its tensor form (mma.sync.m8n8k4.f64 with independent accumulators): the
 * GEMM-shaped cross-check of the DFMA number; bench.py reports both in roofline.peak_detail */
int ccgp_measure_fp64_peak_dmma(ccgp_ctx* ctx, double* flops_per_s);

/* ---- training design (D.train, y) shared by every candidate -------------- */
/* X is n x d column-major, y has n entries.  Replaces the (D.train, y)
 * arguments every logpost / cond.like / predict.post call receives. */
int ccgp_set_design(ccgp_ctx* ctx, const double* X, int n, int d, const double* y);

/* ---- batched likelihood: logpost's log.like [A]:444-455 / cond.like [V]:564-575
 * cand is B x k column-major with leading dimension ldc (>= B).
 * out_nll[b]  = -log.like of candidate b  (NaN when status 1)
 * out_beta[b] = beta.MLE [A]:385-389 (may be NULL)
 * out_status[b] in {0 ok, 1 not positive definite} (may be NULL) */
int ccgp_nll_batch(ccgp_ctx* ctx, int family, int scale, const double* cand, int64_t B, int64_t ldc,
                   double sigma2, int mean_mode, double tau,
                   double* out_nll, double* out_beta, int32_t* out_status);
int ccgp_nll_batch_dev(ccgp_ctx* ctx, int family, int scale, const double* d_cand, int64_t B, int64_t ldc,
                       double sigma2, int mean_mode, double tau,
                       double* d_nll, double* d_beta, int32_t* d_status);
/* argmin over a device vector with lowest-index tie-break (which.min); NaNs are
 * skipped.  Used for the likelihood argmin / entropy argmax ([V]:598, [M]:944). */
int ccgp_argmin_dev(ccgp_ctx* ctx, const double* d_vals, int64_t B, double* best_val, int64_t* best_idx);
int ccgp_nll_argmin(ccgp_ctx* ctx, int family, int scale, const double* cand, int64_t B, int64_t ldc,
                    double sigma2, int mean_mode, double tau, double* best_nll, int64_t* best_idx);

/* ---- R.Inv + beta for a few candidates: logpost's list(beta, R.Inv) [A]:448-466
 * out_Rinv is n*n*B (each n x n column-major), out_beta B. */
int ccgp_rinv_batch(ccgp_ctx* ctx, int family, int scale, const double* cand, int64_t B, int64_t ldc,
                    double* out_Rinv, double* out_beta, int32_t* out_status);
/* rcond_1(R) = 1 / (||R||_1 ||R^-1||_1) per candidate (both norms exact, from the explicit inverse):
 * the number base R's solve() compares with .Machine$double.eps before logpost's `try(solve(R))`
 * turns into NA ([A]:448-449; R takes LAPACK dgecon's ESTIMATE of it).  The R / Python `logpost`
 * wrappers return NA when status != 0 or rcond < 2.220446e-16.  out_rcond[b] = 0 when the
 * factorisation itself broke down (status 1).  out_beta / out_status may be NULL. */
int ccgp_rcond_batch(ccgp_ctx* ctx, int family, int scale, const double* cand, int64_t B, int64_t ldc,
                     double* out_rcond, double* out_beta, int32_t* out_status);

/* ---- prediction table: predict.post [A]:604-623 over S posterior rows x T sites
 * pars is S x k natural-scale (family); Xnew is T x d column-major.
 * vec_family/pars_vec (S x k', may be -1/NULL = same as family/pars) give the
 * parameters of the correlation VECTOR r(x) when they differ from the matrix's
 * (quirk at [V]:672).  out_mean/out_var are T x S column-major.
 * var uses sigma2 WITHOUT (p^2+(1-p)^2), exactly as [A]:619. */
int ccgp_predict(ccgp_ctx* ctx, int family, const double* pars, int64_t S, int64_t ldp,
                 int vec_family, const double* pars_vec, int64_t ldpv,
                 const double* Xnew, int64_t T, double sigma2,
                 double* out_mean, double* out_var, int32_t* out_status);
int ccgp_predict_dev(ccgp_ctx* ctx, int family, const double* d_pars, int64_t S, int64_t ldp,
                     int vec_family, const double* d_pars_vec, int64_t ldpv,
                     const double* d_Xnew, int64_t T, double sigma2,
                     double* d_mean, double* d_var, int32_t* d_status);

/* ---- factors kept on the device: factors.frame [A]:572-592 + prediction [A]:637-654
 * The reference factors every posterior row once (`factors`, [A]:550-559), ships R.Inv, beta and the factor vectors
 * through a data.frame (n^2 + 2n + 5 doubles per row, [A]:587-591) and lets `prediction` walk rows x sites.  Here the
 * per-row state is the Cholesky factor, kept in HBM and keyed by the row index (= sample id):
 *   ccgp_factors_create   factors the S rows of `pars` (same arguments as ccgp_predict) on the design in place;
 *   ccgp_factors_predict  the T x S tables at new sites from the stored factors (only the site phase runs; values are
 *                         bit-identical to ccgp_predict on the same rows), any number of times;
 *   ccgp_factors_info     rows, whether the factors are stored (Gaussian families within the shared-memory kernel;
 *                         otherwise only the parameters are kept and every prediction re-factors), bytes of HBM held;
 *   ccgp_factors_destroy  releases them.  A ccgp_set_design with a different design invalidates the object
 *                         (CCGP_ERR_ARG from predict).  On a ccgp_create_multi context the rows are split over the GPUs. */
typedef struct ccgp_factors ccgp_factors;
int ccgp_factors_create(ccgp_ctx* ctx, int family, const double* pars, int64_t S, int64_t ldp,
                        int vec_family, const double* pars_vec, int64_t ldpv, ccgp_factors** out);
int ccgp_factors_predict(ccgp_ctx* ctx, const ccgp_factors* f, const double* Xnew, int64_t T, double sigma2,
                         double* out_mean, double* out_var, int32_t* out_status);
int ccgp_factors_predict_dev(ccgp_ctx* ctx, const ccgp_factors* f, const double* d_Xnew, int64_t T, double sigma2,
                             double* d_mean, double* d_var, int32_t* d_status);
int ccgp_factors_info(ccgp_ctx* ctx, const ccgp_factors* f, int64_t* rows, int* stored, int64_t* device_bytes);
int ccgp_factors_destroy(ccgp_ctx* ctx, ccgp_factors* f);

/* ---- maximum-entropy design criteria -------------------------------------
 * ccgp_me_schur_batch: Augmented.Mixed.Entropy [M]:869-877 for C candidate
 * second-batch designs x P parameter rows (GAUSS_ISO rows (p,theta1,theta2),
 * P x 3 column-major, ldq >= P).
 *   D_old: n_old x d column-major (n_old may be 0: then the value is Entropy(D)
 *          [M]:856-861, -det of the candidate's own mixed correlation matrix)
 *   D_new: C blocks of n_new*d doubles, block c = candidate c's n_new x d matrix
 *          column-major (= optim's `par` vector c(D.new), [M]:927-930)
 *   out_negdet: C x P column-major, -det(R.new - R.cross R.old^-1 R.cross')
 *   out_logdet: same shape, log of that determinant (may be NULL) */
int ccgp_me_schur_batch(ccgp_ctx* ctx, const double* D_old, int n_old, int d,
                        const double* D_new, int n_new, int64_t C,
                        const double* params, int64_t P, int64_t ldq,
                        double* out_negdet, double* out_logdet, int32_t* out_status);
int ccgp_me_schur_batch_dev(ccgp_ctx* ctx, const double* d_D_old, int n_old, int d,
                            const double* d_D_new, int n_new, int64_t C,
                            const double* d_params, int64_t P, int64_t ldq,
                            double* d_negdet, double* d_logdet, int32_t* d_status);
/* Paired form of the criterion: C = P * group designs, design c evaluated against parameter row
 * c / group only (out_negdet[C]).  This is the shape `Batch.Entropy.optim` ([M]:920-948) takes when
 * it is run for many posterior draws and starts in lock-step: each (draw, start) brings its own
 * finite-difference stencil of designs. */
int ccgp_me_schur_paired(ccgp_ctx* ctx, const double* D_old, int n_old, int d,
                         const double* D_new, int n_new, int64_t group,
                         const double* params, int64_t P, int64_t ldq,
                         double* out_negdet, int32_t* out_status);
/* Stencil form for the finite-difference gradients of optim's L-BFGS-B inside Batch.Entropy.optim ([M]:936):
 * X holds K = P * group base designs as c(D.new) vectors (n_new*d doubles each, problem k -> parameter row
 * k / group); every one is evaluated at its 2m+1 central-difference points (m = n_new*d, step h, clipped
 * to [lo, hi]) generated on the device.  out_vals[k*(2m+1) + s]: s = 0 base, 1+2i: coordinate i + h,
 * 2+2i: coordinate i - h. */
int ccgp_me_schur_stencil(ccgp_ctx* ctx, const double* D_old, int n_old, int d,
                          const double* X, int n_new, int64_t group,
                          const double* params, int64_t P, int64_t ldq,
                          double h, double lo, double hi,
                          double* out_vals, int32_t* out_status);
/* which.min over the candidates of each parameter row ([M]:944-945):
 * best_idx[q] = first c minimising out_negdet[c,q], best_val[q] its value. */
int ccgp_me_argmin(ccgp_ctx* ctx, const double* D_old, int n_old, int d,
                   const double* D_new, int n_new, int64_t C,
                   const double* params, int64_t P, int64_t ldq,
                   double* best_val, int64_t* best_idx);

/* k-medoids (PAM: BUILD + steepest SWAP, Euclidean distance) of n points (n x d column-major,
 * host): the clustering step that condenses the 1000 second-batch designs of All_Subdesigns.txt
 * into `k-medoids ME Design.txt` (reference ReadMe.md:54-60; R's cluster::pam).
 *   out_medoids[k]: 0-based row indices of the medoids (BUILD order, exchanged in place by SWAP)
 *   out_cost: sum of distances to the nearest medoid; out_swaps: exchanges performed (either may be NULL) */
int ccgp_kmedoids_pam(ccgp_ctx* ctx, const double* P, int64_t n, int d, int k, int max_swaps,
                      int32_t* out_medoids, double* out_cost, int32_t* out_swaps);

/* CGP comparator (the composite GP every reference script re-states, [A]:60-319 = "2D Combined GP
 * Anisotropic Public.R").  Xs: n x p column-major design STANDARDISED to [0,1] per column ([A]:70); W: B parameter
 * rows (lambda, theta_1..theta_p, kappa, bw), column-major with leading dimension ldw.
 * ccgp_cgp_objective_batch replaces `apply(starts, 1, var.MLE.DK)` ([A]:104-135, 148): out_val[b] = log(det(Q)) +
 *   n log(tau2) after the four re-weighting passes, 1e6 where the reference's value is not finite.
 * ccgp_cgp_jackknife replaces the leave-one-out loop ([A]:166-199) for ONE parameter row w[p+3]:
 *   out_yp[jf] = Yp_jackknife[jf].  (theta = Stand_theta / scales^2 on the raw design, [A]:164-165, is the same
 *   correlation as Stand_theta on the standardised one.)  out_status (may be NULL): 1 = a pivot of Q was not positive. */
int ccgp_cgp_objective_batch(ccgp_ctx* ctx, const double* Xs, const double* y, int n, int p, const double* W, int64_t B,
                             int64_t ldw, double* out_val, int32_t* out_status);
int ccgp_cgp_jackknife(ccgp_ctx* ctx, const double* Xs, const double* y, int n, int p, const double* w, double* out_yp,
                       int32_t* out_status);

/* log det R[S,S] for C index subsets (0-based, C x m column-major, ldi >= C) of a
 * pool of N points (N x d column-major): the ME subset log-dets of the scaling
 * case.  One natural-scale parameter row of `family`. */
int ccgp_subset_logdet_batch(ccgp_ctx* ctx, const double* pool, int64_t N, int d,
                             const int32_t* idx, int m, int64_t C, int64_t ldi,
                             int family, const double* params,
                             double* out_logdet, int32_t* out_status);
int ccgp_subset_logdet_batch_dev(ccgp_ctx* ctx, const double* d_pool, int64_t N, int d,
                                 const int32_t* d_idx, int m, int64_t C, int64_t ldi,
                                 int family, const double* params_host,
                                 double* d_logdet, int32_t* d_status);

/* ---- plain correlation blocks: out[i + na*j] = mixed correlation of A_i and B_j
 * (A is na x d, B is nb x d, column-major; B == NULL means B = A with an exact unit
 * diagonal).  Mixed.corr.matrix [A]:399-406 (B = NULL), Mixed.corr.vec [A]:416-422
 * (A = x.new as 1 x d, B = D.train), cross.corr.matrix [M]:835-848 (GAUSS_ISO row
 * (1, theta, theta)).  One natural-scale parameter row. */
int ccgp_mixed_corr(ccgp_ctx* ctx, int family, const double* params, const double* A, int na,
                    const double* B, int nb, int d, double* out);

/* debug aid: clock counters per kernel phase (block 0; see tools/phase_timing.py).
 * enable != 0 allocates/zeroes the 32-slot buffer, out32 (may be NULL) receives the
 * current counters first; enable == 0 frees it. */
int ccgp_debug_phase_timing(ccgp_ctx* ctx, int enable, long long* out32);

#ifdef __cplusplus
}
#endif
#endif /* CCGP_H */
