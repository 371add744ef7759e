// ccgp_ctx.h -- the context object and host-side helpers shared by the translation units of libccgp.so
#pragma once
#include <cuda_runtime.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <stdint.h>
#include <map>
#include <vector>
#include <algorithm>
#include "../../include/ccgp.h"
#include "factor_engine.cuh"
#include "bigchol_ws.h"

using namespace ccgp;

struct ccgp_ctx {
    int device = 0;
    int num_sm = 0;
    int max_smem_optin = 0;
    cudaStream_t stream = nullptr;
    cudaStream_t own_stream = nullptr;
    cudaStream_t copy_stream = nullptr;   // host<->device copies of the chunk-pipelined host-pointer entry points
    cudaEvent_t ev_h2d[2] = {nullptr, nullptr}, ev_kern[2] = {nullptr, nullptr};
    int n = 0, d = 0;
    double* d_X = nullptr;
    double* d_y = nullptr;
    size_t cap_X = 0, cap_y = 0;          // bytes allocated behind d_X / d_y (set_design reuses them)
    std::vector<double> h_X, h_y;         // host copy of the design in place: an identical set_design is a no-op
    double span2[MAXD] = {0};   // squared coordinate ranges of the shared design
    int twonu = 10;             // Matern smoothness of the 1-D families, 2*nu (reference default nu = 5)
    double mnorm = 1.0 / 384.0; // 1 / (Gamma(nu) 2^(nu-1))
    std::map<std::vector<int>, uint32_t*> tiletabs;   // (n, naug, TR, TC) -> tile table
    void* ws = nullptr;       // device workspace for the host-pointer entry points
    size_t ws_bytes = 0;
    void* ws2 = nullptr;      // small device workspace (params, reductions)
    size_t ws2_bytes = 0;
    char err[512] = {0};
    int64_t launches = 0;
    int last_team = 0, last_smem = 0, last_ctas = 0, last_variant = -1;
    BigCholWorkspace big;
    long long* dbg = nullptr;  // phase-timing buffer (debug)
    int* sm_slots = nullptr;   // per-SM CTA arrival counters of the DMMA kernel
    double* last_bv_dev = nullptr;     // device results of the last argmin_columns call (value, index per column):
    long long* last_bi_dev = nullptr;  // what the multi-GPU front end feeds to the NCCL (min, index) all-reduce
    double* fac_scratch = nullptr;     // ccgp_predict with few rows and many sites: factors of the rows between its two launches
    size_t fac_scratch_bytes = 0;
    uint64_t design_gen = 0;           // bumped by every ccgp_set_design that changes the design (ccgp_factors are tied to it)
    struct MultiCtx* multi = nullptr;  // multi.cu: set on the front context of ccgp_create_multi
};

// Cholesky factors of S posterior rows kept in HBM (ccgp_factors_*, include/ccgp.h): what `factors.frame` ([A]:572-592)
// ships through a data.frame as R.Inv + factor vectors per row.
struct ccgp_factors {
    ccgp_ctx* owner = nullptr;
    uint64_t design_gen = 0;
    int family = 0, vec_family = -1, k = 0, kv = 0;
    int64_t S = 0;
    int stored = 0;                    // 1: factors in HBM (tensor-path kernel); 0: parameters only, every prediction re-factors
    double* d_pars = nullptr;          // S x k
    double* d_pv = nullptr;            // S x kv (parameters of the correlation vector when they differ, quirk Q2) or NULL
    double* d_fac = nullptr;           // S rows of fac_ld doubles: L | inverse diagonal tiles | bad flag, pad
    int32_t* d_status = nullptr;       // S
    int64_t fac_ld = 0;
    std::vector<ccgp_factors*> child;  // front context of ccgp_create_multi: one slice of rows per GPU
};


#define CK(call)                                                                                   \
    do {                                                                                           \
        cudaError_t e_ = (call);                                                                   \
        if (e_ != cudaSuccess) {                                                                   \
            snprintf(ctx->err, sizeof(ctx->err), "%s:%d %s: %s", __FILE__, __LINE__, #call,        \
                     cudaGetErrorString(e_));                                                      \
            return CCGP_ERR_CUDA;                                                                  \
        }                                                                                          \
    } while (0)

#define RC(call)                                                                                   \
    do {                                                                                           \
        int rc_ = (call);                                                                          \
        if (rc_) return rc_;                                                                       \
    } while (0)

#define ARG(cond)                                                                                  \
    do {                                                                                           \
        if (!(cond)) {                                                                             \
            snprintf(ctx->err, sizeof(ctx->err), "bad argument: %s", #cond);                       \
            return CCGP_ERR_ARG;                                                                   \
        }                                                                                          \
    } while (0)

inline int env_int(const char* name, int dflt) {
    const char* s = getenv(name);
    return (s && *s) ? atoi(s) : dflt;
}


typedef void (*factor_fn)(const FactorArgs);
int get_tiletab(ccgp_ctx* ctx, const Layout& l, int TR, int TC, const uint32_t** out);
// factor_launch.cu / factor_mma_launch.cu
int launch_factor(ccgp_ctx* ctx, FactorArgs& A);
int launch_factor_mma(ccgp_ctx* ctx, FactorArgs& A, int* launched);
int launch_factor_pack(ccgp_ctx* ctx, FactorArgs& A, int* launched, int min_warps = 1);   // factor_pack_launch.cu
int launch_factor_pc(ccgp_ctx* ctx, FactorArgs& A, int* launched);     // factor_pc_launch.cu
// multi.cu -- the front context fans every batch out over its per-GPU contexts
void ccgp_set_create_error(const char* msg);
void multi_destroy(ccgp_ctx* front);
int64_t multi_launches(const ccgp_ctx* front);
int multi_set_design(ccgp_ctx* front, const double* X, int n, int d, const double* y);
int multi_set_matern_nu(ccgp_ctx* front, double nu);
int multi_nll_batch(ccgp_ctx* front, int family, int scale, const double* cand, int64_t B, int64_t ldc, double sigma2,
                    int mean_mode, double tau, double* out_nll, double* out_beta, int32_t* out_status);
int multi_nll_argmin(ccgp_ctx* front, int family, int scale, const double* cand, int64_t B, int64_t ldc, double sigma2,
                     int mean_mode, double tau, double* best_nll, int64_t* best_idx);
int multi_predict(ccgp_ctx* front, int family, const double* pars, int64_t S, int64_t ldp, int vec_family,
                  const double* pars_vec, int64_t ldpv, const double* Xnew, int64_t T, double sigma2, double* out_mean,
                  double* out_var, int32_t* out_status);
int multi_factors_create(ccgp_ctx* front, ccgp_factors* f, const double* pars, int64_t ldp, const double* pars_vec, int64_t ldpv);
int multi_factors_predict(ccgp_ctx* front, const ccgp_factors* f, const double* Xnew, int64_t T, double sigma2, double* out_mean,
                          double* out_var, int32_t* out_status);
int multi_me_schur_batch(ccgp_ctx* front, const double* D_old, int n_old, int d, const double* D_new, int n_new, int64_t C,
                         const double* params, int64_t P, int64_t ldq, double* out_negdet, double* out_logdet, int32_t* out_status);
int multi_me_argmin(ccgp_ctx* front, const double* D_old, int n_old, int d, const double* D_new, int n_new, int64_t C,
                    const double* params, int64_t P, int64_t ldq, double* best_val, int64_t* best_idx);
