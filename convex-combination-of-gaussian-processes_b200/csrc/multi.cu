// multi.cu -- all GPUs of one box behind ONE context, reachable from a single-threaded caller (R's .Call).
//
// candidates / posterior rows / (design, parameter row) pairs are independent given the
// broadcast training design, so the front context owns one ordinary context per GPU, splits every batch
// into contiguous slices (slice g -> GPU g, results land in the caller's own output slice) and drives the
// GPUs from one host thread each.  The only exchange is the (min value, index) pair of
// `which.min` ([V]:598 choose.hyperpars, [M]:944-945 Batch.Entropy.optim): NCCL has no MINLOC, so
//   1. all-reduce(MIN) over the per-GPU best values,
//   2. all-reduce(MIN) over  (own value == global min ? own global index : INT64_MAX)
// -- lowest index wins ties, so the result is the same for every GPU count.  The communicators come from
// ncclCommInitAll (single process) over NVLink / NVSwitch; libnccl is dlopen()ed when a multi-GPU context is
// created, so the single-GPU library has no NCCL dependency.
#include <dlfcn.h>
#include <nccl.h>
#include <thread>
#include "ccgp_ctx.h"

struct NcclApi {
    void* handle = nullptr;
    ncclResult_t (*CommInitAll)(ncclComm_t*, int, const int*) = nullptr;
    ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
    ncclResult_t (*AllReduce)(const void*, void*, size_t, ncclDataType_t, ncclRedOp_t, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*GroupStart)() = nullptr;
    ncclResult_t (*GroupEnd)() = nullptr;
    const char* (*GetErrorString)(ncclResult_t) = nullptr;
};

struct MultiCtx {
    int G = 0;
    std::vector<ccgp_ctx*> child;
    NcclApi nccl;
    std::vector<ncclComm_t> comms;
    std::vector<double*> d_val;        // per GPU: [2 * cap] doubles (own values | reduced values)
    std::vector<long long*> d_idx;     // per GPU: [2 * cap] (own candidates | reduced indices)
    int64_t cap = 0;
    int64_t collectives = 0;
};

static int load_nccl(NcclApi& n, char* err, size_t errlen) {
    const char* names[] = {"libnccl.so.2", "libnccl.so", nullptr};
    const char* env = getenv("CCGP_NCCL_LIB");
    if (env && *env) n.handle = dlopen(env, RTLD_NOW | RTLD_LOCAL);
    for (int i = 0; !n.handle && names[i]; ++i) n.handle = dlopen(names[i], RTLD_NOW | RTLD_LOCAL);
    if (!n.handle) {
        snprintf(err, errlen, "ccgp_create_multi: cannot load libnccl.so.2 (%s); set CCGP_NCCL_LIB", dlerror());
        return CCGP_ERR_UNSUPPORTED;
    }
#define SYM(field, name) do { *(void**)(&n.field) = dlsym(n.handle, name); \
        if (!n.field) { snprintf(err, errlen, "ccgp_create_multi: libnccl lacks %s", name); return CCGP_ERR_UNSUPPORTED; } } while (0)
    SYM(CommInitAll, "ncclCommInitAll");
    SYM(CommDestroy, "ncclCommDestroy");
    SYM(AllReduce, "ncclAllReduce");
    SYM(GroupStart, "ncclGroupStart");
    SYM(GroupEnd, "ncclGroupEnd");
    SYM(GetErrorString, "ncclGetErrorString");
#undef SYM
    return 0;
}

#define NC(call) do { ncclResult_t r_ = (call); if (r_ != ncclSuccess) { \
        snprintf(ctx->err, sizeof(ctx->err), "%s:%d %s: %s", __FILE__, __LINE__, #call, M->nccl.GetErrorString(r_)); \
        return CCGP_ERR_CUDA; } } while (0)

extern "C" int ccgp_create_multi(ccgp_ctx** out, int n_gpus) {
    if (!out) return CCGP_ERR_ARG;
    *out = nullptr;
    int count = 0;
    if (cudaGetDeviceCount(&count) != cudaSuccess || count < 1) {
        ccgp_ctx* probe = nullptr;
        return ccgp_create(&probe, 0);            // sets the "no usable CUDA device" message
    }
    if (n_gpus <= 0) n_gpus = count;
    ccgp_ctx* front = nullptr;
    int rc = ccgp_create(&front, 0);
    if (rc) return rc;
    if (n_gpus > count) {
        snprintf(front->err, sizeof(front->err), "ccgp_create_multi: %d GPUs requested, %d visible", n_gpus, count);
        ccgp_set_create_error(front->err);
        ccgp_destroy(front);
        return CCGP_ERR_ARG;
    }
    MultiCtx* M = new MultiCtx();
    M->G = n_gpus;
    front->multi = M;
    auto fail = [&](int code) { ccgp_set_create_error(front->err); ccgp_destroy(front); return code; };
    for (int g = 0; g < n_gpus; ++g) {
        ccgp_ctx* c = nullptr;
        rc = ccgp_create(&c, g);
        if (rc) { snprintf(front->err, sizeof(front->err), "%s", ccgp_last_error(nullptr)); return fail(rc); }
        M->child.push_back(c);
    }
    if (n_gpus > 1) {
        rc = load_nccl(M->nccl, front->err, sizeof(front->err));
        if (rc) return fail(rc);
        std::vector<int> devs(n_gpus);
        for (int g = 0; g < n_gpus; ++g) devs[g] = g;
        M->comms.resize(n_gpus);
        ncclResult_t r = M->nccl.CommInitAll(M->comms.data(), n_gpus, devs.data());
        if (r != ncclSuccess) {
            M->comms.clear();
            snprintf(front->err, sizeof(front->err), "ncclCommInitAll(%d): %s", n_gpus, M->nccl.GetErrorString(r));
            return fail(CCGP_ERR_CUDA);
        }
    }
    *out = front;
    return CCGP_OK;
}

extern "C" int ccgp_num_gpus(const ccgp_ctx* ctx) { return !ctx ? 0 : (ctx->multi ? ctx->multi->G : 1); }
extern "C" int64_t ccgp_collective_count(const ccgp_ctx* ctx) { return (ctx && ctx->multi) ? ctx->multi->collectives : 0; }

void multi_destroy(ccgp_ctx* front) {
    MultiCtx* M = front->multi;
    if (!M) return;
    for (size_t g = 0; g < M->comms.size(); ++g) M->nccl.CommDestroy(M->comms[g]);
    for (size_t g = 0; g < M->d_val.size(); ++g) {
        cudaSetDevice(M->child[g]->device);
        if (M->d_val[g]) cudaFree(M->d_val[g]);
        if (M->d_idx[g]) cudaFree(M->d_idx[g]);
    }
    for (ccgp_ctx* c : M->child) ccgp_destroy(c);
    // libnccl stays loaded: unloading a library that owns CUDA state at process exit is not safe
    delete M;
    front->multi = nullptr;
}

int64_t multi_launches(const ccgp_ctx* front) {
    int64_t s = 0;
    for (ccgp_ctx* c : front->multi->child) s += c->launches;
    return s;
}

// run fn(g) on one host thread per GPU; first failure wins, its message is copied to the front context
template <class F>
static int fan_out(ccgp_ctx* front, F fn) {
    MultiCtx* M = front->multi;
    std::vector<int> rc(M->G, 0);
    if (M->G == 1) {
        rc[0] = fn(0);
    } else {
        std::vector<std::thread> th;
        for (int g = 0; g < M->G; ++g) th.emplace_back([&, g] { rc[g] = fn(g); });
        for (auto& t : th) t.join();
    }
    for (int g = 0; g < M->G; ++g)
        if (rc[g]) { snprintf(front->err, sizeof(front->err), "GPU %d: %.480s", g, M->child[g]->err); return rc[g]; }
    return 0;
}
static inline void slice(int64_t total, int g, int G, int64_t* b0, int64_t* nb) {
    *b0 = total * g / G;
    *nb = total * (g + 1) / G - *b0;
}

int multi_set_design(ccgp_ctx* front, const double* X, int n, int d, const double* y) {
    int rc = fan_out(front, [&](int g) { return ccgp_set_design(front->multi->child[g], X, n, d, y); });
    if (rc == 0) { front->n = n; front->d = d; }
    return rc;
}
int multi_set_matern_nu(ccgp_ctx* front, double nu) {
    return fan_out(front, [&](int g) { return ccgp_set_matern_nu(front->multi->child[g], nu); });
}

int multi_nll_batch(ccgp_ctx* front, int family, int scale, const double* cand, int64_t B, int64_t ldc, double sigma2,
                    int mean_mode, double tau, double* out_nll, double* out_beta, int32_t* out_status) {
    const int G = front->multi->G;
    return fan_out(front, [&](int g) {
        int64_t b0, nb;
        slice(B, g, G, &b0, &nb);
        return ccgp_nll_batch(front->multi->child[g], family, scale, cand + b0, nb, ldc, sigma2, mean_mode, tau,
                              out_nll + b0, out_beta ? out_beta + b0 : nullptr, out_status ? out_status + b0 : nullptr);
    });
}

int multi_predict(ccgp_ctx* front, int family, const double* pars, int64_t S, int64_t ldp, int vec_family,
                  const double* pars_vec, int64_t ldpv, const double* Xnew, int64_t T, double sigma2, double* out_mean,
                  double* out_var, int32_t* out_status) {
    const int G = front->multi->G;
    return fan_out(front, [&](int g) {
        int64_t s0, ns;
        slice(S, g, G, &s0, &ns);
        return ccgp_predict(front->multi->child[g], family, pars + s0, ns, ldp, vec_family, pars_vec ? pars_vec + s0 : nullptr, ldpv,
                            Xnew, T, sigma2, out_mean + (size_t)T * s0, out_var + (size_t)T * s0, out_status ? out_status + s0 : nullptr);
    });
}

// factors kept on the device (ccgp_factors_*): GPU g factors and keeps rows slice(S, g, G); predictions land in the
// caller's column slices like multi_predict's
int multi_factors_create(ccgp_ctx* front, ccgp_factors* f, const double* pars, int64_t ldp, const double* pars_vec, int64_t ldpv) {
    const int G = front->multi->G;
    f->child.assign(G, nullptr);
    return fan_out(front, [&](int g) {
        int64_t s0, ns;
        slice(f->S, g, G, &s0, &ns);
        return ccgp_factors_create(front->multi->child[g], f->family, pars + s0, ns, ldp, f->vec_family,
                                   pars_vec ? pars_vec + s0 : nullptr, ldpv, &f->child[g]);
    });
}
int multi_factors_predict(ccgp_ctx* front, const ccgp_factors* f, const double* Xnew, int64_t T, double sigma2, double* out_mean,
                          double* out_var, int32_t* out_status) {
    const int G = front->multi->G;
    return fan_out(front, [&](int g) {
        int64_t s0, ns;
        slice(f->S, g, G, &s0, &ns);
        return ccgp_factors_predict(front->multi->child[g], f->child[g], Xnew, T, sigma2, out_mean + (size_t)T * s0,
                                    out_var + (size_t)T * s0, out_status ? out_status + s0 : nullptr);
    });
}

// ---- the (min, index) all-reduce over NCCL ----------------------------------------------------------------------
// per GPU, L slots: own best value (NaN / idx < 0 -> +inf) and own best GLOBAL index
__global__ void multi_prepare_kernel(const double* bv, const long long* bi, long long offset, double* val, long long* idx, int64_t L) {
    const int64_t q = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (q >= L) return;
    const bool ok = bi[q] >= 0 && bv[q] == bv[q];
    val[q] = ok ? bv[q] : __longlong_as_double(0x7ff0000000000000LL);
    idx[q] = ok ? bi[q] + offset : 0x7fffffffffffffffLL;
}
__global__ void multi_select_kernel(const double* val, const double* gmin, long long* idx, int64_t L) {
    const int64_t q = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (q >= L) return;
    if (!(val[q] == gmin[q])) idx[q] = 0x7fffffffffffffffLL;
}

// children hold their local results (value, index relative to their slice) in last_bv_dev / last_bi_dev
static int multi_reduce_argmin(ccgp_ctx* ctx, const std::vector<int64_t>& offset, int64_t L, double* best_val, int64_t* best_idx) {
    MultiCtx* M = ctx->multi;
    const int G = M->G;
    if ((int)M->d_val.size() < G) { M->d_val.resize(G, nullptr); M->d_idx.resize(G, nullptr); }
    if (L > M->cap) {
        for (int g = 0; g < G; ++g) {
            CK(cudaSetDevice(M->child[g]->device));
            if (M->d_val[g]) CK(cudaFree(M->d_val[g]));
            if (M->d_idx[g]) CK(cudaFree(M->d_idx[g]));
            CK(cudaMalloc(&M->d_val[g], (size_t)L * 16));
            CK(cudaMalloc(&M->d_idx[g], (size_t)L * 16));
        }
        M->cap = L;
    }
    const unsigned blocks = (unsigned)((L + 255) / 256);
    for (int g = 0; g < G; ++g) {
        ccgp_ctx* c = M->child[g];
        CK(cudaSetDevice(c->device));
        multi_prepare_kernel<<<blocks, 256, 0, c->stream>>>(c->last_bv_dev, c->last_bi_dev, offset[g], M->d_val[g], M->d_idx[g], L);
        CK(cudaGetLastError());
        c->launches++;
    }
    if (G > 1) {
        NC(M->nccl.GroupStart());
        for (int g = 0; g < G; ++g)
            NC(M->nccl.AllReduce(M->d_val[g], M->d_val[g] + L, (size_t)L, ncclDouble, ncclMin, M->comms[g], M->child[g]->stream));
        NC(M->nccl.GroupEnd());
        for (int g = 0; g < G; ++g) {
            ccgp_ctx* c = M->child[g];
            CK(cudaSetDevice(c->device));
            multi_select_kernel<<<blocks, 256, 0, c->stream>>>(M->d_val[g], M->d_val[g] + L, M->d_idx[g], L);
            CK(cudaGetLastError());
            c->launches++;
        }
        NC(M->nccl.GroupStart());
        for (int g = 0; g < G; ++g)
            NC(M->nccl.AllReduce(M->d_idx[g], M->d_idx[g] + L, (size_t)L, ncclInt64, ncclMin, M->comms[g], M->child[g]->stream));
        NC(M->nccl.GroupEnd());
        M->collectives += 2;
    }
    ccgp_ctx* c0 = M->child[0];
    CK(cudaSetDevice(c0->device));
    std::vector<long long> hi((size_t)L);
    const size_t red = (G > 1) ? (size_t)L : 0;              // reduced halves (G == 1: the prepared values themselves)
    CK(cudaMemcpyAsync(best_val, M->d_val[0] + red, (size_t)L * 8, cudaMemcpyDeviceToHost, c0->stream));
    CK(cudaMemcpyAsync(hi.data(), M->d_idx[0] + red, (size_t)L * 8, cudaMemcpyDeviceToHost, c0->stream));
    for (int g = 0; g < G; ++g) { CK(cudaSetDevice(M->child[g]->device)); CK(cudaStreamSynchronize(M->child[g]->stream)); }
    for (int64_t q = 0; q < L; ++q) {
        const bool none = hi[q] == 0x7fffffffffffffffLL;
        best_idx[q] = none ? -1 : (int64_t)hi[q];
        if (none) best_val[q] = __builtin_nan("");
    }
    return 0;
}

int multi_nll_argmin(ccgp_ctx* front, int family, int scale, const double* cand, int64_t B, int64_t ldc, double sigma2,
                     int mean_mode, double tau, double* best_nll, int64_t* best_idx) {
    const int G = front->multi->G;
    std::vector<int64_t> off(G);
    std::vector<double> lv(G);
    std::vector<int64_t> li(G);
    int rc = fan_out(front, [&](int g) {
        int64_t b0, nb;
        slice(B, g, G, &b0, &nb);
        off[g] = b0;
        ccgp_ctx* c = front->multi->child[g];
        if (nb == 0) { c->last_bv_dev = nullptr; return 0; }
        return ccgp_nll_argmin(c, family, scale, cand + b0, nb, ldc, sigma2, mean_mode, tau, &lv[g], &li[g]);
    });
    if (rc) return rc;
    for (int g = 0; g < G; ++g)
        if (!front->multi->child[g]->last_bv_dev) {
            snprintf(front->err, sizeof(front->err), "ccgp_nll_argmin: batch of %lld is smaller than the %d GPUs", (long long)B, G);
            return CCGP_ERR_ARG;
        }
    ccgp_ctx* ctx = front;
    RC(multi_reduce_argmin(ctx, off, 1, best_nll, best_idx));
    return 0;
}

int multi_me_schur_batch(ccgp_ctx* front, const double* D_old, int n_old, int d, const double* D_new, int n_new, int64_t C,
                         const double* params, int64_t P, int64_t ldq, double* out_negdet, double* out_logdet, int32_t* out_status) {
    const int G = front->multi->G;
    if (P >= G) {                                       // parameter rows are whole output columns: no scatter
        return fan_out(front, [&](int g) {
            int64_t q0, nq;
            slice(P, g, G, &q0, &nq);
            return ccgp_me_schur_batch(front->multi->child[g], D_old, n_old, d, D_new, n_new, C, params + q0, nq, ldq,
                                       out_negdet ? out_negdet + (size_t)C * q0 : nullptr, out_logdet ? out_logdet + (size_t)C * q0 : nullptr,
                                       out_status ? out_status + (size_t)C * q0 : nullptr);
        });
    }
    // few parameter rows, many designs: slice the designs, scatter the column pieces
    std::vector<std::vector<double>> tn(G), tl(G);
    std::vector<std::vector<int32_t>> ts(G);
    int rc = fan_out(front, [&](int g) {
        int64_t c0, nc;
        slice(C, g, G, &c0, &nc);
        if (nc == 0) return 0;
        if (out_negdet) tn[g].resize((size_t)nc * P);
        if (out_logdet) tl[g].resize((size_t)nc * P);
        if (out_status) ts[g].resize((size_t)nc * P);
        int r = ccgp_me_schur_batch(front->multi->child[g], D_old, n_old, d, D_new + (size_t)c0 * n_new * d, n_new, nc, params, P, ldq,
                                    out_negdet ? tn[g].data() : nullptr, out_logdet ? tl[g].data() : nullptr, out_status ? ts[g].data() : nullptr);
        if (r) return r;
        for (int64_t q = 0; q < P; ++q) {
            if (out_negdet) memcpy(out_negdet + (size_t)C * q + c0, tn[g].data() + (size_t)nc * q, (size_t)nc * 8);
            if (out_logdet) memcpy(out_logdet + (size_t)C * q + c0, tl[g].data() + (size_t)nc * q, (size_t)nc * 8);
            if (out_status) memcpy(out_status + (size_t)C * q + c0, ts[g].data() + (size_t)nc * q, (size_t)nc * 4);
        }
        return 0;
    });
    return rc;
}

int multi_me_argmin(ccgp_ctx* front, const double* D_old, int n_old, int d, const double* D_new, int n_new, int64_t C,
                    const double* params, int64_t P, int64_t ldq, double* best_val, int64_t* best_idx) {
    const int G = front->multi->G;
    if (P >= G && !env_int("CCGP_MULTI_ME_SPLIT_DESIGNS", 0)) {   // every parameter row's argmin lives on one GPU: no collective
        return fan_out(front, [&](int g) {
            int64_t q0, nq;
            slice(P, g, G, &q0, &nq);
            return ccgp_me_argmin(front->multi->child[g], D_old, n_old, d, D_new, n_new, C, params + q0, nq, ldq, best_val + q0, best_idx + q0);
        });
    }
    if (C < G) { snprintf(front->err, sizeof(front->err), "ccgp_me_argmin: %lld designs for %d GPUs", (long long)C, G); return CCGP_ERR_ARG; }
    std::vector<int64_t> off(G);
    std::vector<std::vector<double>> lv(G);
    std::vector<std::vector<int64_t>> li(G);
    int rc = fan_out(front, [&](int g) {
        int64_t c0, nc;
        slice(C, g, G, &c0, &nc);
        off[g] = c0;
        lv[g].resize((size_t)P); li[g].resize((size_t)P);
        return ccgp_me_argmin(front->multi->child[g], D_old, n_old, d, D_new + (size_t)c0 * n_new * d, n_new, nc, params, P, ldq,
                              lv[g].data(), li[g].data());
    });
    if (rc) return rc;
    ccgp_ctx* ctx = front;
    RC(multi_reduce_argmin(ctx, off, P, best_val, best_idx));
    return 0;
}
