// bigchol_ws.h -- device workspace of the large-n blocked path (bigchol.cuh), owned by the context
#pragma once
#include <cuda_runtime.h>
#include "factor_engine.cuh"

namespace ccgp {

// everything that parameterises the launches of one large-n batch (bigchol_nll_batch): two identical keys = the same schedule
struct BigGraphKey {
    const void *A, *prm, *linv, *X, *y, *cand, *nll, *beta, *status, *idx;
    cudaStream_t stream;
    int64_t B, ldc, ldi, ldx;
    double sigma2, tau;
    int n, d, family, scale, mean_mode, nrp, flags, pad_;
};

struct BigCholWorkspace {
    double* A = nullptr;       // chunk * nrp * ncp
    double* logdet = nullptr;  // chunk
    int* bad = nullptr;        // chunk
    Prm* prm = nullptr;        // chunk
    double* linv = nullptr;    // chunk * 64 * 64: inverse of the current diagonal block, transposed (k-major B operand)
    size_t bytesA = 0;
    int cap = 0;
    cudaStream_t side = nullptr;          // lookahead: high-priority stream of the serial per-column chain (bigchol.cuh)
    cudaEvent_t ev_main = nullptr, ev_side[2] = {nullptr, nullptr};   // ev_side[k & 1]: the early part of column k is in
    cudaGraphExec_t gexec = nullptr;      // the schedule of the last repeated call as a CUDA graph (bigchol_nll_batch)
    BigGraphKey gkey, last_key;
    bool have_last = false;
    int glaunches = 0;
    void release() {
        if (A) cudaFree(A);
        if (logdet) cudaFree(logdet);
        if (bad) cudaFree(bad);
        if (prm) cudaFree(prm);
        if (linv) cudaFree(linv);
        if (gexec) cudaGraphExecDestroy(gexec);
        gexec = nullptr; have_last = false;
        if (side) cudaStreamDestroy(side);
        if (ev_main) cudaEventDestroy(ev_main);
        for (int i = 0; i < 2; ++i) { if (ev_side[i]) cudaEventDestroy(ev_side[i]); ev_side[i] = nullptr; }
        side = nullptr; ev_main = nullptr;
        A = nullptr; logdet = nullptr; bad = nullptr; prm = nullptr; linv = nullptr; bytesA = 0; cap = 0;
    }
};

}  // namespace ccgp
