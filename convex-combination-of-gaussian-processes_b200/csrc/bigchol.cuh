// bigchol.cuh -- (stub until the blocked large-n path lands)
#pragma once
#include "factor_engine.cuh"
namespace ccgp {
struct BigCholWorkspace { void release() {} };
inline int bigchol_nll_batch(BigCholWorkspace&, cudaStream_t, int, const double*, const double*, int n, int, int, int,
                             const double*, int64_t, int64_t, double, int, double, double*, double*, int32_t*, int64_t*,
                             char* err, size_t errlen) {
    snprintf(err, errlen, "n=%d exceeds the shared-memory path and the blocked path is not built", n);
    return -3;
}
}
