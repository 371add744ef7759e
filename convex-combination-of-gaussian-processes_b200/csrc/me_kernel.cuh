// me_kernel.cuh -- (stub until the specialised ME kernel lands; generic engine is used)
#pragma once
#include "factor_engine.cuh"
namespace ccgp {
inline bool me_fast_supported(int, int, int) { return false; }
inline int me_fast_launch(cudaStream_t, int, const double*, int, int, const double*, int, int64_t, const double*, int64_t,
                          int64_t, double*, double*, int32_t*, char*, size_t) { return -3; }
}
