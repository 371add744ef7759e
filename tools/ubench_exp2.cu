// Per-warp cost of the two device exponentials (polynomial vs table-driven) with ONE warp per
// SM sub-partition (the occupancy of factor_warp_kernel), in clocks per warp-level exp.
#include <cstdio>
#include <cuda_runtime.h>
#include "../convex-combination-of-gaussian-processes_b200/csrc/ccgp_math.h"

template <int ILP, int MODE, bool RND>
__global__ void __launch_bounds__(512, 1) k_exp(double* out, long long* cyc, int iters, double seed) {
    __shared__ double T[128];
    for (int e = threadIdx.x; e < 128; e += blockDim.x) T[e] = CCGP_EXP2_TAB[e];
    __syncthreads();
    double s[ILP], acc[ILP];
    for (int j = 0; j < ILP; ++j) { unsigned h = (threadIdx.x * 2654435761u + j * 40503u) >> 8; s[j] = seed + (RND ? (h & 0xffff) * (40.0 / 65536.0) : threadIdx.x * 0.173 + j * 0.37); acc[j] = 0; }
    long long t0 = clock64();
    for (int i = 0; i < iters; ++i) {
#pragma unroll
        for (int j = 0; j < ILP; ++j) {
            double v;
            if (MODE == 0) v = dexp_neg_dev<false>(s[j]);
            else if (MODE == 1) v = dexp_neg_tab_dev<false>(s[j], T);
            else v = dexp_neg_tab_dev<false>(s[j], T + 0 * (threadIdx.x & 0)) ;
            acc[j] += v; s[j] += 1e-3;
        }
    }
    long long t1 = clock64();
    if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
    double t = 0; for (int j = 0; j < ILP; ++j) t += acc[j];
    if (t == 123.0) out[0] = t;
}

template <int ILP, int MODE, bool RND>
void run(const char* name, int threads) {
    double* out; long long* cyc; cudaMalloc(&out, 8); cudaMalloc(&cyc, 8 * 148);
    const int iters = 4000;
    k_exp<ILP, MODE, RND><<<148, threads>>>(out, cyc, 10, 0.5);
    k_exp<ILP, MODE, RND><<<148, threads>>>(out, cyc, iters, 0.5);
    cudaDeviceSynchronize();
    long long h[148]; cudaMemcpy(h, cyc, 8 * 148, cudaMemcpyDeviceToHost);
    printf("%-34s threads/SM %4d: %6.1f clk per warp-exp per warp, %6.1f per SMSP (incl. 2 FP64 adds)\n", name, threads, (double)h[0] / iters / ILP, (double)h[0] / iters / ILP / (threads / 128));
    cudaFree(out); cudaFree(cyc);
}

int main() {
    for (int threads : {128, 256, 512}) {
        run<8, 0, false>("poly  ILP8 regular args", threads);
        run<8, 0, true>("poly  ILP8 random args", threads);
        run<8, 1, false>("table ILP8 regular args", threads);
        run<8, 1, true>("table ILP8 random args", threads);
        run<4, 1, true>("table ILP4 random args", threads);
    }
    printf("%s\n", cudaGetErrorString(cudaGetLastError()));
    return 0;
}
