// Single-warp latency microbenchmarks for the FP64 path on sm_100a (B200).
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o ubench ubench.cu && ./ubench
#include <cstdio>
#include <cuda_runtime.h>

__global__ void k_dfma_dep(double* out, long long* cyc, int iters, int nwarps_active) {
    double a = threadIdx.x * 1e-3 + 1.0;
    const double m = 0.999999, c = 1e-9;
    __syncthreads();
    long long t0 = clock64();
    for (int i = 0; i < iters; ++i) { a = fma(a, m, c); a = fma(a, m, c); a = fma(a, m, c); a = fma(a, m, c); }
    long long t1 = clock64();
    if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
    if (a == 123.0) out[0] = a;
}
template <int ILP>
__global__ void k_dfma_ilp(double* out, long long* cyc, int iters) {
    double a[ILP];
    for (int j = 0; j < ILP; ++j) a[j] = threadIdx.x * 1e-3 + j;
    const double m = 0.999999, c = 1e-9;
    long long t0 = clock64();
    for (int i = 0; i < iters; ++i) {
#pragma unroll
        for (int j = 0; j < ILP; ++j) a[j] = fma(a[j], m, c);
    }
    long long t1 = clock64();
    if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
    double s = 0; for (int j = 0; j < ILP; ++j) s += a[j];
    if (s == 123.0) out[0] = s;
}
__global__ void k_rsqrt_dep(double* out, long long* cyc, int iters) {
    double a = threadIdx.x * 1e-3 + 1.5;
    long long t0 = clock64();
    for (int i = 0; i < iters; ++i) { a = rsqrt(a) + 1.0; }
    long long t1 = clock64();
    if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
    if (a == 123.0) out[0] = a;
}
__global__ void k_rcp_dep(double* out, long long* cyc, int iters) {
    double a = threadIdx.x * 1e-3 + 1.5;
    long long t0 = clock64();
    for (int i = 0; i < iters; ++i) { a = 1.0 / a + 1.0; }
    long long t1 = clock64();
    if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
    if (a == 123.0) out[0] = a;
}
__global__ void k_shfl_dep(double* out, long long* cyc, int iters) {
    double a = threadIdx.x * 1e-3 + 1.5;
    long long t0 = clock64();
    for (int i = 0; i < iters; ++i) { a = __shfl_sync(0xffffffffu, a, (threadIdx.x + 1) & 31); }
    long long t1 = clock64();
    if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
    if (a == 123.0) out[0] = a;
}
__global__ void k_lds_dep(double* out, long long* cyc, int iters) {
    __shared__ int idx[1024];
    for (int i = threadIdx.x; i < 1024; i += blockDim.x) idx[i] = (i * 33 + 7) & 1023;
    __syncthreads();
    int p = threadIdx.x;
    long long t0 = clock64();
    for (int i = 0; i < iters; ++i) { p = idx[p]; }
    long long t1 = clock64();
    if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
    if (p == -5) out[0] = p;
}
__global__ void k_bar(double* out, long long* cyc, int iters) {
    long long t0 = clock64();
    for (int i = 0; i < iters; ++i) { __syncthreads(); }
    long long t1 = clock64();
    if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}
// FP32 MUFU.RSQ seed + 2 Newton steps in FP64
__global__ void k_rsqrt_fast(double* out, long long* cyc, int iters) {
    double a = threadIdx.x * 1e-3 + 1.5;
    long long t0 = clock64();
    for (int i = 0; i < iters; ++i) {
        double y = (double)rsqrtf((float)a);
        double h = 0.5 * a;
        y = y * fma(-h * y, y, 1.5);
        y = y * fma(-h * y, y, 1.5);
        a = y + 1.0;
    }
    long long t1 = clock64();
    if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
    if (a == 123.0) out[0] = a;
}

int main() {
    double* out; long long* cyc; long long h[4096];
    cudaMalloc(&out, 64); cudaMalloc(&cyc, 4096 * 8);
    const int it = 4096;
    auto rep = [&](const char* name, double per, int nb) {
        cudaMemcpy(h, cyc, nb * 8, cudaMemcpyDeviceToHost);
        long long mx = 0; for (int i = 0; i < nb; ++i) mx = h[i] > mx ? h[i] : mx;
        printf("%-46s %8.2f clk\n", name, mx / per);
    };
    k_dfma_dep<<<1, 32>>>(out, cyc, it, 1); rep("dependent DFMA, 1 warp", it * 4.0, 1);
    k_dfma_dep<<<1, 128>>>(out, cyc, it, 4); rep("dependent DFMA, 4 warps (1/SMSP)", it * 4.0, 1);
    k_dfma_dep<<<1, 512>>>(out, cyc, it, 16); rep("dependent DFMA, 16 warps (4/SMSP) per-warp", it * 4.0, 1);
    k_dfma_dep<<<148 * 4, 128>>>(out, cyc, it, 16); rep("dependent DFMA, 4 CTAs x 4 warps / SM", it * 4.0, 148 * 4);
    k_dfma_ilp<2><<<1, 32>>>(out, cyc, it); rep("DFMA ILP2 1 warp (per DFMA)", it * 2.0, 1);
    k_dfma_ilp<4><<<1, 32>>>(out, cyc, it); rep("DFMA ILP4 1 warp (per DFMA)", it * 4.0, 1);
    k_dfma_ilp<8><<<1, 32>>>(out, cyc, it); rep("DFMA ILP8 1 warp (per DFMA)", it * 8.0, 1);
    k_dfma_ilp<16><<<1, 32>>>(out, cyc, it); rep("DFMA ILP16 1 warp (per DFMA)", it * 16.0, 1);
    k_dfma_ilp<8><<<1, 128>>>(out, cyc, it); rep("DFMA ILP8 4 warps (per DFMA per warp)", it * 8.0, 1);
    k_dfma_ilp<8><<<1, 512>>>(out, cyc, it); rep("DFMA ILP8 16 warps (per DFMA per warp)", it * 8.0, 1);
    k_rsqrt_dep<<<1, 32>>>(out, cyc, it); rep("dependent rsqrt(double)+add", it, 1);
    k_rsqrt_fast<<<1, 32>>>(out, cyc, it); rep("dependent rsqrtf seed + 2 Newton + add", it, 1);
    k_rcp_dep<<<1, 32>>>(out, cyc, it); rep("dependent 1/x (double) + add", it, 1);
    k_shfl_dep<<<1, 32>>>(out, cyc, it); rep("dependent shfl of a double (2 SHFL)", it, 1);
    k_lds_dep<<<1, 32>>>(out, cyc, it); rep("dependent LDS.32", it, 1);
    k_bar<<<1, 128>>>(out, cyc, it); rep("__syncthreads, 4 warps", it, 1);
    cudaDeviceSynchronize();
    printf("%s\n", cudaGetErrorString(cudaGetLastError()));
    return 0;
}
