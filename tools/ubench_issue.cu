// Does a DMMA / DFMA stream on one warp block the issue of OTHER instruction classes from a second warp of the
// same SM sub-partition?  Warps w and w+4 of a CTA share sub-partition w mod 4 (tools/ubench_smsp.cu).
// Warps 0..3 run stream A, warps 4..7 stream B; both loops are timed separately (clock64), one CTA per SM.
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o ubench_issue ubench_issue.cu && ./ubench_issue
#include <cstdio>
#include <cuda_runtime.h>

__device__ __forceinline__ void dmma(double& c0, double& c1, double a, double b) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                 : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
}
enum { S_NONE = 0, S_DMMA = 1, S_DFMA = 2, S_IMAD = 3, S_LDS = 4, S_FFMA = 5 };

template <int KIND>
__device__ __forceinline__ long long stream(int iters, double* out, const double* sm) {
    double c0[8], c1[8];
    int x[8];
    float f[8];
    for (int j = 0; j < 8; ++j) { c0[j] = threadIdx.x * 1e-3 + j; c1[j] = j; x[j] = threadIdx.x + j; f[j] = threadIdx.x * 0.5f + j; }
    const double a = 0.999999, b = 1e-6;
    long long t0 = clock64();
    for (int i = 0; i < iters; ++i) {
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            if (KIND == S_DMMA) dmma(c0[j], c1[j], a, b);
            if (KIND == S_DFMA) c0[j] = fma(c0[j], a, b);
            if (KIND == S_IMAD) x[j] = x[j] * 3 + i;
            if (KIND == S_FFMA) f[j] = fmaf(f[j], 0.999f, 1e-3f);
            if (KIND == S_LDS) c0[j] += sm[(x[j] + i) & 255];
        }
    }
    long long t1 = clock64();
    double s = 0; for (int j = 0; j < 8; ++j) s += c0[j] + c1[j] + x[j] + f[j];
    if (s == 123.456) out[0] = s;
    return t1 - t0;
}
template <int KA, int KB>
__global__ void k_pair(double* out, long long* cyc, int iters) {
    __shared__ double sm[256];
    for (int e = threadIdx.x; e < 256; e += blockDim.x) sm[e] = e;
    __syncthreads();
    const bool second = (threadIdx.x >> 5) >= 4;
    long long dt = 0;
    if (!second) { if (KA != S_NONE) dt = stream<KA>(iters, out, sm); }
    else { if (KB != S_NONE) dt = stream<KB>(iters, out, sm); }
    if ((threadIdx.x & 31) == 0 && (threadIdx.x >> 5) % 4 == 0) cyc[blockIdx.x * 2 + (second ? 1 : 0)] = dt;
}
int main() {
    double* out; long long* cyc;
    cudaMalloc(&out, 64); cudaMalloc(&cyc, 16 * 4096);
    long long h[8192];
    cudaDeviceProp p; cudaGetDeviceProperties(&p, 0);
    const int nsm = p.multiProcessorCount, iters = 4000;
    auto run = [&](const char* name, auto kern) {
        cudaMemset(cyc, 0, 16 * 4096);
        kern<<<nsm, 256>>>(out, cyc, 100);
        kern<<<nsm, 256>>>(out, cyc, iters);
        cudaDeviceSynchronize();
        cudaMemcpy(h, cyc, 16 * nsm, cudaMemcpyDeviceToHost);
        printf("%-34s A: %7.2f clk/instr   B: %7.2f clk/instr   (per warp, one warp of each kind per sub-partition)\n", name,
               h[0] / (8.0 * iters), h[1] / (8.0 * iters));
    };
    run("DMMA alone", k_pair<S_DMMA, S_NONE>);
    run("DFMA alone", k_pair<S_DFMA, S_NONE>);
    run("IMAD alone", k_pair<S_IMAD, S_NONE>);
    run("FFMA alone", k_pair<S_FFMA, S_NONE>);
    run("LDS+DADD alone", k_pair<S_LDS, S_NONE>);
    run("DMMA | IMAD", k_pair<S_DMMA, S_IMAD>);
    run("DMMA | FFMA", k_pair<S_DMMA, S_FFMA>);
    run("DMMA | DFMA", k_pair<S_DMMA, S_DFMA>);
    run("DMMA | LDS+DADD", k_pair<S_DMMA, S_LDS>);
    run("DFMA | IMAD", k_pair<S_DFMA, S_IMAD>);
    run("DFMA | FFMA", k_pair<S_DFMA, S_FFMA>);
    run("DFMA | DFMA", k_pair<S_DFMA, S_DFMA>);
    run("DMMA | DMMA", k_pair<S_DMMA, S_DMMA>);
    run("IMAD | IMAD", k_pair<S_IMAD, S_IMAD>);
    printf("err=%s\n", cudaGetErrorString(cudaGetLastError()));
    return 0;
}
