// DMMA (mma.sync.m8n8k4.f64) throughput / latency on sm_100a, next to plain DFMA.
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o ubench_dmma ubench_dmma.cu && ./ubench_dmma
#include <cstdio>
#include <cuda_runtime.h>

__device__ __forceinline__ void dmma(double& c0, double& c1, double a, double b) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                 : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
}

template <int ILP>
__global__ void k_dmma(double* out, long long* cyc, int iters) {
    double c0[ILP], c1[ILP];
    for (int j = 0; j < ILP; ++j) { c0[j] = threadIdx.x * 1e-3 + j; c1[j] = j; }
    double a = 1e-6 * threadIdx.x, b = 1e-6;
    __syncthreads();
    long long t0 = clock64();
    for (int i = 0; i < iters; ++i) {
#pragma unroll
        for (int j = 0; j < ILP; ++j) dmma(c0[j], c1[j], a, b);
    }
    long long t1 = clock64();
    if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
    double s = 0; for (int j = 0; j < ILP; ++j) s += c0[j] + c1[j];
    if (s == 123.0) out[0] = s;
}
template <int ILP>
__global__ void k_dfma(double* out, long long* cyc, int iters) {
    double c0[ILP];
    for (int j = 0; j < ILP; ++j) { c0[j] = threadIdx.x * 1e-3 + j; }
    double a = 0.999999, b = 1e-6;
    __syncthreads();
    long long t0 = clock64();
    for (int i = 0; i < iters; ++i) {
#pragma unroll
        for (int j = 0; j < ILP; ++j) c0[j] = fma(c0[j], a, b);
    }
    long long t1 = clock64();
    if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
    double s = 0; for (int j = 0; j < ILP; ++j) s += c0[j];
    if (s == 123.0) out[0] = s;
}
// half the warps DMMA, half DFMA: do they share a pipe?
template <int ILP>
__global__ void k_mixed(double* out, long long* cyc, int iters) {
    double c0[ILP], c1[ILP];
    for (int j = 0; j < ILP; ++j) { c0[j] = threadIdx.x * 1e-3 + j; c1[j] = j; }
    double a = 0.999999, b = 1e-6;
    const bool tens = ((threadIdx.x >> 5) & 4) != 0;   // warps 4..7 of each 8
    __syncthreads();
    long long t0 = clock64();
    if (tens) {
        for (int i = 0; i < iters; ++i) {
#pragma unroll
            for (int j = 0; j < ILP; ++j) dmma(c0[j], c1[j], a, b);
        }
    } else {
        for (int i = 0; i < iters; ++i) {
#pragma unroll
            for (int j = 0; j < ILP; ++j) c0[j] = fma(c0[j], a, b);
        }
    }
    long long t1 = clock64();
    if ((threadIdx.x & 31) == 0) cyc[blockIdx.x * 2 + (tens ? 1 : 0)] = t1 - t0;
    double s = 0; for (int j = 0; j < ILP; ++j) s += c0[j] + c1[j];
    if (s == 123.0) out[0] = s;
}

int main() {
    double* out; long long* cyc;
    cudaMalloc(&out, 64); cudaMalloc(&cyc, 8 * 4096);
    long long h[4096];
    int dev = 0; cudaDeviceProp p; cudaGetDeviceProperties(&p, dev);
    const int nsm = p.multiProcessorCount;
    const int iters = 20000;
    auto run = [&](const char* name, auto kern, int threads, int ilp, double fma_per_instr) {
        kern<<<nsm, threads>>>(out, cyc, 100);
        cudaDeviceSynchronize();
        cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
        cudaEventRecord(e0);
        kern<<<nsm, threads>>>(out, cyc, iters);
        cudaEventRecord(e1);
        cudaDeviceSynchronize();
        float ms; cudaEventElapsedTime(&ms, e0, e1);
        cudaMemcpy(h, cyc, 8 * nsm, cudaMemcpyDeviceToHost);
        double warps = threads / 32.0;
        double instr = (double)iters * ilp * warps;              // warp instructions per SM
        double clk_per_instr_sm = (double)h[0] / instr;
        double tflops = 2.0 * fma_per_instr * instr * nsm / (ms * 1e-3) / 1e12;
        printf("%-10s thr=%4d ilp=%2d  clk/warp-instr/SM=%6.2f  FMA/clk/SM=%6.1f  %6.2f TFLOP/s  (%.3f ms)\n", name, threads,
               ilp, clk_per_instr_sm, fma_per_instr / clk_per_instr_sm, tflops, ms);
    };
    run("dmma dep", k_dmma<1>, 32, 1, 256);
    run("dfma dep", k_dfma<1>, 32, 1, 32);
    run("dmma", k_dmma<2>, 32, 2, 256);
    run("dmma", k_dmma<4>, 32, 4, 256);
    run("dmma", k_dmma<8>, 32, 8, 256);
    run("dmma", k_dmma<1>, 128, 1, 256);
    run("dmma", k_dmma<2>, 128, 2, 256);
    run("dmma", k_dmma<4>, 128, 4, 256);
    run("dmma", k_dmma<8>, 128, 8, 256);
    run("dmma", k_dmma<4>, 256, 4, 256);
    run("dmma", k_dmma<8>, 256, 8, 256);
    run("dmma", k_dmma<8>, 512, 8, 256);
    run("dmma", k_dmma<4>, 1024, 4, 256);
    run("dfma", k_dfma<8>, 128, 8, 32);
    run("dfma", k_dfma<8>, 256, 8, 32);
    run("dfma", k_dfma<8>, 512, 8, 32);
    run("dfma", k_dfma<8>, 1024, 8, 32);
    // mixed
    {
        k_mixed<8><<<nsm, 256>>>(out, cyc, iters);
        cudaDeviceSynchronize();
        cudaMemcpy(h, cyc, 16 * nsm, cudaMemcpyDeviceToHost);
        double instr = (double)iters * 8 * 4;
        printf("mixed 256 thr (4 warps DFMA + 4 warps DMMA): dfma clk/warp-instr/SM=%.2f  dmma clk/warp-instr/SM=%.2f\n",
               h[0] / instr, h[1] / instr);
    }
    printf("err=%s\n", cudaGetErrorString(cudaGetLastError()));
    return 0;
}
