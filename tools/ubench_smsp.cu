// Which warps of a CTA share an SM sub-partition?  Warps 0 and j run a saturating DFMA loop
// (ILP 8: one warp alone fills its sub-partition's FP64 pipe); the loop takes twice as long when
// both sit on the same sub-partition.
#include <cstdio>
#include <cuda_runtime.h>

__global__ void k(double* out, long long* cyc, int iters, int j) {
    const int w = threadIdx.x >> 5;
    double a[8];
    for (int q = 0; q < 8; ++q) a[q] = threadIdx.x * 1e-3 + q;
    __syncthreads();
    if (w == 0 || w == j) {
        long long t0 = clock64();
        for (int i = 0; i < iters; ++i) {
#pragma unroll
            for (int q = 0; q < 8; ++q) a[q] = fma(a[q], 0.999999, 1e-9);
        }
        long long t1 = clock64();
        if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
    }
    double s = 0; for (int q = 0; q < 8; ++q) s += a[q];
    if (s == 123.0) out[0] = s;
}
int main() {
    double* out; long long* cyc; cudaMalloc(&out, 8); cudaMalloc(&cyc, 8 * 148);
    const int iters = 20000;
    for (int threads : {256, 512}) {
        for (int j = 0; j < threads / 32; ++j) {
            k<<<148, threads>>>(out, cyc, iters, j);
            cudaDeviceSynchronize();
            long long h; cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
            printf("threads %d: warps 0 and %2d: %.2f clk per DFMA (warp 0)\n", threads, j, (double)h / iters / 8);
        }
    }
    printf("%s\n", cudaGetErrorString(cudaGetLastError()));
    return 0;
}
