// Throughput of the device exp under the factor kernel's occupancy (16 warps/SM), by ILP.
#include <cstdio>
#include <cuda_runtime.h>
#include "../convex-combination-of-gaussian-processes_b200/csrc/ccgp_math.h"

template <int ILP, bool CL>
__global__ void __launch_bounds__(128, 4) k_exp(double* out, int iters, double seed) {
    double s[ILP], acc[ILP];
    for (int j = 0; j < ILP; ++j) { s[j] = seed + threadIdx.x * 1e-3 + j * 0.37; acc[j] = 0; }
    for (int i = 0; i < iters; ++i) {
#pragma unroll
        for (int j = 0; j < ILP; ++j) { acc[j] += dexp_neg_dev<CL>(s[j]); s[j] += 1e-3; }
    }
    double t = 0; for (int j = 0; j < ILP; ++j) t += acc[j];
    if (t == 123.0) out[0] = t;
}

template <int ILP, bool CL>
void run(const char* name, int grid) {
    double* out; cudaMalloc(&out, 8);
    const int iters = 2000;
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    k_exp<ILP, CL><<<grid, 128>>>(out, iters, 0.5);
    cudaEventRecord(e0);
    k_exp<ILP, CL><<<grid, 128>>>(out, iters, 0.5);
    cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    double nexp = (double)grid * 128 * iters * ILP;
    // 16 FP64-pipe instructions per exp (incl. the accumulate and the argument step)
    printf("%-28s grid %5d: %7.1f Gexp/s  -> %5.1f clk per exp per SMSP-warp-instr, FP64 instr rate %.2f /clk/SMSP\n", name, grid,
           nexp / ms / 1e6, 0.0, nexp / 32 * 17 / (ms * 1e-3) / (148 * 4) / 1.965e9);
    cudaFree(out);
}

int main() {
    run<1, false>("ILP1 noclamp", 148 * 4);
    run<2, false>("ILP2 noclamp", 148 * 4);
    run<4, false>("ILP4 noclamp", 148 * 4);
    run<8, false>("ILP8 noclamp", 148 * 4);
    run<4, true>("ILP4 clamp", 148 * 4);
    run<4, false>("ILP4 noclamp 1 CTA/SM", 148);
    run<8, false>("ILP8 noclamp 1 CTA/SM", 148);
    printf("%s\n", cudaGetErrorString(cudaGetLastError()));
    return 0;
}
