// cgp.cu -- the CGP comparator's batched objective: `var.MLE.DK` of the composite GP every
// reference script carries ([A]:104-135 = "2D Combined GP Anisotropic Public.R"), evaluated for the 505 start
// candidates of its sweep ([A]:137-151) in one launch, and the leave-one-out loop behind `rmscv` / `Yp_jackknife`
// ([A]:166-201) as n work items of the same kernel.
//
// Per work item (one CTA, all matrices resident in shared memory, packed lower triangles):
//   G = exp(-s), L = exp(-(s + kappa r^2)), Gbw = exp(-bw s),  s_ij = sum_k theta_k d_k^2, r^2_ij = sum_k d_k^2
//   Sig = I;  4 x { Q = G + lambda Sig^1/2 L Sig^1/2;  Q = C C' (Cholesky; the reference's solve(Q) is an LU of the same
//                   SPD matrix);  beta = 1'Q^-1 y / 1'Q^-1 1;  temp = Q^-1 (y - beta 1);  e = y - beta 1 - G temp;
//                   Sig = diag(Gbw e^2 / Gbw 1), normalised by its mean }
//   once more Q, C, beta, temp;  tau2 = (y - beta 1)' temp / n;
//   objective:  log(det(Q)) + n log(tau2), 1e6 when not finite -- det(Q) itself is formed (it underflows to 0 for the
//               larger designs, which the reference turns into 1e6: reproduced);
//   jackknife:  the same on the n-1 points without jf, then Yp[jf] = beta + q' temp,
//               q = g + lambda sqrt(v) Sig^1/2 l,  v = (gbw' e^2 / gbw' 1) / mean(Sig)   ([A]:190-197).
// The two right-hand sides ride along the factorisation as two extra rows (z_y = C^-1 y, z_1 = C^-1 1), then one backward
// substitution each.  Throughput is not the point here (505 + n small factorisations); residency and one launch are.
#include "ccgp_ctx.h"

namespace {

constexpr int CGP_THREADS = 128;

struct CgpArgs {
    const double* Xs;     // n x p column-major, standardised to [0, 1] ([A]:70)
    const double* y;      // n
    const double* W;      // rows (lambda, theta_1..theta_p, kappa, bw), column-major, ld ldw
    int64_t ldw;
    int64_t B;            // work items
    int n, p;
    int jack;             // 0: item b = parameter row b, all n points; 1: parameter row 0, item b leaves point b out
    double* out;          // objective values (jack = 0) / leave-one-out predictions (jack = 1)
    int32_t* status;      // 1: a pivot was not positive (optional)
};

__device__ __forceinline__ int tri(int i, int j) { return i * (i + 1) / 2 + j; }          // i >= j

__device__ __forceinline__ double block_sum(double v, double* red) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    __syncthreads();
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = v;
    __syncthreads();
    double s = 0.0;
    for (int w = 0; w < CGP_THREADS / 32; ++w) s += red[w];
    return s;
}

__global__ void __launch_bounds__(CGP_THREADS) cgp_kernel(const CgpArgs A) {
    extern __shared__ __align__(16) double sm[];
    const int n = A.n, p = A.p, tid = threadIdx.x;
    const int m = A.jack ? n - 1 : n;
    const int T = m * (m + 1) / 2;
    double* G = sm;
    double* Lm = G + T;
    double* Gb = Lm + T;
    double* Q = Gb + T;
    double* xs = Q + T;              // m x p, point-major
    double* yv = xs + m * p;
    double* S = yv + m;              // Sig diagonal
    double* e2 = S + m;
    double* zy = e2 + m;             // C^-1 y, then Q^-1 y
    double* z1 = zy + m;             // C^-1 1, then Q^-1 1
    double* tmp = z1 + m;            // temp = Q^-1 (y - beta 1)
    double* red = tmp + m;           // 8 doubles
    double* xo = red + 8;            // the left-out point (p coordinates)

    for (int64_t b = blockIdx.x; b < A.B; b += gridDim.x) {
        const int jf = A.jack ? (int)b : -1;
        const double* w = A.W + (A.jack ? 0 : b);
        const double lam = w[0], kappa = w[(int64_t)(p + 1) * A.ldw], bw = w[(int64_t)(p + 2) * A.ldw];
        __syncthreads();
        for (int i = tid; i < m; i += CGP_THREADS) {
            const int src = (jf >= 0 && i >= jf) ? i + 1 : i;
            for (int k = 0; k < p; ++k) xs[i * p + k] = A.Xs[(int64_t)k * n + src];
            yv[i] = A.y[src];
            S[i] = 1.0;
        }
        if (jf >= 0 && tid < p) xo[tid] = A.Xs[(int64_t)tid * n + jf];
        __syncthreads();
        // ---- the three correlation matrices, row per thread ----
        for (int i = tid; i < m; i += CGP_THREADS) {
            for (int j = 0; j <= i; ++j) {
                double s = 0.0, r2 = 0.0;
                for (int k = 0; k < p; ++k) {
                    const double d = xs[i * p + k] - xs[j * p + k];
                    s = fma(w[(int64_t)(1 + k) * A.ldw] * d, d, s);
                    r2 = fma(d, d, r2);
                }
                const int e = tri(i, j);
                G[e] = exp(-s);
                Lm[e] = exp(-(s + kappa * r2));
                Gb[e] = exp(-bw * s);
            }
        }
        __syncthreads();
        int bad = 0;
        double beta = 0.0, sig2 = 1.0, logdet = 0.0;
        for (int pass = 0; pass < 5; ++pass) {
            // ---- Q = G + lambda Sig^1/2 L Sig^1/2 ----
            for (int i = tid; i < m; i += CGP_THREADS) {
                const double si = sqrt(S[i]);
                for (int j = 0; j <= i; ++j) Q[tri(i, j)] = fma(lam * si * sqrt(S[j]), Lm[tri(i, j)], G[tri(i, j)]);
            }
            __syncthreads();
            // ---- Cholesky by columns (left-looking, thread <-> row), y and 1 as rows m and m+1 ----
            logdet = 0.0;
            for (int k = 0; k < m; ++k) {
                double t = 0.0;
                const double* rk = Q + tri(k, 0);
                for (int i = tid; i < m + 2; i += CGP_THREADS) {
                    if (i < k) continue;
                    double acc;
                    const double* ri;
                    if (i < m) { ri = Q + tri(i, 0); acc = ri[k]; }
                    else { ri = (i == m) ? zy : z1; acc = (i == m) ? yv[k] : 1.0; }
                    for (int j = 0; j < k; ++j) acc = fma(-ri[j], rk[j], acc);
                    t = acc;
                    if (i < m) Q[tri(i, k)] = acc; else if (i == m) zy[k] = acc; else z1[k] = acc;
                }
                __syncthreads();
                const double piv = Q[tri(k, k)];
                if (!(piv > 0.0)) bad = 1;
                const double ckk = sqrt(piv);
                logdet += 2.0 * log(ckk);
                __syncthreads();
                for (int i = tid; i < m + 2; i += CGP_THREADS) {
                    if (i < k) continue;
                    if (i == k) Q[tri(k, k)] = ckk;
                    else if (i < m) Q[tri(i, k)] = t / ckk;
                    else if (i == m) zy[k] = t / ckk;
                    else z1[k] = t / ckk;
                }
                __syncthreads();
            }
            // ---- backward substitutions: zy <- Q^-1 y, z1 <- Q^-1 1 ----
            for (int k = m - 1; k >= 0; --k) {
                const double ckk = Q[tri(k, k)];
                const double uy = zy[k] / ckk, u1 = z1[k] / ckk;
                __syncthreads();
                const double* rk = Q + tri(k, 0);
                for (int j = tid; j < k; j += CGP_THREADS) { zy[j] = fma(-rk[j], uy, zy[j]); z1[j] = fma(-rk[j], u1, z1[j]); }
                if (tid == 0) { zy[k] = uy; z1[k] = u1; }
                __syncthreads();
            }
            double a = 0.0, c = 0.0;
            for (int i = tid; i < m; i += CGP_THREADS) { a += zy[i]; c += z1[i]; }
            a = block_sum(a, red);
            c = block_sum(c, red);
            beta = a / c;
            for (int i = tid; i < m; i += CGP_THREADS) tmp[i] = fma(-beta, z1[i], zy[i]);
            __syncthreads();
            if (pass == 4) break;
            // ---- residuals of the global part, new Sig ----
            for (int i = tid; i < m; i += CGP_THREADS) {
                double gip = beta;
                for (int j = 0; j < m; ++j) gip = fma(G[i >= j ? tri(i, j) : tri(j, i)], tmp[j], gip);
                const double e = yv[i] - gip;
                e2[i] = e * e;
            }
            __syncthreads();
            double ssum = 0.0;
            for (int i = tid; i < m; i += CGP_THREADS) {
                double num = 0.0, den = 0.0;
                for (int j = 0; j < m; ++j) {
                    const double g = Gb[i >= j ? tri(i, j) : tri(j, i)];
                    num = fma(g, e2[j], num);
                    den += g;
                }
                S[i] = num / den;
                ssum += S[i];
            }
            sig2 = block_sum(ssum, red) / m;
            for (int i = tid; i < m; i += CGP_THREADS) S[i] /= sig2;
            __syncthreads();
        }
        if (!A.jack) {
            double q = 0.0;
            for (int i = tid; i < m; i += CGP_THREADS) q = fma(yv[i] - beta, tmp[i], q);
            q = block_sum(q, red);
            if (tid == 0) {
                const double tau2 = q / m;
                double val = log(exp(logdet)) + m * log(tau2);      // log(det(Q)): the determinant itself, as in [A]:131
                if (bad || !isfinite(val)) val = 1e6;
                A.out[b] = val;
                if (A.status) A.status[b] = bad;
            }
        } else {
            // ---- prediction at the left-out point ([A]:190-198) ----
            double num = 0.0, den = 0.0;
            for (int j = tid; j < m; j += CGP_THREADS) {
                double s = 0.0;
                for (int k = 0; k < p; ++k) { const double d = xo[k] - xs[j * p + k]; s = fma(w[(int64_t)(1 + k) * A.ldw] * d, d, s); }
                const double g = exp(-bw * s);
                num = fma(g, e2[j], num);
                den += g;
            }
            num = block_sum(num, red);
            den = block_sum(den, red);
            const double v = (num / den) / sig2;
            double acc = 0.0;
            for (int j = tid; j < m; j += CGP_THREADS) {
                double s = 0.0, r2 = 0.0;
                for (int k = 0; k < p; ++k) {
                    const double d = xo[k] - xs[j * p + k];
                    s = fma(w[(int64_t)(1 + k) * A.ldw] * d, d, s);
                    r2 = fma(d, d, r2);
                }
                const double qj = exp(-s) + lam * sqrt(v) * sqrt(S[j]) * exp(-(s + kappa * r2));
                acc = fma(qj, tmp[j], acc);
            }
            acc = block_sum(acc, red);
            if (tid == 0) {
                A.out[b] = bad ? __longlong_as_double(0x7ff8000000000000LL) : beta + acc;
                if (A.status) A.status[b] = bad;
            }
        }
    }
}

size_t cgp_smem_bytes(int m, int p) {
    const size_t T = (size_t)m * (m + 1) / 2;
    return (4 * T + (size_t)m * p + 7 * (size_t)m + 8 + MAXD + 2) * 8;
}

int cgp_run(ccgp_ctx* ctx, const double* Xs, const double* y, int n, int p, const double* W, int64_t rows, int64_t ldw, int jack,
            double* out, int32_t* status) {
    ARG(Xs && y && W && out);
    ARG(n >= 3 && n + 2 <= CGP_THREADS && p >= 1 && p <= MAXD && rows >= 1 && ldw >= rows);
    CK(cudaSetDevice(ctx->device));
    const int64_t B = jack ? n : rows;
    const int m = jack ? n - 1 : n;
    const size_t smem = cgp_smem_bytes(m, p);
    if (smem > (size_t)ctx->max_smem_optin) {
        snprintf(ctx->err, sizeof(ctx->err), "ccgp_cgp: n=%d needs %zu B of shared memory (> %d)", n, smem, ctx->max_smem_optin);
        return CCGP_ERR_UNSUPPORTED;
    }
    const int k = p + 3;
    const size_t bytes = ((size_t)n * p + n + (size_t)rows * k + B) * 8 + (size_t)B * 4 + 64;
    void* buf = nullptr;
    CK(cudaMalloc(&buf, bytes));
    double* d_X = (double*)buf;
    double* d_y = d_X + (size_t)n * p;
    double* d_W = d_y + n;
    double* d_out = d_W + (size_t)rows * k;
    int32_t* d_st = (int32_t*)(d_out + B);
    int rc = CCGP_OK;
    auto body = [&]() -> int {
        CK(cudaMemcpyAsync(d_X, Xs, (size_t)n * p * 8, cudaMemcpyHostToDevice, ctx->stream));
        CK(cudaMemcpyAsync(d_y, y, (size_t)n * 8, cudaMemcpyHostToDevice, ctx->stream));
        CK(cudaMemcpy2DAsync(d_W, (size_t)rows * 8, W, (size_t)ldw * 8, (size_t)rows * 8, k, cudaMemcpyHostToDevice, ctx->stream));
        CgpArgs A;
        A.Xs = d_X; A.y = d_y; A.W = d_W; A.ldw = rows; A.B = B; A.n = n; A.p = p; A.jack = jack; A.out = d_out; A.status = d_st;
        CK(cudaFuncSetAttribute(cgp_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        const int grid = (int)std::min<int64_t>(B, (int64_t)ctx->num_sm * std::max<int64_t>(1, ctx->max_smem_optin / (int64_t)smem));
        cgp_kernel<<<grid, CGP_THREADS, smem, ctx->stream>>>(A);
        CK(cudaGetLastError());
        ctx->launches++;
        CK(cudaMemcpyAsync(out, d_out, (size_t)B * 8, cudaMemcpyDeviceToHost, ctx->stream));
        if (status) CK(cudaMemcpyAsync(status, d_st, (size_t)B * 4, cudaMemcpyDeviceToHost, ctx->stream));
        CK(cudaStreamSynchronize(ctx->stream));
        return CCGP_OK;
    };
    rc = body();
    cudaStreamSynchronize(ctx->stream);
    cudaFree(buf);
    return rc;
}

}  // namespace

// var.MLE.DK ([A]:104-135) for B parameter rows (lambda, theta_1..theta_p, kappa, bw) on the standardised design
extern "C" int ccgp_cgp_objective_batch(ccgp_ctx* ctx, const double* Xs, const double* y, int n, int p, const double* W, int64_t B,
                                        int64_t ldw, double* out_val, int32_t* out_status) {
    if (!ctx) return CCGP_ERR_ARG;
    if (B == 0) return CCGP_OK;
    return cgp_run(ctx, Xs, y, n, p, W, B, ldw, 0, out_val, out_status);
}

// Yp_jackknife ([A]:166-199) for ONE parameter row: out_yp[jf] = the composite GP's prediction at point jf from the other n-1
extern "C" int ccgp_cgp_jackknife(ccgp_ctx* ctx, const double* Xs, const double* y, int n, int p, const double* w, double* out_yp,
                                  int32_t* out_status) {
    if (!ctx) return CCGP_ERR_ARG;
    return cgp_run(ctx, Xs, y, n, p, w, 1, 1, 1, out_yp, out_status);
}
