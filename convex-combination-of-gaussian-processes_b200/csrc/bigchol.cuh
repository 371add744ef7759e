// bigchol.cuh -- blocked path for designs too large for shared memory (n up to a few
// thousand; SURVEY 8d ME-B: the synthetic n = 2048, 2-D anisotropic scaling case).
//
// Same likelihood as factor_engine.cuh (logpost [A]:444-455 / cond.like [V]:564-575), same
// augmented-matrix trick, but every candidate's matrix lives in HBM (column-major, leading
// dimension nrp) and the factorisation is a LEFT-looking blocked Cholesky, block columns of 64,
// batched over the candidates of a chunk; per block column k:
//   update<false> : C(rows >= 64k, k) -= sum_{j<k} L(rows, j) L(k, j)'  for all 128-row tiles at once, the whole
//                   K = 64k contraction accumulated in registers on the FP64 tensor path
//                   (mma.sync.m8n8k4.f64 -> DMMA; tcgen05 has no FP64 kind), operands staged with cp.async
//   potrf         : the 64x64 diagonal block in shared memory (one CTA per candidate) + its inverse, as 8x8 tiles on the
//                   same tensor path (big_potrf_mma_kernel)
//   update<true>  : rows below = C * inv(L_kk)' on the same tensor path (K = 64) instead of a 64-step substitution
// plus build (mixed correlation tiles, 2 exp per entry, identity padding, rows y', 1') and finish (z-dots, beta,
// Q_R, log det -> NLL, or the log-determinant alone for index subsets of a point pool).
// CCGP_BIG_RIGHT=1 runs the first version of this path (right-looking: potrf, big_trsm_kernel, one K = 64
// big_syrk_kernel launch per panel) for A/B timing.
// Row layout: [0,n) design points, [n,ncp) identity padding (ncp = n rounded up to 64),
// rows ncp and ncp+1 are y' and 1', zero rows up to nrp = ncp + 8 (ncp + 64 under CCGP_BIG_RIGHT: the first version's
// kernels work on 64-row tiles): the update kernel's last row tile then holds 8 or 72 rows and its warps skip the
// quadrants beyond them, instead of multiplying 64 rows of padding per block column.
#pragma once
#include <algorithm>
#include <stdlib.h>
#include "factor_mma.cuh"
#include "bigchol_ws.h"

namespace ccgp {

struct BigArgs {
    double* A;
    int n, d, ncp, nrp;
    int64_t stride;            // nrp * ncp
    const double* X;           // n x d column-major
    const double* y;
    const Prm* prm;
    double* logdet;
    int* bad;
    double* linv;              // per candidate: inv(L_kk)' of the block just factored, linv[q*64 + c] = inv(L_kk)(c, q)
    const int32_t* idx;        // optional gather (subset log-dets): row i of candidate b is pool row idx[(b0 + b) + ldi * i]
    int64_t ldi, ldx, b0;      // ldx: leading dimension of X (n for a shared design, pool rows when gathering)
};

__global__ void __launch_bounds__(128) big_params_kernel(FactorArgs F, int64_t b0, int nb, Prm* prm, double* logdet, int* bad) {
    int b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b < nb) {
        load_params(F, F.shared_row ? 0 : b0 + b, prm + b);
        logdet[b] = 0.0;
        bad[b] = 0;
    }
}

// one CTA (256 threads) per (64x64 tile on or below the diagonal incl. the extra rows, candidate)
__global__ void __launch_bounds__(256) big_build_kernel(BigArgs G) {
    const int b = blockIdx.z, ti = blockIdx.y, tj = blockIdx.x;
    if (ti < tj) return;
    const Prm* pr = G.prm + b;
    __shared__ double xi[MAXD][64], xj[MAXD][64], wts[MAXD], etab[128];
    const int tid = threadIdx.x, n = G.n, d = G.d;
    if (tid < 128) etab[tid] = CCGP_EXP2_TAB[tid];
    for (int e = tid; e < d * 64; e += 256) {
        int k = e / 64, r = e % 64;
        int i = ti * 64 + r, j = tj * 64 + r;
        if (G.idx) {
            xi[k][r] = (i < n) ? G.X[(size_t)k * G.ldx + G.idx[(G.b0 + b) + G.ldi * i]] : 0.0;
            xj[k][r] = (j < n) ? G.X[(size_t)k * G.ldx + G.idx[(G.b0 + b) + G.ldi * j]] : 0.0;
        } else {
            xi[k][r] = (i < n) ? G.X[(size_t)k * G.ldx + i] : 0.0;
            xj[k][r] = (j < n) ? G.X[(size_t)k * G.ldx + j] : 0.0;
        }
    }
    if (tid < d) wts[tid] = pr->wts[tid];
    __syncthreads();
    const double rho = pr->rho, a = pr->a, bb = pr->b;
    double* Ab = G.A + (size_t)b * G.stride;
    for (int e = tid; e < 64 * 64; e += 256) {
        const int r = e % 64, c = e / 64;
        const int i = ti * 64 + r, j = tj * 64 + c;
        if (i >= G.nrp) continue;                         // (the tile of the extra rows is 8 rows tall)
        double v = 0.0;
        if (i < n && j < n) {
            if (i == j) v = 1.0;
            else if (i > j) {
                double s1 = 0.0;
                for (int k = 0; k < d; ++k) { double df = xi[k][r] - xj[k][c]; s1 = fma(wts[k] * df, df, s1); }
                v = fma(bb, dexp_neg_tab_dev<true>(rho * s1, etab), a * dexp_neg_tab_dev<true>(s1, etab));
            }
        } else if (i == j) v = 1.0;                       // identity padding (i, j in [n, ncp))
        else if (j < n && i == G.ncp) v = G.y ? G.y[j] : 0.0;   // row y' (absent in determinant mode)
        else if (j < n && i == G.ncp + 1) v = 1.0;        // row 1'
        Ab[(size_t)j * G.nrp + i] = v;
    }
}

// 64x64 Cholesky of diagonal block k in shared memory; one CTA per candidate
__global__ void __launch_bounds__(256) big_potrf_kernel(BigArgs G, int k) {
    __shared__ double S[64 * 65];
    __shared__ int s_bad;
    const int b = blockIdx.x, tid = threadIdx.x;
    double* Ab = G.A + (size_t)b * G.stride + (size_t)(k * 64) * G.nrp + k * 64;
    for (int e = tid; e < 4096; e += 256) { int r = e % 64, c = e / 64; S[c * 65 + r] = Ab[(size_t)c * G.nrp + r]; }
    if (tid == 0) s_bad = 0;
    __syncthreads();
    // blocked in 8-column panels: warp 0 factors the panel (lane <-> rows lane, lane+32; warp-synchronous, no
    // CTA barrier per column), then all 256 threads apply its rank-8 update to the trailing block
    double ld = 0.0;
    int bad = 0;
    const int lane = tid & 31, warp = tid >> 5;
    for (int j0 = 0; j0 < 64; j0 += 8) {
        if (warp == 0) {
            for (int c = j0; c < j0 + 8; ++c) {
                const double piv = S[c * 65 + c];
                const bool live = (k * 64 + c) < G.n;
                if (live && !(piv > PIVOT_MIN)) bad = 1;
                if (live && lane == 0) ld += log(piv);
                const double ri = fast_rsqrt(piv);
                __syncwarp();
#pragma unroll
                for (int h = 0; h < 2; ++h) {
                    const int r = lane + 32 * h;
                    if (r > c) S[c * 65 + r] *= ri;
                    else if (r == c) S[c * 65 + r] = piv * ri;
                }
                __syncwarp();
                for (int c2 = c + 1; c2 < j0 + 8; ++c2) {
                    const double lc = S[c * 65 + c2];
#pragma unroll
                    for (int h = 0; h < 2; ++h) {
                        const int r = lane + 32 * h;
                        if (r >= c2) S[c2 * 65 + r] = fma(-S[c * 65 + r], lc, S[c2 * 65 + r]);
                    }
                }
                __syncwarp();
            }
        }
        __syncthreads();
        const int nt = 56 - j0;                                 // trailing block: columns/rows j0+8 .. 63
        for (int e = tid; e < nt * nt; e += 256) {
            const int c2 = j0 + 8 + e / nt, r = j0 + 8 + e % nt;
            if (r >= c2) {
                double v = S[c2 * 65 + r];
#pragma unroll
                for (int c = 0; c < 8; ++c) v = fma(-S[(j0 + c) * 65 + r], S[(j0 + c) * 65 + c2], v);
                S[c2 * 65 + r] = v;
            }
        }
        __syncthreads();
    }
    if (warp == 0 && __any_sync(0xffffffffu, bad) && lane == 0) s_bad = 1;
    __syncthreads();
    for (int e = tid; e < 4096; e += 256) { int r = e % 64, c = e / 64; if (r >= c) Ab[(size_t)c * G.nrp + r] = S[c * 65 + r]; }
    if (tid == 0) {
        G.logdet[b] += ld;
        if (s_bad) G.bad[b] = 1;
    }
    // inverse of the factored block, column per thread (forward substitution on e_j); stored transposed so the
    // rows below become a product X = A inv(L_kk)' on the tensor path (big_update_kernel, TRSM mode)
    if (tid < 64) {
        const int j = tid;
        double* out = G.linv + (size_t)b * 4096;
        double x[64];
#pragma unroll
        for (int r = 0; r < 64; ++r) {
            double s = (r == j) ? 1.0 : 0.0;
#pragma unroll
            for (int q = 0; q < r; ++q) s = fma(-S[q * 65 + r], x[q], s);      // L(r, q) at S[q*65 + r]
            x[r] = (r >= j) ? s / S[r * 65 + r] : 0.0;
        }
#pragma unroll
        for (int r = 0; r < 64; ++r) out[(size_t)j * 64 + r] = x[r];            // linv[q = j][c = r] = inv(L)(r, j)
    }
}

// The same step on the FP64 tensor path (the default; the kernel above remains for A/B runs, CCGP_BIG_POTRF_OLD=1).
// The 64x64 block is 8x8 tiles of 8x8 in the fragment layout of factor_mma.cuh; 8 warps run the left-looking tile
// algorithm of the small-n kernels over the 8 tile columns, and the inverse comes out of the SAME column steps:
// X = inv(L)' solves X L' = I, so its tiles are "rows below" whose right-hand side is the identity,
//   X(i, c) = (delta_ic I - sum_{J=i}^{c-1} X(i, J) L(c, J)') inv(L_cc)',  i <= c   (upper block-triangular).
// Step c: warp c accumulates and factors the diagonal tile (mma_diag_impl: 8x8 Cholesky + inverse, every lane holding the
// whole tile), warps w > c own L(w, c), warps w < c own X(w, c): seven products with inv(L_cc)' after one barrier.
// 78 us -> 22 us per block column (ncu, 128 candidates) (the column-at-a-time panel loop and the 64-step substitution per
// thread of the first version were the serial part of the large-n path: 2.5 of 12.4 ms at n = 2048).
__global__ void __launch_bounds__(256) big_potrf_mma_kernel(BigArgs G, int k) {
    __shared__ __align__(16) double Lt[36 * 64];     // tile (r, c), r >= c, at (r (r + 1) / 2 + c) * 64
    __shared__ __align__(16) double Xt[36 * 64];     // tile X(i, c), i <= c, at (c (c + 1) / 2 + i) * 64
    __shared__ __align__(16) double linv_t[64];      // inv(L_cc), row-major
    __shared__ double s_ld[8];
    __shared__ int s_badw[8];
    const int b = blockIdx.x, tid = threadIdx.x, lane = tid & 31, w = tid >> 5;
    double* Ab = G.A + (size_t)b * G.stride + (size_t)(k * 64) * G.nrp + k * 64;
    for (int e = tid; e < 4096; e += 256) {
        const int r = e & 63, c = e >> 6, tr = r >> 3, tc = c >> 3;
        if (tr >= tc) Lt[(tr * (tr + 1) / 2 + tc) * 64 + (r & 7) * 8 + (c & 7)] = Ab[(size_t)c * G.nrp + r];
    }
    __syncthreads();
    FactorArgs FA;                                   // mma_diag_impl reads lay.n (all eight columns live) and tail0 only
    FA.lay.n = 64; FA.tail0 = 0;
    FactorResult res;
    res.mant_all = 1.0; res.mant_tail = 1.0; res.es_all = 0; res.es_tail = 0; res.bad = 0;
    for (int c = 0; c < 8; ++c) {
        double2 acc = make_double2(0.0, 0.0), alt = make_double2(0.0, 0.0);
        const double* ap;                            // this warp's operand tiles of panel J: ap + J * astep
        int J0, astep;
        if (w >= c) {
            acc = ld2(Lt + (w * (w + 1) / 2 + c) * 64 + 2 * lane);
            ap = Lt + (w * (w + 1) / 2) * 64 + 2 * lane; astep = 64; J0 = 0;                  // L(w, J)
        } else {
            ap = Xt + 2 * lane + w * 64; astep = 0; J0 = w;                                     // X(w, J) at (J (J + 1) / 2 + w) * 64
        }
        const double* bp = Lt + (c * (c + 1) / 2) * 64 + 2 * lane;                            // L(c, J)
        for (int J = J0; J < c; ++J) {
            const double2 a = (w >= c) ? ld2(ap + J * astep) : ld2(ap + (J * (J + 1) / 2) * 64);
            const double2 bt = ld2(bp + J * 64);
            mma884(acc.x, acc.y, a.x, negd(bt.x));
            mma884(alt.x, alt.y, a.y, negd(bt.y));
        }
        acc.x += alt.x; acc.y += alt.y;
        if (w == c) {
            double* blk = Lt + (c * (c + 1) / 2 + c) * 64;
            st2(blk + 2 * lane, acc.x, acc.y);
            __syncwarp();
            mma_diag_impl<true>(FA, blk, linv_t, c, lane, res, true);
        }
        __syncthreads();
        if (w != c) {
            const double2 li = ld2(linv_t + 2 * lane);
            double2 x = make_double2(0.0, 0.0);
            mma884(x.x, x.y, acc.x, li.x);
            mma884(x.x, x.y, acc.y, li.y);
            double* dst = (w > c) ? Lt + (w * (w + 1) / 2 + c) * 64 : Xt + (c * (c + 1) / 2 + w) * 64;
            st2(dst + 2 * lane, x.x, x.y);
        } else {                                     // X(c, c) = inv(L_cc)'
            const int row = lane >> 2, col = 2 * (lane & 3);
            st2(Xt + (c * (c + 1) / 2 + c) * 64 + 2 * lane, linv_t[col * 8 + row], linv_t[(col + 1) * 8 + row]);
        }
        __syncthreads();
    }
    if (lane == 0) { s_ld[w] = log(res.mant_all) + res.es_all * LN2; s_badw[w] = res.bad; }      // warp w factored tile column w
    for (int e = tid; e < 4096; e += 256) {
        const int r = e & 63, c = e >> 6, tr = r >> 3, tc = c >> 3;
        if (r >= c) Ab[(size_t)c * G.nrp + r] = Lt[(tr * (tr + 1) / 2 + tc) * 64 + (r & 7) * 8 + (c & 7)];
    }
    double* out = G.linv + (size_t)b * 4096;         // out[q * 64 + cc] = inv(L)(cc, q) = X(q, cc)
    for (int e = tid; e < 4096; e += 256) {
        const int cc = e & 63, q = e >> 6, ti = q >> 3, tc = cc >> 3;
        out[e] = (ti <= tc) ? Xt[(tc * (tc + 1) / 2 + ti) * 64 + (q & 7) * 8 + (cc & 7)] : 0.0;
    }
    __syncthreads();
    if (tid == 0) {
        double ld = 0.0;
        int bad = 0;
        for (int c = 0; c < 8; ++c) { ld += s_ld[c]; bad |= s_badw[c]; }
        G.logdet[b] += ld;
        if (bad) G.bad[b] = 1;
    }
}

// rows of block column k below the diagonal block: X L_kk' = A, one thread per row
__global__ void __launch_bounds__(64) big_trsm_kernel(BigArgs G, int k) {
    extern __shared__ __align__(16) double smt[];
    double* Lk = smt;                   // L_kk(c, c1) at Lk[c1*65 + c]
    double* Xs = smt + 64 * 65;         // the 64 rows of this CTA, Xs[c*65 + r]
    const int b = blockIdx.y, rt = blockIdx.x, tid = threadIdx.x;
    const size_t colk = (size_t)(k * 64) * G.nrp;
    const double* Ld = G.A + (size_t)b * G.stride + colk + k * 64;
    double* Ar = G.A + (size_t)b * G.stride + colk + (size_t)(k + 1 + rt) * 64;
    for (int c = 0; c < 64; ++c) {
        Lk[c * 65 + tid] = Ld[(size_t)c * G.nrp + tid];
        Xs[c * 65 + tid] = Ar[(size_t)c * G.nrp + tid];
    }
    __syncthreads();
    for (int c = 0; c < 64; ++c) {
        double x = Xs[c * 65 + tid];
        for (int c1 = 0; c1 < c; ++c1) x = fma(-Xs[c1 * 65 + tid], Lk[c1 * 65 + c], x);
        Xs[c * 65 + tid] = x / Lk[c * 65 + c];
    }
    for (int c = 0; c < 64; ++c) Ar[(size_t)c * G.nrp + tid] = Xs[c * 65 + tid];
}

__device__ __forceinline__ void dmma884(double& c0, double& c1, double a, double b) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                 : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
}

// trailing update with panel k: C(ti, tj) -= P_ti P_tj', 64x64 tiles, ti >= tj > k.
// 128 threads = 4 warps, warp (wr, wc) owns a 32x32 quadrant = 4x4 DMMA m8n8k4 tiles.
// Panel blocks are staged k-major with leading dimension 72 (k-stride = 64 B mod 128 B) so one
// fragment load of a warp (8 rows x 4 k) hits every bank exactly once per wavefront.
__global__ void __launch_bounds__(128) big_syrk_kernel(BigArgs G, int k, int ntile) {
    extern __shared__ __align__(16) double sm[];
    double* Pi = sm;                // [64 k][72]
    double* Pj = sm + 64 * 72;
    const int b = blockIdx.y;
    // decode the lower-triangular tile index -> (ti, tj), ti >= tj, both counted from k+1
    int t = blockIdx.x, tj = 0, rowlen = ntile;      // ntile = row tiles below/at block k+1
    while (t >= rowlen) { t -= rowlen; ++tj; --rowlen; }
    const int ti = tj + t;
    // column tiles only exist while tj is a real block column
    const int gi = k + 1 + ti, gj = k + 1 + tj;
    const size_t colk = (size_t)(k * 64) * G.nrp;
    const double* Ab = G.A + (size_t)b * G.stride;
    const int tid = threadIdx.x;
    for (int e = tid; e < 4096; e += 128) {
        const int r = e % 64, kk = e / 64;
        Pi[kk * 72 + r] = Ab[colk + (size_t)kk * G.nrp + (size_t)gi * 64 + r];
        Pj[kk * 72 + r] = Ab[colk + (size_t)kk * G.nrp + (size_t)gj * 64 + r];
    }
    __syncthreads();
    const int warp = tid >> 5, lane = tid & 31;
    const int wr = (warp >> 1) * 32, wc = (warp & 1) * 32;
    const int fr = lane >> 2, fk = lane & 3;
    double acc[4][4][2];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) { acc[i][j][0] = 0.0; acc[i][j][1] = 0.0; }
#pragma unroll 4
    for (int k0 = 0; k0 < 64; k0 += 4) {
        double af[4], bf[4];
#pragma unroll
        for (int i = 0; i < 4; ++i) af[i] = Pi[(k0 + fk) * 72 + wr + 8 * i + fr];
#pragma unroll
        for (int j = 0; j < 4; ++j) bf[j] = Pj[(k0 + fk) * 72 + wc + 8 * j + fr];
#pragma unroll
        for (int i = 0; i < 4; ++i)
#pragma unroll
            for (int j = 0; j < 4; ++j) dmma884(acc[i][j][0], acc[i][j][1], af[i], bf[j]);
    }
    double* C = G.A + (size_t)b * G.stride + (size_t)(gj * 64) * G.nrp + (size_t)gi * 64;
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const int r = wr + 8 * i + fr, c = wc + 8 * j + 2 * fk;
            C[(size_t)c * G.nrp + r] -= acc[i][j][0];
            C[(size_t)(c + 1) * G.nrp + r] -= acc[i][j][1];
        }
}

// ---- left-looking update of block column k (all row tiles at once, every previous panel in one launch) ----
//   C(rows, k) -= sum_{j<k} L(rows, j) L(k, j)'        rows: 128-row tiles from the diagonal block down to y', 1'
// The accumulators stay in registers over the whole K = 64 k contraction (the right-looking big_syrk_kernel
// re-read and re-wrote every trailing tile once per panel: HBM-bound at 2 FMA per byte).  256 threads = 8 warps
// (4 x 2), each a 32x32 quadrant = 4x4 DMMA m8n8k4 fragments; operands travel HBM/L2 -> shared memory in
// 32-column chunks with cp.async, double buffered (the chunk after next is in flight while this one is
// multiplied); k-major staging with leading dimensions 136 / 72 (k-stride = 64 B mod 128 B) keeps every fragment
// load conflict-free.
constexpr int BU_KC = 32, BU_LDB = 72;
__device__ __forceinline__ void cp_async16(double* dst_smem, const double* src) {
    const unsigned d = (unsigned)__cvta_generic_to_shared(dst_smem);
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(d), "l"(src) : "memory");
}
// TRSM = true: rows below the diagonal block of column k, X = A inv(L_kk)' (K = 64, B operand = G.linv), in place.
// chlo, chhi: the 32-column chunks of the contraction this launch covers (update mode: panels chlo/2 .. chhi/2 - 1 -- the
// lookahead splits a column's update into "everything but the last panel", launched early on a side stream, and the last
// panel; TRSM mode: 0, 2).
// ROWS = 128: 8 warps as 4 x 2 quadrants of 32 x 32; ROWS = 64: 2 x 4 pieces of 32 x 16 -- half the work per CTA, chosen by
// the host for block columns whose 128-row tiles would leave most SMs idle in the last round (big_update_rows).
template <int ROWS> __host__ __device__ constexpr int bu_lda() { return ROWS + 8; }           // leading dimension of the staged A chunk (k-stride = 64 B mod 128 B)
template <int ROWS> __host__ __device__ constexpr int bu_stage() { return BU_KC * (bu_lda<ROWS>() + BU_LDB); }
template <bool TRSM, int ROWS>
__global__ void __launch_bounds__(256, ROWS == 128 ? 2 : 3) big_update_kernel(BigArgs G, int k, int chlo, int chhi) {
    static_assert(ROWS == 128 || ROWS == 64, "ROWS");
    constexpr int NJF = ROWS == 128 ? 4 : 2;                          // 8-column fragments per warp
    constexpr int BU_LDA = bu_lda<ROWS>(), BU_STAGE = bu_stage<ROWS>();
    extern __shared__ __align__(16) double sm[];
    const int b = blockIdx.y, tid = threadIdx.x;
    const int row0 = (TRSM ? (k + 1) * 64 : k * 64) + blockIdx.x * ROWS;  // first row of this tile
    const int rows_here = min(ROWS, G.nrp - row0);                    // ROWS, or 72 / 8 in the last tile (64 with the old layout)
    const double* Ab = G.A + (size_t)b * G.stride;
    const int nch = chhi - chlo;                                      // 32-column chunks: panels of the range, or block column k itself
    const double* Lt = G.linv + (size_t)b * 4096;
    auto load = [&](int stage, int ch) {
        double* As = sm + stage * BU_STAGE;
        double* Bs = As + BU_KC * BU_LDA;
        const size_t col0 = (size_t)(chlo + ch) * BU_KC + (TRSM ? (size_t)k * 64 : 0);
        for (int e = tid; e < BU_KC * (ROWS / 2); e += 256) {         // A: 32 columns x ROWS / 2 chunks of 2 rows
            const int kk = e / (ROWS / 2), r2 = (e % (ROWS / 2)) * 2;
            double* dst = As + kk * BU_LDA + r2;
            if (r2 < rows_here) cp_async16(dst, Ab + (col0 + kk) * G.nrp + row0 + r2);
            else { dst[0] = 0.0; dst[1] = 0.0; }
        }
        for (int e = tid; e < BU_KC * 32; e += 256) {                 // B: rows of block k
            const int kk = e >> 5, r2 = (e & 31) * 2;
            if (TRSM) cp_async16(Bs + kk * BU_LDB + r2, Lt + (size_t)(ch * BU_KC + kk) * 64 + r2);
            else cp_async16(Bs + kk * BU_LDB + r2, Ab + (col0 + kk) * G.nrp + (size_t)k * 64 + r2);
        }
        asm volatile("cp.async.commit_group;" ::: "memory");
    };
    const int warp = tid >> 5, lane = tid & 31;
    const int wr = (ROWS == 128 ? (warp >> 1) : (warp >> 2)) * 32, wc = ROWS == 128 ? (warp & 1) * 32 : (warp & 3) * 16;
    const int fr = lane >> 2, fk = lane & 3;
    const int nfrag = min(4, max(0, (rows_here - wr + 7) >> 3));     // live 8-row fragment rows of this warp's 32 rows
    double acc[4][NJF][2];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < NJF; ++j) { acc[i][j][0] = 0.0; acc[i][j][1] = 0.0; }
    load(0, 0);
    for (int ch = 0; ch < nch; ++ch) {
        if (ch + 1 < nch) {
            load((ch + 1) & 1, ch + 1);
            asm volatile("cp.async.wait_group 1;" ::: "memory");
        } else {
            asm volatile("cp.async.wait_group 0;" ::: "memory");
        }
        __syncthreads();
        const double* As = sm + (ch & 1) * BU_STAGE;
        const double* Bs = As + BU_KC * BU_LDA;
        if (nfrag == 4) {
#pragma unroll
            for (int k0 = 0; k0 < BU_KC; k0 += 4) {
                double af[4], bf[NJF];
#pragma unroll
                for (int i = 0; i < 4; ++i) af[i] = As[(k0 + fk) * BU_LDA + wr + 8 * i + fr];
#pragma unroll
                for (int j = 0; j < NJF; ++j) bf[j] = Bs[(k0 + fk) * BU_LDB + wc + 8 * j + fr];
#pragma unroll
                for (int i = 0; i < 4; ++i)
#pragma unroll
                    for (int j = 0; j < NJF; ++j) dmma884(acc[i][j][0], acc[i][j][1], af[i], bf[j]);
            }
        } else if (nfrag > 0) {                          // the last tile's partial piece: 8 live rows per fragment row
#pragma unroll
            for (int k0 = 0; k0 < BU_KC; k0 += 4) {
                double bf[NJF];
#pragma unroll
                for (int j = 0; j < NJF; ++j) bf[j] = Bs[(k0 + fk) * BU_LDB + wc + 8 * j + fr];
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                    if (i < nfrag) {                     // warp-uniform
                        const double af = As[(k0 + fk) * BU_LDA + wr + 8 * i + fr];
#pragma unroll
                        for (int j = 0; j < NJF; ++j) dmma884(acc[i][j][0], acc[i][j][1], af, bf[j]);
                    }
                }
            }
        }
        __syncthreads();
    }
    double* C = G.A + (size_t)b * G.stride + (size_t)(k * 64) * G.nrp + row0;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const int r = wr + 8 * i + fr;
        if (r < rows_here) {
#pragma unroll
            for (int j = 0; j < NJF; ++j) {
                const int c = wc + 8 * j + 2 * fk;
                if (TRSM) {
                    C[(size_t)c * G.nrp + r] = acc[i][j][0];
                    C[(size_t)(c + 1) * G.nrp + r] = acc[i][j][1];
                } else {
                    C[(size_t)c * G.nrp + r] -= acc[i][j][0];
                    C[(size_t)(c + 1) * G.nrp + r] -= acc[i][j][1];
                }
            }
        }
    }
}
// Rows per tile.  64-row tiles measured faster than 128-row tiles at every shape tried (n = 2048 x 64: 10.4 vs 11.1 ms;
// n = 1024 x 128: 3.6 vs 3.75 ms): the launches of the late block columns have few tiles per SM, and the finer tiles fill the
// last round.  CCGP_BIG_ROWS = 128 forces the tall tiles.
inline int big_update_rows(int rows, int nb, int num_sm) {
    const char* e = getenv("CCGP_BIG_ROWS");
    if (e && *e) return atoi(e) == 64 ? 64 : 128;
    (void)rows; (void)nb; (void)num_sm;
    return 64;
}
template <bool TRSM>
inline void big_update_launch(const BigArgs& G, int k, int chlo, int chhi, int rows, int nb, int num_sm, cudaStream_t st) {
    if (big_update_rows(rows, nb, num_sm) == 64)
        big_update_kernel<TRSM, 64><<<dim3((rows + 63) / 64, nb), 256, 2 * bu_stage<64>() * 8, st>>>(G, k, chlo, chhi);
    else
        big_update_kernel<TRSM, 128><<<dim3((rows + 127) / 128, nb), 256, 2 * bu_stage<128>() * 8, st>>>(G, k, chlo, chhi);
}

__global__ void __launch_bounds__(256) big_finish_kernel(BigArgs G, int64_t b0, double sigma2, int mean_mode, double tau,
                                                        double* out_nll, double* out_beta, int32_t* out_status) {
    __shared__ double red[64];
    const int b = blockIdx.x, tid = threadIdx.x, n = G.n;
    const double* Ab = G.A + (size_t)b * G.stride;
    double s11 = 0.0, s1y = 0.0;
    for (int kk = tid; kk < n; kk += 256) {
        const double zy = Ab[(size_t)kk * G.nrp + G.ncp], z1 = Ab[(size_t)kk * G.nrp + G.ncp + 1];
        s11 = fma(z1, z1, s11);
        s1y = fma(z1, zy, s1y);
    }
    team_sum2<256>(s11, s1y, red);
    const double beta = s1y / s11;
    double qr = 0.0, dummy = 0.0;
    for (int kk = tid; kk < n; kk += 256) {
        const double zy = Ab[(size_t)kk * G.nrp + G.ncp], z1 = Ab[(size_t)kk * G.nrp + G.ncp + 1];
        const double rz = fma(-beta, z1, zy);
        qr = fma(rz, rz, qr);
    }
    team_sum2<256>(qr, dummy, red);
    if (tid == 0) {
        const double c = G.prm[b].c;
        const double logdet = G.logdet[b];
        double nll;
        if (mean_mode == 0) nll = 0.5 * (qr / c + n * LOG2PI + n * log(c) + logdet);
        else {
            const double g = 1.0 + tau * tau * s11 / c;
            nll = 0.5 * (qr / c + s1y * s1y / (c * s11 * g) + n * LOG2PI + n * log(c) + logdet + log(g));
        }
        const bool bad = G.bad[b] || !(nll == nll);
        const double nanv = __longlong_as_double(0x7ff8000000000000LL);
        out_nll[b0 + b] = bad ? nanv : nll;
        if (out_beta) out_beta[b0 + b] = bad ? nanv : beta;
        if (out_status) out_status[b0 + b] = bad ? 1 : 0;
    }
}

// determinant mode: log det of every candidate's matrix (the sum of log pivots big_potrf_kernel accumulated)
__global__ void __launch_bounds__(128) big_logdet_out_kernel(BigArgs G, int nb, double* out_logdet, int32_t* out_status) {
    const int b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= nb) return;
    const bool bad = G.bad[b] || !(G.logdet[b] == G.logdet[b]);
    out_logdet[G.b0 + b] = bad ? __longlong_as_double(0x7ff8000000000000LL) : G.logdet[b];
    if (out_status) out_status[G.b0 + b] = bad ? 1 : 0;
}

// d_idx != NULL: determinant mode over index subsets (d_X is then the pool with leading dimension ldx, d_cand ONE
// shared natural-scale parameter row, d_nll receives log det R[S,S]); otherwise the likelihood of the shared design.
inline int bigchol_nll_batch(BigCholWorkspace& ws, cudaStream_t stream, int num_sm, const double* d_X, const double* d_y,
                             int n, int d, int family, int scale, const double* d_cand, int64_t B, int64_t ldc,
                             double sigma2, int mean_mode, double tau, double* d_nll, double* d_beta, int32_t* d_status,
                             int64_t* launches, char* err, size_t errlen,
                             const int32_t* d_idx = nullptr, int64_t ldi = 0, int64_t ldx = 0) {
#define BIGCK(call) do { cudaError_t e_ = (call); if (e_ != cudaSuccess) { \
        snprintf(err, errlen, "bigchol %s: %s", #call, cudaGetErrorString(e_)); return -2; } } while (0)
    const bool right_looking = getenv("CCGP_BIG_RIGHT") && atoi(getenv("CCGP_BIG_RIGHT"));   // the old schedule, for A/B runs
    const int ncp = (n + 63) / 64 * 64, nrp = ncp + (right_looking ? 64 : 8), T = ncp / 64;
    const size_t per = (size_t)nrp * ncp * 8;
    size_t freeb = 0, totalb = 0;
    BIGCK(cudaMemGetInfo(&freeb, &totalb));
    size_t budget = std::min<size_t>((freeb + ws.bytesA) / 2, (size_t)24 << 30);
    int chunk = (int)std::min<int64_t>(std::min<int64_t>(B, 65535), std::max<size_t>(1, budget / per));   // (grid.z / grid.y limit)
    if (chunk < 1 || per > budget) { snprintf(err, errlen, "n=%d does not fit in device memory", n); return -3; }
    if ((size_t)chunk * per > ws.bytesA || chunk > ws.cap) {
        ws.release();
        BIGCK(cudaMalloc(&ws.A, (size_t)chunk * per));
        BIGCK(cudaMalloc(&ws.logdet, (size_t)chunk * 8));
        BIGCK(cudaMalloc(&ws.bad, (size_t)chunk * 4));
        BIGCK(cudaMalloc(&ws.prm, (size_t)chunk * sizeof(Prm)));
        BIGCK(cudaMalloc(&ws.linv, (size_t)chunk * 4096 * 8));
        ws.bytesA = (size_t)chunk * per;
        ws.cap = chunk;
    }
    BIGCK(cudaFuncSetAttribute(big_syrk_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 2 * 64 * 72 * 8));
    BIGCK(cudaFuncSetAttribute(big_update_kernel<false, 128>, cudaFuncAttributeMaxDynamicSharedMemorySize, 2 * bu_stage<128>() * 8));
    BIGCK(cudaFuncSetAttribute(big_update_kernel<true, 128>, cudaFuncAttributeMaxDynamicSharedMemorySize, 2 * bu_stage<128>() * 8));
    BIGCK(cudaFuncSetAttribute(big_update_kernel<false, 64>, cudaFuncAttributeMaxDynamicSharedMemorySize, 2 * bu_stage<64>() * 8));
    BIGCK(cudaFuncSetAttribute(big_update_kernel<true, 64>, cudaFuncAttributeMaxDynamicSharedMemorySize, 2 * bu_stage<64>() * 8));
    BIGCK(cudaFuncSetAttribute(big_trsm_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 2 * 64 * 65 * 8));
    FactorArgs F;
    memset(&F, 0, sizeof(F));
    F.d = d; F.cand = d_cand; F.ldc = ldc; F.n_params = B; F.family = family; F.logscale = scale; F.sigma2 = sigma2;
    F.force_clamp = 1;
    F.shared_row = d_idx ? 1 : 0;
    BigArgs G;
    G.A = ws.A; G.n = n; G.d = d; G.ncp = ncp; G.nrp = nrp; G.stride = (int64_t)nrp * ncp;
    G.X = d_X; G.y = d_y; G.prm = ws.prm; G.logdet = ws.logdet; G.bad = ws.bad; G.linv = ws.linv;
    G.idx = d_idx; G.ldi = ldi; G.ldx = d_idx ? ldx : n; G.b0 = 0;
    const bool potrf_old = getenv("CCGP_BIG_POTRF_OLD") && atoi(getenv("CCGP_BIG_POTRF_OLD"));
    const char* la_env = getenv("CCGP_BIG_LOOKAHEAD");
    // Lookahead (left-looking): column k's update = panels 0..k-1.  Everything but the last panel only needs columns
    // <= k-2, so it is launched as soon as column k-2 is final and overlaps the short serial kernels of column k-1
    // (last-panel update, 64x64 factor + inverse on nb CTAs, solve).  Those run on a HIGH-PRIORITY internal stream so that
    // their few CTAs take the first slots the bulk update's CTAs free; the bulk stays on the caller's stream.
    // On from 24 block columns (n = 1024 x 128 got slower with it), CCGP_BIG_LOOKAHEAD = 0 / 1 forces it.
    const bool lookahead = !right_looking && ((la_env && *la_env) ? atoi(la_env) != 0 : T >= 24);
    if (lookahead && !ws.side) {
        int lo = 0, hi = 0;
        BIGCK(cudaDeviceGetStreamPriorityRange(&lo, &hi));
        BIGCK(cudaStreamCreateWithPriority(&ws.side, cudaStreamNonBlocking, hi));
        BIGCK(cudaEventCreateWithFlags(&ws.ev_main, cudaEventDisableTiming));
        BIGCK(cudaEventCreateWithFlags(&ws.ev_side[0], cudaEventDisableTiming));
        BIGCK(cudaEventCreateWithFlags(&ws.ev_side[1], cudaEventDisableTiming));
    }
    // every launch of one chunk of candidates (b0 .. b0 + nb), in stream order; returns the number of launches or < 0
    auto enqueue_chunk = [&](int64_t b0, int nb) -> int {
        int nl = 0;
        G.b0 = b0;
        big_params_kernel<<<(nb + 127) / 128, 128, 0, stream>>>(F, b0, nb, ws.prm, ws.logdet, ws.bad);
        big_build_kernel<<<dim3(T, T + 1, nb), 256, 0, stream>>>(G);
        nl += 2;
        cudaStream_t crit = lookahead ? ws.side : stream;       // the serial chain of every block column
        if (lookahead) {                                        // the chain starts after the build
            BIGCK(cudaEventRecord(ws.ev_side[0], stream));
            BIGCK(cudaStreamWaitEvent(crit, ws.ev_side[0], 0));
        }
        for (int k = 0; k < T; ++k) {
            if (!right_looking && k > 0) {
                const bool split = lookahead && k >= 2;
                if (split) BIGCK(cudaStreamWaitEvent(crit, ws.ev_side[k & 1], 0));     // panels 0..k-2 are in (caller's stream)
                big_update_launch<false>(G, k, split ? 2 * (k - 1) : 0, 2 * k, nrp - k * 64, nb, num_sm, crit);
                nl += 1;
            }
            if (potrf_old) big_potrf_kernel<<<nb, 256, 0, crit>>>(G, k);
            else big_potrf_mma_kernel<<<nb, 256, 0, crit>>>(G, k);
            const int rows_below = nrp - (k + 1) * 64;          // rows below the diagonal block (>= 8: the extra rows)
            const int rt = rows_below / 64;                     // (64-row tiles of the right-looking schedule)
            if (rt > 0 && right_looking) big_trsm_kernel<<<dim3(rt, nb), 64, 2 * 64 * 65 * 8, crit>>>(G, k);
            else if (rows_below > 0) big_update_launch<true>(G, k, 0, 2, rows_below, nb, num_sm, crit);
            nl += 2;
            if (lookahead && k + 2 < T) {                       // column k is final: panels 0..k of column k+2, the bulk
                BIGCK(cudaEventRecord(ws.ev_main, crit));
                BIGCK(cudaStreamWaitEvent(stream, ws.ev_main, 0));
                big_update_launch<false>(G, k + 2, 0, 2 * (k + 1), nrp - (k + 2) * 64, nb, num_sm, stream);
                BIGCK(cudaEventRecord(ws.ev_side[k & 1], stream));              // (k + 2) & 1: waited on at step k + 2, re-recorded after that wait
                nl += 1;
            }
            const int ct = T - (k + 1);                         // real block columns still to update
            if (right_looking && ct > 0) {
                // tiles (ti, tj) with tj < ct, ti in [tj, rt): sum_{tj<ct} (rt - tj)
                const int ntiles = ct * rt - ct * (ct - 1) / 2;
                big_syrk_kernel<<<dim3(ntiles, nb), 128, 2 * 64 * 72 * 8, stream>>>(G, k, rt);
                nl += 1;
            }
        }
        if (lookahead) {                                        // back on the caller's stream
            BIGCK(cudaEventRecord(ws.ev_main, crit));
            BIGCK(cudaStreamWaitEvent(stream, ws.ev_main, 0));
        }
        if (d_idx) big_logdet_out_kernel<<<(nb + 127) / 128, 128, 0, stream>>>(G, nb, d_nll, d_status);
        else big_finish_kernel<<<nb, 256, 0, stream>>>(G, b0, sigma2, mean_mode, tau, d_nll, d_beta, d_status);
        nl += 1;
        return nl;
    };
    // A batch is ~130 short launches on two streams: a caller that repeats the same call (same buffers, same sizes -- the
    // host-pointer API always does, it works out of the context's workspace) gets the schedule as a CUDA graph from the second
    // identical call on: one cudaGraphLaunch instead of ~130 launches + 2 x 60 event operations, so a busy host no longer
    // starves the GPU between the 75 us kernels.  CCGP_BIG_GRAPH = 0 turns it off.
    BigGraphKey key;
    memset(&key, 0, sizeof(key));
    key.A = ws.A; key.prm = ws.prm; key.linv = ws.linv; key.X = d_X; key.y = d_y; key.cand = d_cand; key.nll = d_nll; key.beta = d_beta;
    key.status = d_status; key.idx = d_idx; key.B = B; key.ldc = ldc; key.ldi = ldi; key.ldx = ldx; key.sigma2 = sigma2; key.tau = tau;
    key.n = n; key.d = d; key.family = family; key.scale = scale; key.mean_mode = mean_mode; key.nrp = nrp;
    key.flags = (right_looking ? 1 : 0) | (lookahead ? 2 : 0) | (potrf_old ? 4 : 0) | (big_update_rows(0, 0, 0) == 64 ? 8 : 0);
    key.stream = stream;
    const char* g_env = getenv("CCGP_BIG_GRAPH");
    const bool graphs = !(g_env && *g_env && atoi(g_env) == 0) && B <= chunk && stream != nullptr;
    if (graphs && ws.gexec && memcmp(&key, &ws.gkey, sizeof(key)) == 0) {
        BIGCK(cudaGraphLaunch(ws.gexec, stream));
        *launches += ws.glaunches;
        return 0;
    }
    if (graphs && ws.have_last && memcmp(&key, &ws.last_key, sizeof(key)) == 0) {      // second identical call: capture, keep, launch
        cudaGraph_t graph = nullptr;
        if (cudaStreamBeginCapture(stream, cudaStreamCaptureModeThreadLocal) == cudaSuccess) {
            const int nl = enqueue_chunk(0, (int)B);
            const cudaError_t ee = cudaStreamEndCapture(stream, &graph);
            cudaGraphExec_t ge = nullptr;
            if (nl > 0 && ee == cudaSuccess && graph && cudaGraphInstantiate(&ge, graph, 0) == cudaSuccess) {
                if (ws.gexec) cudaGraphExecDestroy(ws.gexec);
                ws.gexec = ge; ws.gkey = key; ws.glaunches = nl;
                cudaGraphDestroy(graph);
                BIGCK(cudaGraphLaunch(ws.gexec, stream));
                *launches += nl;
                return 0;
            }
            if (graph) cudaGraphDestroy(graph);
            cudaGetLastError();                                 // capture failed: fall through to the plain launches
        } else {
            cudaGetLastError();
        }
    }
    ws.last_key = key; ws.have_last = true;
    for (int64_t b0 = 0; b0 < B; b0 += chunk) {
        const int nb = (int)std::min<int64_t>(chunk, B - b0);
        const int nl = enqueue_chunk(b0, nb);
        if (nl < 0) return nl;
        *launches += nl;
        BIGCK(cudaGetLastError());
    }
#undef BIGCK
    return 0;
}

}  // namespace ccgp
