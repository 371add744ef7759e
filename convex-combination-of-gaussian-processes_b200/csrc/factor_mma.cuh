// factor_mma.cuh -- the FP64 tensor-core (DMMA) variant of the fused build + Cholesky + solves
// kernel of factor_engine.cuh: same inputs (FactorArgs), same outputs, same arithmetic contract.
//
// Replaces, per candidate, the body of `logpost` up to `log.like` ([A]:444-455) / `cond.like`
// ([V]:564-575) and, in determinant mode, `Entropy` ([M]:856-861) / subset log-dets -- for the
// Gaussian component families (1-D Matern / spline designs are tiny and stay on factor_kernel).
//
// Why a second kernel: measured on B200 (tools/ubench_dmma.cu) `mma.sync.m8n8k4.f64` runs at the
// same 64 FMA/clk/SM as DFMA, on the same pipe, but ONE warp instruction carries 256 FMAs: a single
// warp with two independent accumulators saturates its sub-partition's FP64 pipe with one issue
// slot per 16 clocks.  factor_kernel's DFMA tiles spent ~2/3 of the issue slots on LDS / integer /
// control instructions and 58 % of the shared-memory bandwidth; here every update is a DMMA whose
// operands are perfectly coalesced 128-bit shared loads.
//
// Shared-memory layout ("block-column trapezoid, row-major blocks"): block column c (8 columns)
// stores rows 8c..npad-1, each row's 8 entries contiguous (64 B).  The 8x8 tile (r, c) is then 64
// consecutive doubles, and lane l's double2 at tile + 2l holds entries (row l/4, columns 2(l%4),
// 2(l%4)+1) -- which is at once
//   * the DMMA accumulator fragment of the tile (C/D operand),
//   * the A fragment (and, for L(c,J)', the B fragment) of BOTH k-halves of the 8-wide panel, if
//     the contraction index is enumerated as k = 2(l%4) + h for half h (any permutation of k is
//     legal as long as A and B agree).
// So a tile moves between shared memory and any DMMA operand slot with one conflict-free
// LDS.128/STS.128 and no shuffles, and an accumulator can be fed back as an A operand directly
// (used for the triangular solve).
//
// Schedule per candidate (NW warps: warp 0 = serial "diagonal" warp, the others update):
//   build: all warps, tile by tile (lane <-> row and column pair, 4 exponentials in flight).
//   step c = 0..NJ-1 (one CTA barrier + one producer/consumer named barrier per step):
//     diag warp : tile(c,c) (already holds C - sum_{J<c-1}) -= L(c,c-1) L(c,c-1)'   [2 DMMA]
//                 8x8 Cholesky, every lane redundantly in registers (diag of factor_engine.cuh),
//                 with the inverse of the 8x8 factor built column-per-lane alongside; publish.
//     update warps, tiles (r,c) r>c held in REGISTERS across steps:
//                 cur -= L(r,c-1) L(c,c-1)'                                        [2 DMMA/tile]
//                 lookahead: nxt = tile(r,c+1) - sum_{J<c} L(r,J) L(c+1,J)'  (overlaps the
//                 serial diagonal block; also pre-accumulates the next diagonal tile)
//                 wait for the diagonal warp;  L(r,c) = cur * inv(L_cc)'           [2 DMMA/tile]
//                 (triangular solve as a product with the explicit 8x8 inverse), store.
// Every tile is owned by a fixed warp and accumulated in a fixed order: results are bit-identical
// for any grid size, shard or GPU count.
#pragma once
#include "factor_engine.cuh"

namespace ccgp {

// D = A(8x4) B(4x8) + D on the FP64 tensor path
__device__ __forceinline__ void mma884(double& c0, double& c1, double a, double b) {
    asm("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
        : "+d"(c0), "+d"(c1)
        : "d"(a), "d"(b));
}
__device__ __forceinline__ void named_sync(int id, int count) {
    asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(count) : "memory");
}
__device__ __forceinline__ void named_arrive(int id, int count) {
    asm volatile("bar.arrive %0, %1;" ::"r"(id), "r"(count) : "memory");
}
__device__ __forceinline__ void st2(double* p, double x, double y) { *reinterpret_cast<double2*>(p) = make_double2(x, y); }
__device__ __forceinline__ double negd(double x) { return __hiloint2double(__double2hiint(x) ^ 0x80000000, __double2loint(x)); }

// tile (r, c), r >= c: 64 doubles
__device__ __forceinline__ int tile_off(int r, int c, int npad) { return blk_base(c, npad) + (r - c) * 64; }
// element (i, k) of the row-major-block layout
__device__ __forceinline__ int elem_off_rm(int i, int k, int npad) {
    const int c = k >> 3;
    return blk_base(c, npad) + (i - 8 * c) * 8 + (k & 7);
}

// shared bytes of one candidate: L | Xs[d*npx] | ys[npx] | linv[64] | red[64] | raw[2*(MAXD+2)] | etab[128] | Prm | ints[8]
inline size_t mma_smem_bytes(const Layout& l, int d) {
    size_t dbl = (size_t)l.total + (size_t)d * l.npx + l.npx + 64 + 64 + 2 * (MAXD + 2) + 128;
    return (dbl * 8 + sizeof(Prm) + 32 + 15) / 16 * 16;
}

// ---- assemble all tiles of A = [R; y'; 1'] ------------------------------------------------------
template <int DT, bool CLAMP>
__device__ __forceinline__ void mma_build(const FactorArgs& A, double* Ls, const double* Xs, const double* ys,
                                          const Prm* prm, const double* T, int warp, int nwarps, int lane) {
    const int n = A.lay.n, npad = A.lay.npad, naug = A.lay.naug, npx = A.lay.npx, d = A.d, NJ = A.lay.NJ;
    const int NR = npad >> 3;
    const int g = lane >> 2, m = lane & 3;
    const double rho = prm->rho, a = prm->a, b = prm->b;
    double wts[DT > 0 ? DT : 1];
    if (DT > 0) {
#pragma unroll
        for (int k = 0; k < DT; ++k) wts[k] = prm->wts[k];
    }
    for (int c = 0; c < NJ; ++c) {
        const int j0 = 8 * c + 2 * m, j1 = j0 + 1;
        const int jc0 = min(j0, n - 1), jc1 = min(j1, n - 1);
        double xj0[DT > 0 ? DT : 1], xj1[DT > 0 ? DT : 1];
        if (DT > 0) {
#pragma unroll
            for (int k = 0; k < DT; ++k) { xj0[k] = Xs[k * npx + jc0]; xj1[k] = Xs[k * npx + jc1]; }
        }
        double* colbase = Ls + blk_base(c, npad) + 2 * lane;
        // the two entries (row 8r + g, columns j0, j1) of tile (r, c)
        auto entry = [&](int r, double& v0, double& v1) {
            const int i = 8 * r + g;
            const int ic = min(i, n - 1);
            double s0 = 0.0, s1 = 0.0;
            if (DT > 0) {
#pragma unroll
                for (int k = 0; k < DT; ++k) {
                    const double xi = Xs[k * npx + ic];
                    const double d0 = xi - xj0[k], d1 = xi - xj1[k];
                    s0 = fma(wts[k] * d0, d0, s0);
                    s1 = fma(wts[k] * d1, d1, s1);
                }
            } else {
                for (int k = 0; k < d; ++k) {
                    const double xi = Xs[k * npx + ic], wk = prm->wts[k];
                    const double d0 = xi - Xs[k * npx + jc0], d1 = xi - Xs[k * npx + jc1];
                    s0 = fma(wk * d0, d0, s0);
                    s1 = fma(wk * d1, d1, s1);
                }
            }
            v0 = fma(b, dexp_neg_tab_dev<CLAMP>(rho * s0, T), a * dexp_neg_tab_dev<CLAMP>(s0, T));
            v1 = fma(b, dexp_neg_tab_dev<CLAMP>(rho * s1, T), a * dexp_neg_tab_dev<CLAMP>(s1, T));
        };
        // diagonal tile / rows beyond the design: unit diagonal, zero upper part, rows y' and 1', zero padding
        auto fixup = [&](int r, double& v0, double& v1) {
            const int i = 8 * r + g;
            if (i < n) {
                if (j0 >= i) v0 = (j0 == i) ? 1.0 : 0.0;
                if (j1 >= i) v1 = (j1 == i) ? 1.0 : 0.0;
            } else if (naug && i == n) {
                v0 = (j0 < n) ? ys[j0] : 0.0;
                v1 = (j1 < n) ? ys[j1] : 0.0;
            } else if (naug && i == n + 1) {
                v0 = (j0 < n) ? 1.0 : 0.0;
                v1 = (j1 < n) ? 1.0 : 0.0;
            } else {
                v0 = 0.0; v1 = 0.0;
            }
        };
        int r = c + warp;
        // two tiles per iteration: eight exponentials in flight per lane (a lone warp has nobody
        // else to hide the table load and the FP64 latencies behind)
        for (; r + nwarps < NR; r += 2 * nwarps) {
            const int r2 = r + nwarps;
            double v0, v1, u0, u1;
            entry(r, v0, v1);
            entry(r2, u0, u1);
            if (r == c || 8 * r + 7 >= n) fixup(r, v0, v1);          // warp-uniform
            if (8 * r2 + 7 >= n) fixup(r2, u0, u1);
            st2(colbase + (r - c) * 64, v0, v1);
            st2(colbase + (r2 - c) * 64, u0, u1);
        }
        if (r < NR) {
            double v0, v1;
            entry(r, v0, v1);
            if (r == c || 8 * r + 7 >= n) fixup(r, v0, v1);
            st2(colbase + (r - c) * 64, v0, v1);
        }
    }
}

// ---- one raw tile (r, c) of A = [R; y'; 1'] straight into this lane's fragment slot ---------------
// Same arithmetic as mma_build, for kernels that never store the unfactored matrix (fused build).
template <int DT, bool CLAMP>
__device__ __forceinline__ double2 mma_raw_tile(const FactorArgs& A, const double* Xs, const double* ys, const Prm* prm,
                                                const double* T, int r, int c, int lane) {
    const int n = A.lay.n, naug = A.lay.naug, npx = A.lay.npx, d = A.d;
    const int g = lane >> 2, m = lane & 3;
    const int j0 = 8 * c + 2 * m, j1 = j0 + 1;
    const int jc0 = min(j0, n - 1), jc1 = min(j1, n - 1);
    const int i = 8 * r + g;
    const int ic = min(i, n - 1);
    double s0 = 0.0, s1 = 0.0;
    if (DT > 0) {
#pragma unroll
        for (int k = 0; k < DT; ++k) {
            const double xi = Xs[k * npx + ic], wk = prm->wts[k];
            const double d0 = xi - Xs[k * npx + jc0], d1 = xi - Xs[k * npx + jc1];
            s0 = fma(wk * d0, d0, s0);
            s1 = fma(wk * d1, d1, s1);
        }
    } else {
        for (int k = 0; k < d; ++k) {
            const double xi = Xs[k * npx + ic], wk = prm->wts[k];
            const double d0 = xi - Xs[k * npx + jc0], d1 = xi - Xs[k * npx + jc1];
            s0 = fma(wk * d0, d0, s0);
            s1 = fma(wk * d1, d1, s1);
        }
    }
    const double rho = prm->rho, a = prm->a, b = prm->b;
    double v0 = fma(b, dexp_neg_tab_dev<CLAMP>(rho * s0, T), a * dexp_neg_tab_dev<CLAMP>(s0, T));
    double v1 = fma(b, dexp_neg_tab_dev<CLAMP>(rho * s1, T), a * dexp_neg_tab_dev<CLAMP>(s1, T));
    if (r == c || 8 * r + 7 >= n) {                       // warp-uniform: diagonal tile or rows beyond the design
        if (i < n) {
            if (j0 >= i) v0 = (j0 == i) ? 1.0 : 0.0;
            if (j1 >= i) v1 = (j1 == i) ? 1.0 : 0.0;
        } else if (naug && i == n) {
            v0 = (j0 < n) ? ys[j0] : 0.0;
            v1 = (j1 < n) ? ys[j1] : 0.0;
        } else if (naug && i == n + 1) {
            v0 = (j0 < n) ? 1.0 : 0.0;
            v1 = (j1 < n) ? 1.0 : 0.0;
        } else {
            v0 = 0.0; v1 = 0.0;
        }
    }
    return make_double2(v0, v1);
}

// ---- 8x8 diagonal tile, one warp: Cholesky + inverse of the factor -------------------------------
// Every lane holds the whole lower triangle (as diag_block of factor_engine.cuh); lane j (mod 8)
// additionally builds column j of inv(L) by forward substitution, interleaved with the factor
// loop (row c of L is final when pivot c is taken, so x_c only waits for 1/L_cc).
// `blk`: the tile (row-major, 64 doubles); `linv`: inv(L) row-major, 64 doubles.
// The determinant bookkeeping is redundant in every lane (the product of the tile's pivots, folded into
// res.mant/es once per tile): res is valid in ALL lanes.  FULL: all 8 columns are design columns (no
// per-column `live` selects).  The factor is written back only when `writeback` (the tile that holds
// the rows y', 1' -- nobody else reads L_cc: the rows below are solved against its inverse).
template <bool FULL>
__device__ __forceinline__ void mma_diag_impl(const FactorArgs& A, double* blk, double* linv, int c, int lane,
                                              FactorResult& res, bool writeback) {
    const int n = A.lay.n;
    double a[8][8];
#pragma unroll
    for (int r = 0; r < 8; ++r)
#pragma unroll
        for (int k = 0; k <= r; k += 2) {
            const double2 v = ld2(blk + r * 8 + k);
            a[r][k] = v.x; a[r][k + 1] = v.y;             // (k+1 > r: an unused upper entry)
        }
    const int jl = lane & 7;
    double x[8];
    double pp = 1.0;                                      // product of the live pivots of this tile
    int minhi = 0x7fffffff;                               // smallest high word among them (sign and size in one compare)
#pragma unroll
    for (int k = 0; k < 8; ++k) {
        const double piv = a[k][k];
        const bool live = FULL || (8 * c + k) < n;
        const double ri = live ? fast_rsqrt(piv) : 0.0;
        if (live) { pp *= piv; minhi = min(minhi, __double2hiint(piv)); }
        // inverse, row k: x_k = (e_k - sum_{q<k} L_kq x_q) / L_kk   (independent of ri until the last multiply)
        double s = (jl == k) ? 1.0 : 0.0;
#pragma unroll
        for (int q = 0; q < k; ++q) s = fma(-a[k][q], x[q], s);
        x[k] = s * ri;
        a[k][k] = piv * ri;
#pragma unroll
        for (int r = k + 1; r < 8; ++r) a[r][k] *= ri;
#pragma unroll
        for (int k2 = k + 1; k2 < 8; ++k2)
#pragma unroll
            for (int r = k2; r < 8; ++r) a[r][k2] = fma(-a[r][k], a[k2][k], a[r][k2]);
    }
    if (lane < 8) {
#pragma unroll
        for (int r = 0; r < 8; ++r) linv[r * 8 + lane] = x[r];
    }
    if (writeback && lane == 8) {               // one lane writes the factor back (upper part zeroed)
#pragma unroll
        for (int r = 0; r < 8; ++r)
#pragma unroll
            for (int k = 0; k <= r; k += 2) st2(blk + r * 8 + k, a[r][k], (k + 1 <= r) ? a[r][k + 1] : 0.0);
    }
    // pivot <= PIVOT_MIN = 2^-46 (or negative, or NaN after an earlier bad pivot -- already flagged): high word test;
    // a pivot of exactly 2^-46 with a non-zero low word passes, as in `piv > PIVOT_MIN`
    if (minhi < 0x3d100000 || !(pp == pp)) res.bad = 1;   // (a NaN pivot makes the product NaN)
    if (minhi == 0x3d100000) {                  // rare: decide the boundary binade exactly
#pragma unroll
        for (int k = 0; k < 8; ++k)
            if ((FULL || (8 * c + k) < n) && !(a[k][k] * a[k][k] > PIVOT_MIN)) res.bad = 1;
    }
    prod_accum(res.mant_all, res.es_all, pp);
    if (A.tail0 > 0) {                          // determinant of the trailing pivots only (ME criterion): rare path
        if (8 * c >= A.tail0) {
            prod_accum(res.mant_tail, res.es_tail, pp);
        } else if (8 * c + 8 > A.tail0) {
            double lt = 1.0;
#pragma unroll
            for (int k = 0; k < 8; ++k)
                if (8 * c + k >= A.tail0 && (FULL || (8 * c + k) < n)) lt *= a[k][k];       // L_kk = sqrt(pivot)
            prod_accum(res.mant_tail, res.es_tail, lt * lt);
        }
    } else {
        res.mant_tail = res.mant_all; res.es_tail = res.es_all;
    }
}
__device__ __forceinline__ void mma_diag(const FactorArgs& A, double* blk, double* linv, int c, int lane,
                                         FactorResult& res) {
    const int n = A.lay.n;
    const bool writeback = (8 * c + 8 > n);     // the tile reaches the rows beyond the design (y', 1')
    if (8 * c + 8 <= n) mma_diag_impl<true>(A, blk, linv, c, lane, res, writeback);
    else mma_diag_impl<false>(A, blk, linv, c, lane, res, writeback);
}

// ---- update-warp pieces, specialised on the number NT of tiles the warp owns in the column ------
// acc[t] -= L(r_t, J) L(rb, J)' for `npan` consecutive panels starting at the one `ap`/`bp` point
// into (ap: this lane's slot of tile (r_0, J); bp: of tile (rb, J)); tiles r_t = r_0 + t NU.
// Stepping to the next panel adds `inc` doubles (the block column's height in tiles, minus the
// one tile row the trapezoid loses), and inc itself shrinks by 64.
// DIAG: also accumulate -L(rb,J) L(rb,J)' (the next diagonal tile) into dg/dg2 (two chains).
template <int NT, int MAXT, int NU, bool DIAG>
__device__ __forceinline__ void mma_panels(double2 (&acc)[MAXT], double2& dg, double2& dg2, const double* ap,
                                           const double* bp, int inc, int npan) {
    for (int J = 0; J < npan; ++J) {
        const double2 b = ld2(bp);
        const double bx = negd(b.x), by = negd(b.y);
        double2 a[NT > 0 ? NT : 1];
#pragma unroll
        for (int t = 0; t < NT; ++t) a[t] = ld2(ap + t * NU * 64);
#pragma unroll
        for (int t = 0; t < NT; ++t) mma884(acc[t].x, acc[t].y, a[t].x, bx);
        if (DIAG) mma884(dg.x, dg.y, b.x, bx);
#pragma unroll
        for (int t = 0; t < NT; ++t) mma884(acc[t].x, acc[t].y, a[t].y, by);
        if (DIAG) mma884(dg2.x, dg2.y, b.y, by);
        ap += inc; bp += inc; inc -= 64;
    }
}
template <int K, int MAXT, int NU, bool DIAG>
__device__ __forceinline__ void mma_panels_nt(int nt, double2 (&acc)[MAXT], double2& dg, double2& dg2, const double* ap,
                                              const double* bp, int inc, int npan) {
    if constexpr (K <= MAXT) {
        if (nt == K) mma_panels<K, MAXT, NU, DIAG>(acc, dg, dg2, ap, bp, inc, npan);
        else mma_panels_nt<K + 1, MAXT, NU, DIAG>(nt, acc, dg, dg2, ap, bp, inc, npan);
    }
}
// L(r_t, c) = acc[t] inv(L_cc)'  (accumulator registers are the A fragments; li = lane's slot of inv(L_cc))
template <int NT, int MAXT, int NU>
__device__ __forceinline__ void mma_solve(const double2 (&acc)[MAXT], double2 li, double* xp) {
    double2 x[NT > 0 ? NT : 1];
#pragma unroll
    for (int t = 0; t < NT; ++t) { x[t] = make_double2(0.0, 0.0); mma884(x[t].x, x[t].y, acc[t].x, li.x); }
#pragma unroll
    for (int t = 0; t < NT; ++t) mma884(x[t].x, x[t].y, acc[t].y, li.y);
#pragma unroll
    for (int t = 0; t < NT; ++t) st2(xp + t * NU * 64, x[t].x, x[t].y);
}
template <int K, int MAXT, int NU>
__device__ __forceinline__ void mma_solve_nt(int nt, const double2 (&acc)[MAXT], double2 li, double* xp) {
    if constexpr (K <= MAXT) {
        if (nt == K) mma_solve<K, MAXT, NU>(acc, li, xp);
        else mma_solve_nt<K + 1, MAXT, NU>(nt, acc, li, xp);
    }
}

__device__ __forceinline__ void cp_async8(double* dst_smem, const double* src) {
    const unsigned d = (unsigned)__cvta_generic_to_shared(dst_smem);
    asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(d), "l"(src) : "memory");
}
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_all;" ::: "memory"); }

// phase timing (debug, tools/phase_timing_mma.py): block 0, lane 0 of the diagonal warp (slot 0..)
// and of the first update warp (slot 16..)
#define CCGP_MT(slot) do { if (A.dbg && blockIdx.x == 0 && lane == 0 && role < 2) { \
        long long t1_ = clock64(); A.dbg[role * 16 + (slot)] += t1_ - t_ph; t_ph = t1_; } } while (0)

template <int NW, int MAXT, int DT>
__global__ void __launch_bounds__(NW * 32, 4) factor_mma_kernel(const FactorArgs A) {
    constexpr int TEAM = NW * 32;
    constexpr int NU = NW - 1;
    constexpr int RAWLD = MAXD + 2;
    extern __shared__ __align__(16) double smem_all[];
    const Layout& lay = A.lay;
    double* Ls = smem_all;
    double* Xs = Ls + lay.total;
    double* ys = Xs + A.d * lay.npx;
    double* linv = ys + lay.npx;
    double* red = linv + 64;
    double* raw = red + 64;                        // 2 x RAWLD staged parameter rows
    double* etab = raw + 2 * RAWLD;                // 2^(j/128), j < 128
    Prm* prm = reinterpret_cast<Prm*>(etab + 128);
    int* ints = reinterpret_cast<int*>(reinterpret_cast<char*>(prm) + sizeof(Prm));
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int n = lay.n, npad = lay.npad, NJ = lay.NJ, NR = npad >> 3;

    // which warp takes the serial diagonal blocks: CTAs sharing an SM pick different warps, so the
    // (pipe-hungry, latency-critical) diagonal chains spread over the four SM sub-partitions
    if (tid == 0) {
        int slot = 0;
        if (A.sm_slots) {
            unsigned smid;
            asm("mov.u32 %0, %%smid;" : "=r"(smid));
            slot = atomicAdd(A.sm_slots + (smid & 1023), 1);
        }
        ints[0] = slot % NW;
    }
    for (int e = tid; e < 128; e += TEAM) etab[e] = CCGP_EXP2_TAB[e];
    if (A.design_mode == DESIGN_SHARED) {
        for (int e = tid; e < n * A.d; e += TEAM) {
            int k = e / n, i = e - k * n;
            Xs[k * lay.npx + i] = A.X[e];
        }
        if (lay.naug) for (int i = tid; i < n; i += TEAM) ys[i] = A.y[i];
    }
    // parameter rows are staged one candidate ahead with cp.async (a cold HBM read per candidate
    // otherwise stalls the whole CTA at its first barrier)
    const int nprm = A.nparams;
    if (tid < nprm && (int64_t)blockIdx.x < A.W) {
        const int64_t pi0 = (A.n_params == 1) ? 0 : (int64_t)blockIdx.x / A.n_designs;
        cp_async8(raw + tid, A.cand + pi0 + (int64_t)tid * A.ldc);
    }
    cp_async_wait_all();
    __syncthreads();
    const int fw = ints[0];
    const int role = (warp - fw + NW) % NW;         // 0: diagonal warp; 1..NU: update warp role-1
    int buf = 0;

    for (int64_t w = blockIdx.x; w < A.W; w += gridDim.x) {
        const int64_t dsg = w % A.n_designs;
        if (tid == 0) load_params_from(A, raw + buf * RAWLD, 1, prm);
        {
            const int64_t wn = w + gridDim.x;
            if (tid >= 32 && tid < 32 + nprm && wn < A.W) {
                const int64_t pin = (A.n_params == 1) ? 0 : wn / A.n_designs;
                cp_async8(raw + (buf ^ 1) * RAWLD + (tid - 32), A.cand + pin + (int64_t)(tid - 32) * A.ldc);
            }
        }
        if (A.design_mode != DESIGN_SHARED) stage_design<TEAM>(A, dsg, Xs, tid);
        __syncthreads();
        long long t_ph = (A.dbg && blockIdx.x == 0) ? clock64() : 0;

        if (prm->clamp) mma_build<DT, true>(A, Ls, Xs, ys, prm, etab, warp, NW, lane);
        else mma_build<DT, false>(A, Ls, Xs, ys, prm, etab, warp, NW, lane);
        CCGP_MT(0);
        __syncthreads();
        CCGP_MT(1);

        FactorResult res;
        res.mant_all = 1.0; res.mant_tail = 1.0; res.es_all = 0; res.es_tail = 0; res.bad = 0;

        if (role == 0) {
            // ---------------- serial warp: diagonal tiles ----------------
            for (int c = 0; c < NJ; ++c) {
                double* blk = Ls + tile_off(c, c, npad);
                if (c > 0) {
                    double2 t = ld2(blk + 2 * lane);
                    const double2 p = ld2(Ls + tile_off(c, c - 1, npad) + 2 * lane);
                    mma884(t.x, t.y, p.x, negd(p.x));
                    mma884(t.x, t.y, p.y, negd(p.y));
                    st2(blk + 2 * lane, t.x, t.y);
                    __syncwarp();
                }
                CCGP_MT(2);
                mma_diag(A, blk, linv, c, lane, res);
                __threadfence_block();
                named_arrive(1, TEAM);                     // inv(L_cc) and L_cc are published
                CCGP_MT(3);
                __syncthreads();                           // step c complete (panel c stored)
                CCGP_MT(4);
            }
        } else {
            // ---------------- update warps ----------------
            const int uw = role - 1;
            double2 cur[MAXT], nxt[MAXT];
            double2 dg = make_double2(0.0, 0.0), dg2 = make_double2(0.0, 0.0);
#pragma unroll
            for (int t = 0; t < MAXT; ++t) {
                const int r = 1 + uw + t * NU;
                cur[t] = (r < NR) ? ld2(Ls + tile_off(r, 0, npad) + 2 * lane) : make_double2(0.0, 0.0);
            }
            const double* Ll = Ls + 2 * lane;
            for (int c = 0; c < NJ; ++c) {
                const int r0 = c + 1 + uw;                 // first own tile row of column c
                const int nt0 = (NR - r0 + NU - 1) / NU;   // own tiles in column c (<= 0: none)
                // (A) last panel into the tiles of column c
                if (c > 0 && nt0 > 0)
                    mma_panels_nt<1, MAXT, NU, false>(nt0, cur, dg, dg2, Ll + tile_off(r0, c - 1, npad),
                                                      Ll + tile_off(c, c - 1, npad), 0, 1);
                // (B) lookahead: tiles of column c+1 through panel c-1 (overlaps the diagonal block)
                if (c + 1 < NJ) {
                    const int r1 = r0 + 1;
                    const int nt1 = max((NR - r1 + NU - 1) / NU, 0);
#pragma unroll
                    for (int t = 0; t < MAXT; ++t)
                        nxt[t] = (t < nt1) ? ld2(Ll + tile_off(r1 + t * NU, c + 1, npad)) : make_double2(0.0, 0.0);
                    if (c > 0) {
                        // tile (r, 0) sits at 64 r; from panel J to J+1 the same tile row moves by 8(npad - 8J) - 64
                        const double* ap = Ll + 64 * r1;
                        const double* bp = Ll + 64 * (c + 1);
                        const int inc = 8 * npad - 64;
                        if (uw == NU - 1) {                // this warp also pre-accumulates tile (c+1, c+1)
                            dg = make_double2(0.0, 0.0); dg2 = make_double2(0.0, 0.0);
                            mma_panels_nt<0, MAXT, NU, true>(nt1, nxt, dg, dg2, ap, bp, inc, c);
                            double* dp = Ls + tile_off(c + 1, c + 1, npad) + 2 * lane;
                            const double2 t0 = ld2(dp);
                            st2(dp, t0.x + (dg.x + dg2.x), t0.y + (dg.y + dg2.y));
                        } else if (nt1 > 0) {
                            mma_panels_nt<1, MAXT, NU, false>(nt1, nxt, dg, dg2, ap, bp, inc, c);
                        }
                    }
                }
                // (C) triangular solve against the diagonal block, as a product with its inverse
                CCGP_MT(2);
                named_sync(1, TEAM);
                CCGP_MT(3);
                if (nt0 > 0) mma_solve_nt<1, MAXT, NU>(nt0, cur, ld2(linv + 2 * lane), Ls + tile_off(r0, c, npad) + 2 * lane);
#pragma unroll
                for (int t = 0; t < MAXT; ++t) cur[t] = nxt[t];
                CCGP_MT(4);
                __syncthreads();                           // step c complete
                CCGP_MT(5);
            }
        }
        if (tid >= 32 && tid < 32 + nprm) cp_async_wait_all();
        buf ^= 1;

        // ---------------- scalars ----------------
        if (role == 0) {
            // res is valid in every lane of this warp (mma_diag keeps the bookkeeping redundantly)
        }
        if (A.out_mode == OUT_NLL) {
            double s11 = 0.0, s1y = 0.0;
            for (int k = tid; k < n; k += TEAM) {
                const int off = elem_off_rm(n, k, npad);
                const double zy = Ls[off], z1 = Ls[off + 8];
                s11 = fma(z1, z1, s11);
                s1y = fma(z1, zy, s1y);
            }
            team_sum2<TEAM>(s11, s1y, red);
            const double beta = s1y / s11;
            double qr = 0.0, dummy = 0.0;
            for (int k = tid; k < n; k += TEAM) {
                const int off = elem_off_rm(n, k, npad);
                const double rz = fma(-beta, Ls[off + 8], Ls[off]);
                qr = fma(rz, rz, qr);
            }
            team_sum2<TEAM>(qr, dummy, red);
            if (role == 0 && lane == 0) {
                const double cc = prm->c;
                const double logdet = log(res.mant_all) + res.es_all * LN2;
                double nll;
                if (A.mean_mode == 0) {
                    nll = 0.5 * (qr / cc + n * LOG2PI + n * log(cc) + logdet);
                } else {
                    const double gg = 1.0 + A.tau * A.tau * s11 / cc;
                    const double quad = qr / cc + s1y * s1y / (cc * s11 * gg);
                    nll = 0.5 * (quad + n * LOG2PI + n * log(cc) + logdet + log(gg));
                }
                const bool bad = res.bad || !(nll == nll);
                const double nanv = __longlong_as_double(0x7ff8000000000000LL);
                A.out0[w] = bad ? nanv : nll;
                if (A.out1) A.out1[w] = bad ? nanv : beta;
                if (A.status) A.status[w] = bad ? 1 : 0;
            }
        } else {
            if (role == 0 && lane == 0) {
                const double nanv = __longlong_as_double(0x7ff8000000000000LL);
                const bool bad = res.bad != 0;
                if (A.out0) A.out0[w] = bad ? nanv : log(res.mant_all) + res.es_all * LN2;
                if (A.out1) A.out1[w] = bad ? nanv : log(res.mant_tail) + res.es_tail * LN2;
                if (A.out2) A.out2[w] = bad ? nanv : -scalbn(res.mant_tail, res.es_tail);
                if (A.status) A.status[w] = bad ? 1 : 0;
            }
        }
        CCGP_MT(6);
        __syncthreads();                                   // candidate fully consumed; staged parameters visible
    }
}

}  // namespace ccgp
