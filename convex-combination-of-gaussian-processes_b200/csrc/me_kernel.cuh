// me_kernel.cuh -- specialised kernel for the maximum-entropy batch criterion
// `Augmented.Mixed.Entropy` ([M]:869-877): -det(R.new - R.cross R.old^-1 R.cross') for C
// candidate second-batch designs x P parameter rows, small sizes (n_new <= 8, n_old <= 32,
// d <= 4; the reference runs n_old = 14, n_new = 7, d = 2).
//
// The Schur complement is the trailing block of the Cholesky factor of
//   R_all = [ R.old  R.cross' ; R.cross  R.new ],
// so with R.old = L_o L_o' (factored ONCE per parameter row, [M]:924-925 computes R.old.Inv
// once per row too) each candidate needs  A = R.cross L_o^-T  (n_new forward solves),
// S = R.new - A A'  and the product of the n_new Cholesky pivots of S.
//
// Mapping: a CTA takes a contiguous range of passes (16 candidates each) of the (parameter row, candidate) grid; it
// factors R.old of every row it touches into shared memory, then every 8-lane group owns one candidate, lane r <-> new
// point r: the cross correlations and the forward solve of row r stay in that lane's registers, the solved rows are
// exchanged through a per-warp shared buffer, S is formed pair by pair (each unordered pair of new points by one lane,
// see SYM below) and the 8x8 factorisation runs on width-8 shuffles.  Deterministic: one fixed evaluation order per
// (candidate, row), whatever the schedule -- CCGP_ME_BALANCED=0 / CCGP_ME_SYM=0 (read at every launch) select the first
// version's chunked schedule / row-wise S block and give the same bits (tests/test_gpu_me_variants.py).
#pragma once
#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include "factor_engine.cuh"

#ifndef CCGP_ME_SYM_DEFAULT
#define CCGP_ME_SYM_DEFAULT true
#endif
#ifndef CCGP_ME_BALANCED_DEFAULT
#define CCGP_ME_BALANCED_DEFAULT true
#endif

namespace ccgp {

struct MeArgs {
    const double* D_old;   // n_old x d column-major
    const double* D_new;   // C blocks of n_new*d (each n_new x d column-major)
    const double* params;  // P x 3 column-major (p, theta1, theta2), ld ldq
    int64_t ldq;
    int n_old, n_new, d;
    int64_t C, P;
    int64_t chunk;         // candidates per work item
    int64_t nchunks;       // ceil(candidates per parameter row / chunk)
    int64_t ppr;           // > 0: balanced schedule -- passes (16 candidates) per parameter row; CTA b takes the
                           // contiguous range [T b / grid, T (b+1) / grid) of the P * ppr passes and factors R.old
                           // once per row it touches (same values: the per-(candidate, row) order is fixed)
    int64_t group;         // 0: every design x every parameter row; G > 0: designs [qG, (q+1)G) belong to row q (C = P G)
    int stencil;           // S > 0: D_new holds C/S base designs; design c is base c/S with the central-difference
                           // perturbation (c%S): 0 none, 1+2i: coordinate i + h, 2+2i: coordinate i - h (clipped to [lo, hi])
    double h, lo, hi;
    double* negdet;        // C x P column-major (may be NULL)
    double* logdet;        // (may be NULL)
    int32_t* status;       // (may be NULL)
};

inline bool me_fast_supported(int n_old, int n_new, int d) {
    // n_old = 0 (the first batch, `Entropy` [M]:856-861): the old block is the identity padding, the cross rows vanish
    return n_new >= 1 && n_new <= 8 && n_old >= 0 && n_old <= 32 && d >= 1 && d <= 4;
}

// NOLD: rows of the old design the kernel is unrolled for; DM: coordinates it is unrolled for (d <= DM)
// STENCIL: designs are generated from base designs (central-difference points), see MeArgs::stencil
// SYM: the symmetric block S is formed pair by pair (lane r takes the partners (r + k) mod n_new, k = 1 .. n_new/2:
// every unordered pair once, twice for k = n_new/2 when n_new is even) and exchanged through shared memory, instead
// of every lane forming its whole row; S(r, c) has the same bits whichever of the two lanes forms it (the squared
// difference and the fma chain of the dot product are symmetric), so the determinants are bit-identical
template <int NOLD, int DM, bool STENCIL, bool SYM>
__global__ void __launch_bounds__(128) me_schur_kernel(const MeArgs M) {
    __shared__ __align__(16) double sbuf[SYM ? 4 : 1][SYM ? 4 : 1][SYM ? 64 : 1];   // per 8-lane group: S, lower triangle
    __shared__ double Lo[NOLD * NOLD];        // L_old, row-major Lo[j*NOLD + k], k <= j
    __shared__ double rio[NOLD];              // 1 / L_old(j,j)
    __shared__ double Xo[DM * NOLD];          // D_old, Xo[k*NOLD + j]
    __shared__ double abuf[4][32][NOLD + 1];  // per warp: the solved cross rows (padded against conflicts)
    __shared__ double prm[4];                 // a, b, theta1, theta2
    // 2^(j/128): table-driven exp (ccgp_math.h).  TREP copies, one per lane of a half-warp (dexp_neg_tab_dev_ic<STRIDE>):
    // measured with 16 copies (profiles/r02_me_schedule_ab.txt) -- bank conflicts 41.9 M -> 14.3 M per launch, the
    // shared-memory pipe 81 % -> 66 % of its peak, and the same 0.62 ms (the kernel is dispatch-bound again), 5 % slower at
    // P = 60 where filling the 16 KB table per CTA shows: one copy kept
    constexpr int TREP = 1;
    __shared__ double etab[128 * TREP];
    __shared__ int s_bad;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int n_old = M.n_old, n_new = M.n_new, d = M.d;
    const int r = lane & 7, grp = lane >> 3;
    const unsigned full = 0xffffffffu;

    for (int e = tid; e < DM * NOLD; e += 128) {
        int k = e / NOLD, j = e - k * NOLD;
        Xo[e] = (k < d && j < n_old) ? M.D_old[k * n_old + j] : 0.0;
    }
    for (int e = tid; e < 128 * TREP; e += 128) etab[e] = CCGP_EXP2_TAB[e / TREP];
    const double* Tl = etab + (threadIdx.x & (TREP - 1));
    if (SYM) for (int e = tid; e < 4 * 4 * 64; e += 128) (&sbuf[0][0][0])[e] = 0.0;

    const bool balanced = M.ppr > 0;
    const int64_t npass = M.P * M.ppr;
    const int64_t ps0 = balanced ? npass * blockIdx.x / gridDim.x : 0, ps1 = balanced ? npass * (blockIdx.x + 1) / gridDim.x : 0;
    const int64_t nitems = balanced ? (ps1 > ps0 ? (ps1 - 1) / M.ppr + 1 : 0) : M.P * M.nchunks;
    for (int64_t item = balanced ? ps0 / M.ppr : blockIdx.x; item < nitems; item += balanced ? 1 : gridDim.x) {
        const int64_t q = balanced ? item : item / M.nchunks, ch = balanced ? 0 : item - q * M.nchunks;
        __syncthreads();
        if (tid == 0) {
            const double p = M.params[q], t1 = M.params[q + M.ldq], t2 = M.params[q + 2 * M.ldq];
            const double w = p * p + (1.0 - p) * (1.0 - p);
            prm[0] = p * p / w; prm[1] = (1.0 - p) * (1.0 - p) / w; prm[2] = t1; prm[3] = t2;
            s_bad = 0;
        }
        __syncthreads();
        const double ca = prm[0], cb = prm[1], t1 = prm[2], t2 = prm[3];
        // ---- R.old (lower) into Lo, identity padding beyond n_old ---------------------------
        for (int e = tid; e < NOLD * NOLD; e += 128) {
            int j = e / NOLD, k = e - j * NOLD;
            double v = (j == k) ? 1.0 : 0.0;
            if (k < j && j < n_old) {
                double s = 0.0;
#pragma unroll
                for (int dd = 0; dd < DM; ++dd) { double df = Xo[dd * NOLD + j] - Xo[dd * NOLD + k]; s = fma(df, df, s); }
                v = SYM ? fma(cb, dexp_neg_tab_dev_ic<TREP>(t2 * s, Tl), ca * dexp_neg_tab_dev_ic<TREP>(t1 * s, Tl))
                        : fma(cb, dexp_neg_tab_dev<true>(t2 * s, etab), ca * dexp_neg_tab_dev<true>(t1 * s, etab));
            }
            Lo[e] = v;
        }
        __syncthreads();
        // ---- Cholesky of R.old by warp 0: lane <-> row, right-looking, column by column --------
        if (warp == 0) {
            int bad = 0;
            for (int j = 0; j < n_old; ++j) {
                const double piv = Lo[j * NOLD + j];
                if (!(piv > PIVOT_MIN)) bad = 1;
                const double ri = fast_rsqrt(piv);
                __syncwarp();
                double lij = 0.0;
                if (lane < NOLD && lane > j) { lij = Lo[lane * NOLD + j] * ri; Lo[lane * NOLD + j] = lij; }
                if (lane == j) { Lo[j * NOLD + j] = piv * ri; rio[j] = ri; }
                __syncwarp();
                if (lane < NOLD && lane > j)
                    for (int k = j + 1; k <= lane; ++k) Lo[lane * NOLD + k] = fma(-lij, Lo[k * NOLD + j], Lo[lane * NOLD + k]);
                __syncwarp();
            }
            if (lane >= n_old && lane < NOLD) rio[lane] = 1.0;
            if (bad && lane == 0) s_bad = 1;
        }
        __syncthreads();
        const int bad_old = s_bad;

        // ---- candidates of this chunk: 16 per pass (4 warps x 4 groups) -----------------------
        const int64_t c_beg = M.group ? q * M.group : 0, c_end = M.group ? c_beg + M.group : M.C;
        int64_t c_lo = c_beg + ch * M.chunk, c_hi = min(c_end, c_lo + M.chunk);
        if (balanced) {                                   // the passes of row q inside [ps0, ps1)
            c_lo = c_beg + (max(ps0, q * M.ppr) - q * M.ppr) * 16;
            c_hi = min(c_end, c_beg + (min(ps1, (q + 1) * M.ppr) - q * M.ppr) * 16);
        }
        for (int64_t c0 = c_lo; c0 < c_hi; c0 += 16) {
            const int64_t c = c0 + warp * 4 + grp;
            const bool valid = c < c_hi;
            const bool rowok = valid && r < n_new;
            double x[DM];
            const int64_t cbase = STENCIL ? c / M.stencil : c;           // base design
            const int sten = STENCIL ? (int)(c - cbase * M.stencil) : 0;
#pragma unroll
            for (int dd = 0; dd < DM; ++dd) {
                x[dd] = (rowok && dd < d) ? M.D_new[cbase * (int64_t)(n_new * d) + dd * n_new + r] : 0.0;
                if (STENCIL && sten > 0 && ((sten - 1) >> 1) == dd * n_new + r)
                    x[dd] = (sten & 1) ? fmin(x[dd] + M.h, M.hi) : fmax(x[dd] - M.h, M.lo);
            }
            // cross correlations of new point r with the old design, then a = L_o^-1 (.)
            double a[NOLD];
#pragma unroll
            for (int j = 0; j < NOLD; ++j) {
                double s = 0.0;
#pragma unroll
                for (int dd = 0; dd < DM; ++dd) { double df = x[dd] - Xo[dd * NOLD + j]; s = fma(df, df, s); }
                double cv;
                if constexpr (SYM) cv = fma(cb, dexp_neg_tab_dev_ic<TREP>(t2 * s, Tl), ca * dexp_neg_tab_dev_ic<TREP>(t1 * s, Tl));
                else cv = fma(cb, dexp_neg_tab_dev<true>(t2 * s, etab), ca * dexp_neg_tab_dev<true>(t1 * s, etab));
                a[j] = (j < n_old) ? cv : 0.0;
            }
#pragma unroll
            for (int j = 0; j < NOLD; ++j) {
#pragma unroll
                for (int k = 0; k < j; ++k) a[j] = fma(-a[k], Lo[j * NOLD + k], a[j]);
                a[j] *= rio[j];
            }
            double* mine = &abuf[warp][lane][0];
#pragma unroll
            for (int j = 0; j < NOLD; ++j) mine[j] = a[j];
            __syncwarp();
            double s8[8];
            double dg = 1.0;
            if constexpr (SYM) {
                double dself = 0.0;
#pragma unroll
                for (int j = 0; j < NOLD; ++j) dself = fma(a[j], a[j], dself);
                double* Sg = &sbuf[warp][grp][0];
                const int K = n_new >> 1;
#pragma unroll
                for (int k = 1; k <= 4; ++k) {
                    if (k <= K) {
                        int c = r + k;
                        if (c >= n_new) c -= n_new;
                        if (r >= n_new) c = r;                  // padding lanes: any valid partner, nothing stored
                        double sq = 0.0;
#pragma unroll
                        for (int dd = 0; dd < DM; ++dd) {
                            const double xc = __shfl_sync(full, x[dd], c, 8);
                            const double df = x[dd] - xc;
                            sq = fma(df, df, sq);
                        }
                        double v = fma(cb, dexp_neg_tab_dev_ic<TREP>(t2 * sq, Tl), ca * dexp_neg_tab_dev_ic<TREP>(t1 * sq, Tl));
                        const double* other = &abuf[warp][(lane & ~7) + c][0];
                        double dot = 0.0;
#pragma unroll
                        for (int j = 0; j < NOLD; ++j) dot = fma(a[j], other[j], dot);
                        v -= dot;
                        if (r < n_new) Sg[max(r, c) * 8 + min(r, c)] = v;
                    }
                }
                __syncwarp();
                {
                    const double2* row = reinterpret_cast<const double2*>(Sg + r * 8);
#pragma unroll
                    for (int c2 = 0; c2 < 8; c2 += 2) { const double2 t = row[c2 >> 1]; s8[c2] = t.x; s8[c2 + 1] = t.y; }
                }
                dg = 1.0 - dself;
                if (r >= n_new) {                               // rows beyond n_new: identity padding
                    dg = 1.0;
#pragma unroll
                    for (int c2 = 0; c2 < 8; ++c2) s8[c2] = 0.0;
                }
            } else {
                // row r of S = R.new - A A' (entries c2 <= r), new-new correlations by shuffled coordinates
    #pragma unroll
                for (int c2 = 0; c2 < 8; ++c2) {
                    double sq = 0.0;
    #pragma unroll
                    for (int dd = 0; dd < DM; ++dd) {
                        const double xc = __shfl_sync(full, x[dd], c2, 8);
                        const double df = x[dd] - xc;
                        sq = fma(df, df, sq);
                    }
                    // entries c2 > r are never read and (r, r) is the unit diagonal: column n_new-1 needs no exponential
                    double v = 1.0;
                    if (c2 + 1 < n_new) v = fma(cb, dexp_neg_tab_dev<true>(t2 * sq, etab), ca * dexp_neg_tab_dev<true>(t1 * sq, etab));
                    if (c2 == r) v = 1.0;
                    const double* other = &abuf[warp][(lane & ~7) + c2][0];
                    double dot = 0.0;
    #pragma unroll
                    for (int j = 0; j < NOLD; ++j) dot = fma(a[j], other[j], dot);
                    v -= dot;
                    // rows/columns beyond n_new: identity padding
                    if (r >= n_new || c2 >= n_new) v = (c2 == r) ? 1.0 : 0.0;
                    s8[c2] = v;
                }
                __syncwarp();
#pragma unroll
                for (int c2 = 0; c2 < 8; ++c2) if (c2 == r) dg = s8[c2];
            }
            // 8x8 Cholesky inside the 8-lane group; only the pivots are needed
            double mant = 1.0;
            int es = 0, bad = bad_old;
#pragma unroll
            for (int cc = 0; cc < 8; ++cc) {
                const double piv = __shfl_sync(full, dg, cc, 8);
                if (cc < n_new) {
                    if (!(piv > PIVOT_MIN)) bad = 1;
                    prod_accum(mant, es, piv);
                }
                const double ri = fast_rsqrt(piv);
                double l = s8[cc] * ri;
                if (r > cc) dg = fma(-l, l, dg);
                double lc2[8];
#pragma unroll
                for (int c2 = cc + 1; c2 < 8; ++c2) lc2[c2] = __shfl_sync(full, l, c2, 8);
#pragma unroll
                for (int c2 = cc + 1; c2 < 8; ++c2) s8[c2] = fma(-l, lc2[c2], s8[c2]);
            }
            if (valid && r == 0) {
                const int64_t o = M.group ? c : c + M.C * q;
                const double nanv = __longlong_as_double(0x7ff8000000000000LL);
                if (M.negdet) M.negdet[o] = bad ? nanv : -scalbn(mant, es);
                if (M.logdet) M.logdet[o] = bad ? nanv : log(mant) + es * LN2;
                if (M.status) M.status[o] = bad ? 1 : 0;
            }
            __syncwarp();
        }
    }
}

inline int me_fast_launch(cudaStream_t stream, int num_sm, const double* d_D_old, int n_old, int d,
                          const double* d_D_new, int n_new, int64_t C, const double* d_params, int64_t P,
                          int64_t ldq, double* d_negdet, double* d_logdet, int32_t* d_status, char* err, size_t errlen,
                          int64_t group = 0, int stencil = 0, double h = 0.0, double lo = 0.0, double hi = 0.0) {
    MeArgs M;
    M.group = group; M.stencil = stencil; M.h = h; M.lo = lo; M.hi = hi;
    const int64_t Cq = group ? group : C;                  // candidates per parameter row
    M.D_old = d_D_old; M.D_new = d_D_new; M.params = d_params; M.ldq = ldq;
    M.n_old = n_old; M.n_new = n_new; M.d = d; M.C = C; M.P = P;
    // enough candidates per work item to amortise the R.old factorisation, enough items to fill the GPU
    int64_t chunk = 256;
    const int64_t target_items = (int64_t)num_sm * 8;
    while (chunk > 16 && P * ((Cq + chunk - 1) / chunk) < target_items) chunk >>= 1;
    M.chunk = chunk;
    M.nchunks = (Cq + chunk - 1) / chunk;
    M.negdet = d_negdet; M.logdet = d_logdet; M.status = d_status;
    const int64_t items = P * M.nchunks;
    int grid = (int)std::min<int64_t>(items, (int64_t)num_sm * 8);
    // Balanced schedule (default; CCGP_ME_BALANCED=0 restores the chunked one; see MeArgs::ppr): equal pass counts per
    // CTA.  The chunked grid above is 8 CTAs per SM against 5 resident (94 registers): its second wave runs 60 % full.
    // Grid: one wave (the occupancy query) until a CTA would hold more than ~48 passes, then more CTAs -- measured
    // (profiles/r02_me_schedule_ab.txt): a staggered second wave de-phases the CTAs' exponential / solve / shuffle
    // sections and gains another 3-4 % at P >= 1000; at P <= 250 the single wave is the fastest.
    const char* ev = getenv("CCGP_ME_BALANCED");
    const bool balanced = ev ? atoi(ev) != 0 : CCGP_ME_BALANCED_DEFAULT;
    const char* ec = getenv("CCGP_ME_CTAS");               // CTAs per SM of the balanced grid (0: the rule above)
    const int ctas_env = ec ? atoi(ec) : 0;
    M.ppr = balanced ? (Cq + 15) / 16 : 0;
    const char* es = getenv("CCGP_ME_SYM");                // 0: every lane forms its whole row of S (the first version)
    const bool sym = es ? atoi(es) != 0 : CCGP_ME_SYM_DEFAULT;
    // 14 = the reference's initial design ([M]:980); d = 2 in the shipped script
#define CCGP_ME_GO(NO, DMV, ST) do { if (sym) CCGP_ME_GO2(NO, DMV, ST, true); else CCGP_ME_GO2(NO, DMV, ST, false); } while (0)
#define CCGP_ME_GO2(NO, DMV, ST, SY) do { \
        if (balanced) { int occ = ctas_env; \
            if (occ <= 0 && cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, me_schur_kernel<NO, DMV, ST, (SY) && ((NO) <= 24)>, 128, 0) != cudaSuccess) occ = 4; \
            int64_t g = (int64_t)num_sm * std::max(occ, 1); \
            if (ctas_env <= 0) g = std::max<int64_t>(g, P * M.ppr / 48); \
            grid = (int)std::min<int64_t>(std::min<int64_t>(P * M.ppr, g), (int64_t)1 << 30); } \
        me_schur_kernel<NO, DMV, ST, (SY) && ((NO) <= 24)><<<grid, 128, 0, stream>>>(M); } while (0)
#define CCGP_ME_LAUNCH(NO) do { if (stencil) { if (d <= 2) CCGP_ME_GO(NO, 2, true); else CCGP_ME_GO(NO, 4, true); } \
                                else if (d <= 2) CCGP_ME_GO(NO, 2, false); else CCGP_ME_GO(NO, 4, false); } while (0)
    if (n_old <= 8) CCGP_ME_LAUNCH(8);
    else if (n_old <= 14) CCGP_ME_LAUNCH(14);
    else if (n_old <= 16) CCGP_ME_LAUNCH(16);
    else if (n_old <= 24) CCGP_ME_LAUNCH(24);
    else CCGP_ME_LAUNCH(32);
#undef CCGP_ME_LAUNCH
#undef CCGP_ME_GO
#undef CCGP_ME_GO2
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) { snprintf(err, errlen, "me_schur_kernel launch: %s", cudaGetErrorString(e)); return -2; }
    return 0;
}

}  // namespace ccgp
