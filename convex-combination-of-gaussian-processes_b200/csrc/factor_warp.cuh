// factor_warp.cuh -- one WARP per candidate on the FP64 tensor path: the production NLL kernel for
// designs up to n ~ 110 (every design the reference ships: 14, 50, 64, 90, 100 points).
//
// Same contract as factor_kernel (factor_engine.cuh) / factor_mma_kernel (factor_mma.cuh): per
// candidate the body of `logpost` up to `log.like` ([A]:444-455) or `cond.like` ([V]:564-575), in
// determinant mode `Entropy` ([M]:856-861) / subset log-dets; shared-memory layout and DMMA
// fragment identities as described in factor_mma.cuh.
//
// Why one warp: measured on B200, the serial 8x8 diagonal factorisation is a dependent FP64 chain
// whose latency triples when other warps queue work on the same sub-partition's FP64 pipe
// (profiles/: stall_math + stall_wait on the chain, update warps idle at the barrier).  A CTA here
// is four independent one-warp teams -- the hardware places the four warps of a CTA on the four
// SM sub-partitions -- so every candidate owns one FP64 pipe: no barriers, no flags, no
// contention, and the DMMA updates (one instruction = 256 FMAs) keep that pipe busy from a single
// instruction stream.  Left-looking by block column: all tiles of column c live in registers while
// panels 0..c-1 are applied (2 DMMA per tile and panel), then the diagonal tile is factored and
// inverted (mma_diag), and the rows below are solved as products with the inverse (2 DMMA/tile).
#pragma once
#include "factor_mma.cuh"

namespace ccgp {

// shared bytes of one team: L | Xs[d*npx] | ys[npx] | linv[64] | raw[2*(MAXD+2)] | Prm   (+ 1 KB table per CTA)
inline size_t warp_team_smem_bytes(const Layout& l, int d) {
    size_t dbl = (size_t)l.total + (size_t)d * l.npx + l.npx + 64 + 2 * (MAXD + 2);
    return (dbl * 8 + sizeof(Prm) + 15) / 16 * 16;
}
constexpr int WARP_TEAMS = 4;                   // one-warp teams per CTA
constexpr size_t WARP_CTA_EXTRA = 128 * 8;      // 2^(j/128) table

// ---- column pieces, specialised on the number NT of tiles in the column (dispatched once per step) ----
// cur[t] -= L(c+t, J) L(c, J)' for panels J = 0..npan-1; tile (c+t, J) sits 64 t after tile (c, J) = ap,
// and moving to the next panel adds `inc` doubles (inc shrinks by 64 per panel).  Tile 0 (the
// diagonal tile's row block) is its own B operand.  Software-pipelined: the operands of panel J+1
// are loaded before the DMMAs of panel J issue (a lone warp cannot hide shared-memory latency
// otherwise).  The load past the last panel reads the column itself: valid memory, value unused.
// SPLIT (short columns): separate accumulators for the two k-halves, so consecutive DMMAs are independent.
// SKIP0: tile 0 only serves as the B operand (its own update belongs to another warp).
template <int NT, int MAXT, bool SKIP0 = false>
__device__ __forceinline__ void warp_panels(double2 (&cur)[MAXT], const double* ap, int inc, int npan) {
    constexpr bool SPLIT = (NT - (SKIP0 ? 1 : 0) <= 2);
    constexpr int T0 = SKIP0 ? 1 : 0;
    double2 a[NT], an[NT];
    double2 alt[SPLIT ? NT : 1];
#pragma unroll
    for (int t = 0; t < NT; ++t) a[t] = ld2(ap + 64 * t);
    if (SPLIT) {
#pragma unroll
        for (int t = 0; t < NT; ++t) alt[t] = make_double2(0.0, 0.0);
    }
    for (int J = 0; J < npan; ++J) {
        ap += inc; inc -= 64;
#pragma unroll
        for (int t = 0; t < NT; ++t) an[t] = ld2(ap + 64 * t);
        const double bx = negd(a[0].x), by = negd(a[0].y);
#pragma unroll
        for (int t = T0; t < NT; ++t) mma884(cur[t].x, cur[t].y, a[t].x, bx);
#pragma unroll
        for (int t = T0; t < NT; ++t) {
            if (SPLIT) mma884(alt[t].x, alt[t].y, a[t].y, by);
            else mma884(cur[t].x, cur[t].y, a[t].y, by);
        }
#pragma unroll
        for (int t = 0; t < NT; ++t) a[t] = an[t];
    }
    if (SPLIT) {
#pragma unroll
        for (int t = T0; t < NT; ++t) { cur[t].x += alt[t].x; cur[t].y += alt[t].y; }
    }
}
// rows below the diagonal: L(c+t, c) = cur[t] inv(L_cc)', t = 1..NT-1 (accumulators are the A fragments)
template <int NT, int MAXT>
__device__ __forceinline__ void warp_solve(const double2 (&cur)[MAXT], double2 li, double* cb) {
    double2 x[NT > 1 ? NT : 2];
#pragma unroll
    for (int t = 1; t < NT; ++t) { x[t] = make_double2(0.0, 0.0); mma884(x[t].x, x[t].y, cur[t].x, li.x); }
#pragma unroll
    for (int t = 1; t < NT; ++t) mma884(x[t].x, x[t].y, cur[t].y, li.y);
#pragma unroll
    for (int t = 1; t < NT; ++t) st2(cb + 64 * t, x[t].x, x[t].y);
}
#define CCGP_NT_CASES(F)                                                                           \
    F(1) F(2) F(3) F(4) F(5) F(6) F(7) F(8) F(9) F(10) F(11) F(12) F(13) F(14)

// phase timing (debug, tools/phase_timing_warp.py): team 0 of block 0
#define CCGP_WT(slot) do { if (A.dbg && blockIdx.x == 0 && threadIdx.x == 0) { \
        long long t1_ = clock64(); A.dbg[slot] += t1_ - t_ph; t_ph = t1_; } } while (0)

// MAXT: even, >= tiles of the first block column (npad / 8)
template <int MAXT, int DT, int MINB>
__global__ void __launch_bounds__(WARP_TEAMS * 32, MINB) factor_warp_kernel(const FactorArgs A) {
    static_assert(MAXT % 2 == 0 && MAXT <= 14, "MAXT");
    constexpr int RAWLD = MAXD + 2;
    extern __shared__ __align__(16) double smem_all[];
    const Layout& lay = A.lay;
    const int lane = threadIdx.x & 31, team = threadIdx.x >> 5;
    double* etab = smem_all;
    double* Ls = smem_all + 128 + (size_t)team * (A.team_smem_bytes / 8);
    double* Xs = Ls + lay.total;
    double* ys = Xs + A.d * lay.npx;
    double* linv = ys + lay.npx;
    double* raw = linv + 64;
    Prm* prm = reinterpret_cast<Prm*>(raw + 2 * RAWLD);
    const int n = lay.n, npad = lay.npad, NJ = lay.NJ, NR = npad >> 3;

    for (int e = threadIdx.x; e < 128; e += WARP_TEAMS * 32) etab[e] = CCGP_EXP2_TAB[e];
    if (A.design_mode == DESIGN_SHARED) {
        for (int e = lane; e < n * A.d; e += 32) {
            int k = e / n, i = e - k * n;
            Xs[k * lay.npx + i] = A.X[e];
        }
        if (lay.naug) for (int i = lane; i < n; i += 32) ys[i] = A.y[i];
    }
    const int64_t w0 = (int64_t)blockIdx.x * WARP_TEAMS + team, wstride = (int64_t)gridDim.x * WARP_TEAMS;
    const int nprm = A.nparams;
    // parameter rows are staged one candidate ahead with cp.async (hides the HBM read)
    if (lane < nprm && w0 < A.W) {
        const int64_t pi0 = (A.n_params == 1) ? 0 : w0 / A.n_designs;
        cp_async8(raw + lane, A.cand + pi0 + (int64_t)lane * A.ldc);
    }
    cp_async_wait_all();
    __syncthreads();
    int buf = 0;
    const double* Ll = Ls + 2 * lane;

    for (int64_t w = w0; w < A.W; w += wstride) {
        const int64_t dsg = w % A.n_designs;
        long long t_top = (A.dbg && blockIdx.x == 0) ? clock64() : 0;
        if (lane == 0) load_params_from(A, raw + buf * RAWLD, 1, prm);
        {
            const int64_t wn = w + wstride;
            if (lane < nprm && wn < A.W) {
                const int64_t pin = (A.n_params == 1) ? 0 : wn / A.n_designs;
                cp_async8(raw + (buf ^ 1) * RAWLD + lane, A.cand + pin + (int64_t)lane * A.ldc);
            }
        }
        if (A.design_mode != DESIGN_SHARED) stage_design<32>(A, dsg, Xs, lane);
        __syncwarp();
        long long t_ph = t_top;
        CCGP_WT(0);

        if (prm->clamp) mma_build<DT, true>(A, Ls, Xs, ys, prm, etab, 0, 1, lane);
        else mma_build<DT, false>(A, Ls, Xs, ys, prm, etab, 0, 1, lane);
        __syncwarp();
        CCGP_WT(1);

        FactorResult res;
        res.mant_all = 1.0; res.mant_tail = 1.0; res.es_all = 0; res.es_tail = 0; res.bad = 0;

        for (int c = 0; c < NJ; ++c) {
            const int nt = NR - c;                          // tiles (c+t, c), t < nt; t = 0 is the diagonal tile
            double* cb = Ls + tile_off(c, c, npad) + 2 * lane;
            double2 cur[MAXT];
#pragma unroll
            for (int t = 0; t < MAXT; ++t) cur[t] = (t < nt) ? ld2(cb + 64 * t) : make_double2(0.0, 0.0);
            // ---- panels 0..c-1 into the column ----
            if (c > 0) {
                const double* ap = Ll + 64 * c;             // tile (c, 0)
                const int inc = 8 * npad - 64;              // tile (r, 0) -> tile (r, 1)
                switch (nt) {
#define CCGP_F(NTv) case NTv: if constexpr (NTv <= MAXT) warp_panels<NTv, MAXT>(cur, ap, inc, c); break;
                    CCGP_NT_CASES(CCGP_F)
#undef CCGP_F
                    default: break;
                }
            }
            // ---- diagonal tile: through shared memory into every lane, factor + inverse ----
            st2(cb, cur[0].x, cur[0].y);
            __syncwarp();
            CCGP_WT(2);
            mma_diag(A, Ls + tile_off(c, c, npad), linv, c, lane, res);
            __syncwarp();
            CCGP_WT(3);
            // ---- rows below: L(c+t, c) = cur[t] inv(L_cc)' ----
            {
                const double2 li = ld2(linv + 2 * lane);
                switch (nt) {
#define CCGP_F(NTv) case NTv: if constexpr (NTv <= MAXT) warp_solve<NTv, MAXT>(cur, li, cb); break;
                    CCGP_NT_CASES(CCGP_F)
#undef CCGP_F
                    default: break;
                }
            }
            __syncwarp();
            CCGP_WT(4);
        }

        // ---------------- scalars ----------------
        {
            // res is valid in every lane of this warp (mma_diag keeps the bookkeeping redundantly)
        }
        if (A.out_mode == OUT_NLL) {
            double s11 = 0.0, s1y = 0.0;
            for (int k = lane; k < n; k += 32) {
                const int off = elem_off_rm(n, k, npad);
                const double zy = Ls[off], z1 = Ls[off + 8];
                s11 = fma(z1, z1, s11);
                s1y = fma(z1, zy, s1y);
            }
            team_sum2<32>(s11, s1y, nullptr);
            const double beta = s1y / s11;
            double qr = 0.0, dummy = 0.0;
            for (int k = lane; k < n; k += 32) {
                const int off = elem_off_rm(n, k, npad);
                const double rz = fma(-beta, Ls[off + 8], Ls[off]);
                qr = fma(rz, rz, qr);
            }
            team_sum2<32>(qr, dummy, nullptr);
            if (lane == 0) {
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
        } else if (lane == 0) {
            const double nanv = __longlong_as_double(0x7ff8000000000000LL);
            const bool bad = res.bad != 0;
            if (A.out0) A.out0[w] = bad ? nanv : log(res.mant_all) + res.es_all * LN2;
            if (A.out1) A.out1[w] = bad ? nanv : log(res.mant_tail) + res.es_tail * LN2;
            if (A.out2) A.out2[w] = bad ? nanv : -scalbn(res.mant_tail, res.es_tail);
            if (A.status) A.status[w] = bad ? 1 : 0;
        }
        cp_async_wait_all();
        __syncwarp();                                       // candidate consumed; staged parameters visible
        buf ^= 1;
        CCGP_WT(5);
        if (A.dbg && blockIdx.x == 0 && threadIdx.x == 0) A.dbg[6] += 1;
    }
}


}  // namespace ccgp
