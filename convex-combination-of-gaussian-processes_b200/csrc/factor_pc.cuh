// factor_pc.cuh -- the packed-residency NLL kernel (factor_pack.cuh) split into PRODUCER and CONSUMER warps.
//
// Contract as factor_kernel (factor_engine.cuh): per candidate the body of `logpost` up to `log.like`
// ([A]:444-455) or `cond.like` ([V]:564-575); in determinant mode `Entropy` ([M]:856-861).  Shared design
// only (DESIGN_SHARED), Gaussian component families.  Same arithmetic, tile by tile and in the same order, as
// factor_pack_kernel: results are bit-identical to it for any grid, shard or GPU count.
//
// Why: measured (profiles/r02_ncu_pack_kernel.txt) with 8 packed candidates per SM the FP64 pipe is still idle
// half the time -- two warps per scheduler cannot cover the dependency latencies of the column assembly (chains of
// ~10 FP64 operations per exponential, 40 % of the stall samples) and of the 8x8 chain, and shared memory rules
// out a third candidate per scheduler.  A third WARP per scheduler costs no shared memory if it works on the
// candidates that are already resident: the assembly of block column c+1 (exponentials only, no dependence on
// the factor) is taken out of the candidate's own instruction stream and given to a producer warp that runs one
// column ahead.
//
// CTA = 8 consumer warps (one candidate each, two per sub-partition) + 4 producer warps (one per sub-partition;
// producer p serves consumers p and p+4, alternating column by column).  Registers follow the roles
// (setmaxnreg): consumers keep the whole block column as DMMA accumulators, producers need few.
//   producer, column c of candidate k:  wait LOADED -> [c == 0: transform the parameter row] -> raw tiles (r, c),
//                                       r >= c, into their slots (diagonal tile: a scratch tile) -> arrive FULL
//   consumer, step c:                   wait FULL -> raw column into registers -> arrive LOADED -> panels 0..c-1 ->
//                                       8x8 Cholesky + inverse -> rows below as products with the inverse -> store
// LOADED(c) releases the raw slots of column c (they are in registers now), so the producer writes column c+1
// during the whole of step c; after the last column it is the go for column 0 of the consumer's next candidate.
// Slot plan (pc_plan): raw column c+1 may not touch tile row c (still read by the panels of step c) nor the slots
// the solved column c will take; the solved tiles of the rows that outlive the last step may not touch the raw
// slots of column 0 (written for the next candidate during that step).  n = 100: 48 slots, as factor_pack_kernel.
#pragma once
#include "factor_pack.cuh"

namespace ccgp {

constexpr int PC_NC = 8;                            // consumer warps = candidates in flight per CTA
constexpr int PC_NP = 4;                            // producer warps
constexpr int PC_REGS_CONSUMER = 192;               // 256 * 192 + 128 * 112 = 63488 of the 65536 registers (an exact fit never gets its allocation)
constexpr int PC_REGS_PRODUCER = 112;

template <int N>
__device__ __forceinline__ void setmaxnreg_inc() { asm volatile("setmaxnreg.inc.sync.aligned.u32 %0;" ::"n"(N)); }
template <int N>
__device__ __forceinline__ void setmaxnreg_dec() { asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(N)); }

// Slot tables (element offsets, slot * 64): off[J * PACK_LD + r] solved tile (r, J); rawoff[c * PACK_LD + r] raw tile
// (r, c), r > c.  Event order per step c, as the kernel's barriers enforce it:
//   raw column c consumed (LOADED) -> raw column c+1 allocated (next candidate's column 0 after the last step) ->
//   tile row c released (c < keep_row) -> solved column c allocated.
// Returns the number of slots, or -1 if the plan is not cyclic (never for NR <= PACK_MAXNR).
inline int pc_plan(int NR, int NJ, int keep_row, uint32_t* off, uint32_t* rawoff) {
    constexpr int MAXS = PACK_LD * PACK_LD;
    bool isfree[MAXS], raw0[MAXS];
    int lslot[PACK_LD][PACK_LD], rslot[PACK_LD][PACK_LD];
    int next = 0;
    for (int i = 0; i < MAXS; ++i) { off[i] = 0; rawoff[i] = 0; isfree[i] = false; raw0[i] = false; }
    auto alloc = [&](bool avoid_raw0) {
        for (int s = 0; s < next; ++s)
            if (isfree[s] && !(avoid_raw0 && raw0[s])) { isfree[s] = false; return s; }
        return next++;
    };
    for (int r = 1; r < NR; ++r) { rslot[r][0] = alloc(false); raw0[rslot[r][0]] = true; }
    for (int c = 0; c < NJ; ++c) {
        for (int r = c + 1; r < NR; ++r) isfree[rslot[r][c]] = true;
        if (c + 1 < NJ) {
            for (int r = c + 2; r < NR; ++r) rslot[r][c + 1] = alloc(false);
        } else {
            for (int r = 1; r < NR; ++r) {
                if (!isfree[rslot[r][0]]) return -1;
                isfree[rslot[r][0]] = false;
            }
        }
        if (c < keep_row)
            for (int J = 0; J < c; ++J) isfree[lslot[c][J]] = true;
        for (int r = c + 1; r < NR; ++r) lslot[r][c] = alloc(r >= NJ - 1);
    }
    for (int c = 0; c < NJ; ++c)
        for (int r = c + 1; r < NR; ++r) {
            off[c * PACK_LD + r] = (uint32_t)(lslot[r][c] * 64);
            rawoff[c * PACK_LD + r] = (uint32_t)(rslot[r][c] * 64);
        }
    return next;
}

// shared bytes of one consumer: slots | diagonal scratch tile | inverse | raw diagonal tile | staged rows x2 | Prm x2
inline size_t pc_warp_smem_bytes(int nslots) {
    size_t dbl = (size_t)nslots * 64 + 64 + 64 + 64 + 2 * (MAXD + 2);
    return (dbl * 8 + 2 * sizeof(Prm) + 15) / 16 * 16;
}
// shared bytes of the CTA-wide part: exp table | solved-slot table | raw-slot table | design | response
inline size_t pc_cta_smem_bytes(const Layout& l, int d) {
    return (size_t)(128 + 128 + 128 + d * l.npx + l.npx) * 8;
}

// ---- producer: raw tiles (r, c), r >= c, of A = [R; y'; 1'] -- arithmetic as pack_build_column / mma_build ----
template <int DT, bool CLAMP>
__device__ __forceinline__ void pc_build_column(const FactorArgs& A, double* Lw, double* draw, const uint32_t* rtc,
                                                const double* Xs, const double* ys, const Prm* prm, const double* T,
                                                int c, int hc, int lane) {
    const int n = A.lay.n, naug = A.lay.naug, npx = A.lay.npx, d = A.d, NR = A.lay.npad >> 3;
    const int g = lane >> 2, m = lane & 3;
    const double rho = prm->rho, a = prm->a, b = prm->b;
    const int j0 = 8 * c + 2 * m, j1 = j0 + 1;
    const int jc0 = min(j0, n - 1), jc1 = min(j1, n - 1);
    double wts[DT > 0 ? DT : 1], xj0[DT > 0 ? DT : 1], xj1[DT > 0 ? DT : 1];
    if (DT > 0) {
#pragma unroll
        for (int k = 0; k < DT; ++k) { wts[k] = prm->wts[k]; xj0[k] = Xs[k * npx + jc0]; xj1[k] = Xs[k * npx + jc1]; }
    }
    auto entry = [&](int r, double& v0, double& v1) {
        const int ic = min(8 * r + g, n - 1);
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
    auto dst = [&](int r) -> double* { return (r == c) ? draw : Lw + rtc[r]; };
    auto general_pair = [&](int r) {
        double v0, v1, u0, u1;
        entry(r, v0, v1);
        entry(r + 1, u0, u1);
        if (r == c || 8 * r + 7 >= n) fixup(r, v0, v1);            // warp-uniform
        if (8 * r + 15 >= n) fixup(r + 1, u0, u1);
        st2(dst(r), v0, v1);
        st2(Lw + rtc[r + 1], u0, u1);
    };
    const int rint = min(n >> 3, NR);                              // interior tile rows: c < r < rint
    int r = c + hc;                                                // the consumer builds the first hc tiles itself
    if (hc == 0 && r + 1 < NR) { general_pair(r); r += 2; }
#pragma unroll 1
    for (; r + 1 < rint; r += 2) {
        double v0, v1, u0, u1;
        entry(r, v0, v1);
        entry(r + 1, u0, u1);
        st2(Lw + rtc[r], v0, v1);
        st2(Lw + rtc[r + 1], u0, u1);
    }
#pragma unroll 1
    for (; r + 1 < NR; r += 2) general_pair(r);
    if (r < NR) {
        double v0, v1;
        entry(r, v0, v1);
        if (r == c || 8 * r + 7 >= n) fixup(r, v0, v1);
        st2(dst(r), v0, v1);
    }
}

// phase timing (debug, tools/phase_timing_pc.py): block 0, lane 0 of consumer warp 0 (slots 0..7) and of its producer (8..15)
#define CCGP_PCT(slot) do { if (A.dbg && blockIdx.x == 0 && lane == 0 && (warp == 0 || warp == PC_NC)) { \
        long long t1_ = clock64(); A.dbg[slot] += t1_ - t_ph; t_ph = t1_; } } while (0)

// barrier ids of consumer cw: FULL (producer arrives, consumer waits), LOADED (consumer arrives, producer waits)
__device__ __forceinline__ int pc_bar_full(int cw) { return 2 * cw; }
__device__ __forceinline__ int pc_bar_loaded(int cw) { return 2 * cw + 1; }

// MAXT: >= tiles of the first block column (npad / 8)
template <int MAXT, int DT>
__global__ void __launch_bounds__((PC_NC + PC_NP) * 32, 1) factor_pc_kernel(const FactorArgs A) {
    static_assert(MAXT <= PACK_MAXNR, "MAXT");
    constexpr int RAWLD = MAXD + 2;
    extern __shared__ __align__(16) double smem_all[];
    const Layout& lay = A.lay;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int n = lay.n, NJ = lay.NJ, NR = lay.npad >> 3;
    double* etab = smem_all;
    uint32_t* tab = reinterpret_cast<uint32_t*>(smem_all + 128);
    uint32_t* rtab = reinterpret_cast<uint32_t*>(smem_all + 256);
    double* Xs = smem_all + 384;
    double* ys = Xs + A.d * lay.npx;
    double* wbase = ys + lay.npx;
    const int64_t wdbl = A.team_smem_bytes / 8;

    for (int e = threadIdx.x; e < 128; e += blockDim.x) etab[e] = CCGP_EXP2_TAB[e];
    for (int e = threadIdx.x; e < PACK_LD * PACK_LD; e += blockDim.x) { tab[e] = A.pack_off[e]; rtab[e] = A.pack_raw[e]; }
    for (int e = threadIdx.x; e < n * A.d; e += blockDim.x) {
        int k = e / n, i = e - k * n;
        Xs[k * lay.npx + i] = A.X[e];
    }
    if (lay.naug) for (int i = threadIdx.x; i < n; i += blockDim.x) ys[i] = A.y[i];
    __syncthreads();
    const int64_t wstride = (int64_t)gridDim.x * PC_NC;
    const int nprm = A.nparams;

    if (warp >= PC_NC) {
        // =========================== producer ===========================
        setmaxnreg_dec<PC_REGS_PRODUCER>();
        const int pw = warp - PC_NC;
        int64_t wk[2];
        double *Lw[2], *draw[2], *raw[2];
        Prm* prm2[2];
        int cwv[2];
#pragma unroll
        for (int h = 0; h < 2; ++h) {
            const int cw = pw + PC_NP * h;
            cwv[h] = cw;
            double* Ls = wbase + (size_t)cw * wdbl;
            Lw[h] = Ls + 2 * lane;
            draw[h] = Ls + A.pack_slots * 64 + 128 + 2 * lane;
            raw[h] = Ls + A.pack_slots * 64 + 192;
            prm2[h] = reinterpret_cast<Prm*>(raw[h] + 2 * RAWLD);
            wk[h] = (int64_t)blockIdx.x * PC_NC + cw;
            if (lane < nprm && wk[h] < A.W) {
                const int64_t pi0 = (A.n_params == 1) ? 0 : wk[h] / A.n_designs;
                cp_async8(raw[h] + lane, A.cand + pi0 + (int64_t)lane * A.ldc);
            }
        }
        const int hc = A.team_map;
        long long t_ph = (A.dbg && blockIdx.x == 0) ? clock64() : 0;
        for (int buf = 0; wk[0] < A.W; buf ^= 1) {
            cp_async_wait_all();
            __syncwarp();
#pragma unroll
            for (int h = 0; h < 2; ++h) {                 // parameter rows one candidate ahead (hides the HBM read)
                const int64_t wn = wk[h] + wstride;
                if (lane < nprm && wn < A.W) {
                    const int64_t pin = (A.n_params == 1) ? 0 : wn / A.n_designs;
                    cp_async8(raw[h] + (buf ^ 1) * RAWLD + lane, A.cand + pin + (int64_t)lane * A.ldc);
                }
            }
            for (int c = 0; c < NJ; ++c) {
#pragma unroll
                for (int h = 0; h < 2; ++h) {
                    if (wk[h] >= A.W) continue;
                    Prm* prm = prm2[h] + buf;
                    CCGP_PCT(8);
                    named_sync(pc_bar_loaded(cwv[h]), 64);
                    CCGP_PCT(9 + h);
                    if (c == 0) {
                        if (lane == 0) load_params_from(A, raw[h] + buf * RAWLD, 1, prm);
                        __syncwarp();
                        CCGP_PCT(11);
                    }
                    const uint32_t* rtc = rtab + c * PACK_LD;
                    if (prm->clamp) pc_build_column<DT, true>(A, Lw[h], draw[h], rtc, Xs, ys, prm, etab, c, hc, lane);
                    else pc_build_column<DT, false>(A, Lw[h], draw[h], rtc, Xs, ys, prm, etab, c, hc, lane);
                    __threadfence_block();
                    named_arrive(pc_bar_full(cwv[h]), 64);
                    CCGP_PCT(12);
                }
            }
            wk[0] += wstride; wk[1] += wstride;
        }
        cp_async_wait_all();
        return;
    }

    // =========================== consumer ===========================
    setmaxnreg_inc<PC_REGS_CONSUMER>();
    double* Ls = wbase + (size_t)warp * wdbl;
    double* dscr = Ls + A.pack_slots * 64;
    double* linv = dscr + 64;
    double* drawt = linv + 64;
    Prm* prm2 = reinterpret_cast<Prm*>(drawt + 64 + 2 * RAWLD);
    double* Lw = Ls + 2 * lane;
    const int fullb = pc_bar_full(warp), loadb = pc_bar_loaded(warp);
    const int64_t w0 = (int64_t)blockIdx.x * PC_NC + warp;
    if (w0 < A.W) named_arrive(loadb, 64);                     // go for column 0 of the first candidate
    const int hc = A.team_map;
    long long t_ph = (A.dbg && blockIdx.x == 0) ? clock64() : 0;
    int buf = 0;
    for (int64_t w = w0; w < A.W; w += wstride, buf ^= 1) {
        const Prm* prm = prm2 + buf;
        const bool more = (w + wstride < A.W);
        FactorResult res;
        res.mant_all = 1.0; res.mant_tail = 1.0; res.es_all = 0; res.es_tail = 0; res.bad = 0;

        for (int c = 0; c < NJ; ++c) {
            const int nt = NR - c;                          // tiles (c+t, c), t < nt; t = 0 is the diagonal tile
            CCGP_PCT(0);
            if (c == 0) named_sync(fullb, 64);              // (the candidate's parameters come with column 0)
            CCGP_PCT(1);
            const uint32_t* rtc = rtab + c * PACK_LD + c;      // rtc[t]: raw tile (c+t, c)
            const uint32_t* tabc = tab + c * PACK_LD + c;      // tabc[t]: solved tile (c+t, c)
            double2 cur[MAXT];
            // the consumer's own share of the assembly: the first hc (<= 2) tiles of the column, straight into registers
#pragma unroll
            for (int t = 0; t < 2; ++t)
                if (t < hc && t < nt)
                    cur[t] = prm->clamp ? mma_raw_tile<DT, true>(A, Xs, ys, prm, etab, c + t, c, lane)
                                        : mma_raw_tile<DT, false>(A, Xs, ys, prm, etab, c + t, c, lane);
            CCGP_PCT(2);
            if (c > 0) named_sync(fullb, 64);
            CCGP_PCT(1);
            if (hc < 1) cur[0] = ld2(drawt + 2 * lane);
            if (hc < 2) cur[1] = (1 < nt) ? ld2(Lw + rtc[1]) : make_double2(0.0, 0.0);
#pragma unroll
            for (int t = 2; t < MAXT; ++t) cur[t] = (t < nt) ? ld2(Lw + rtc[t]) : make_double2(0.0, 0.0);
            if (c + 1 < NJ || more) {                       // the raw slots are free: column c+1 (next candidate's column 0)
                __threadfence_block();
                named_arrive(loadb, 64);
            }
            if (c > 0) {
                switch (nt) {
#define CCGP_F(NTv) case NTv: if constexpr (NTv <= MAXT) pack_panels<NTv, MAXT>(cur, Lw, tab + c, c); break;
                    CCGP_PACK_NT_CASES(CCGP_F)
#undef CCGP_F
                    default: break;
                }
            }
            CCGP_PCT(3);
            st2(dscr + 2 * lane, cur[0].x, cur[0].y);
            __syncwarp();
            mma_diag(A, dscr, linv, c, lane, res);
            __syncwarp();
            CCGP_PCT(4);
            {
                const double2 li = ld2(linv + 2 * lane);
                switch (nt) {
#define CCGP_F(NTv) case NTv: if constexpr (NTv <= MAXT) pack_solve<NTv, MAXT>(cur, li, Lw, tabc); break;
                    CCGP_PACK_NT_CASES(CCGP_F)
#undef CCGP_F
                    default: break;
                }
            }
            __syncwarp();
            CCGP_PCT(5);
        }

        // ---------------- scalars (as factor_pack_kernel; res is valid in every lane) ----------------
        if (A.out_mode == OUT_NLL) {
            const int try_ = n >> 3, tr1 = (n + 1) >> 3;
            const int iy = (n & 7) * 8, i1 = ((n + 1) & 7) * 8;
            auto zat = [&](int tr, int rowoff, int k) -> double {
                const int J = k >> 3;
                const double* base = (tr == J) ? dscr : Ls + tab[J * PACK_LD + tr];
                return base[rowoff + (k & 7)];
            };
            double s11 = 0.0, s1y = 0.0;
            for (int k = lane; k < n; k += 32) {
                const double zy = zat(try_, iy, k), z1 = zat(tr1, i1, k);
                s11 = fma(z1, z1, s11);
                s1y = fma(z1, zy, s1y);
            }
            team_sum2<32>(s11, s1y, nullptr);
            const double beta = s1y / s11;
            double qr = 0.0, dummy = 0.0;
            for (int k = lane; k < n; k += 32) {
                const double rz = fma(-beta, zat(tr1, i1, k), zat(try_, iy, k));
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
        __syncwarp();                                       // candidate consumed
        CCGP_PCT(6);
        if (A.dbg && blockIdx.x == 0 && threadIdx.x == 0) A.dbg[7] += 1;
    }
}

}  // namespace ccgp
