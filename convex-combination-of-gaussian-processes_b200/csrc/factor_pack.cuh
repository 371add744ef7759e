// factor_pack.cuh -- one warp per candidate on the FP64 tensor path with PACKED residency: the raw
// matrix is never stored and dead tiles are recycled, so a candidate needs (c+1)(NR-c)-1 tile slots at
// its peak instead of NR(NR+1)/2 -- n = 100: 48 slots = 24.6 KB instead of 46.6 KB -- and eight to ten
// independent candidates fit one SM (two or more per sub-partition) where factor_team_kernel holds four.
//
// Contract as factor_kernel (factor_engine.cuh): per candidate the body of `logpost` up to `log.like`
// ([A]:444-455) or `cond.like` ([V]:564-575); in determinant mode `Entropy` ([M]:856-861).  Shared design
// only (DESIGN_SHARED), Gaussian component families.  DMMA fragment identities: factor_mma.cuh.
//
// Why: measured (profiles/r01b_ncu_team_kernel.txt) the team kernel leaves the FP64 pipe idle half the
// time -- its three warps per candidate are coupled by two barriers per step and nothing else is resident
// on the sub-partition to fill the gaps.  Independent candidates are the cheapest latency hiding there is
// (no barriers, no flags); what limited them was shared memory.
//
// Left-looking by block column c, everything a warp needs at step c:
//   L(r, J), J < c, r >= c   -- stored tiles, c (NR - c) of them
//   column c itself          -- assembled just in time (2 table exponentials per entry) into the slots its
//                               solved tiles will occupy; the diagonal tile stays in registers
// Tile row c is dead once the panels of step c are applied (rows n, n+1 = z_y, z_1 are kept to the end), so
// its slots are handed to column c+1.  The slot of tile (r, J) comes from a host-built table (pack_plan),
// 32-bit element offsets in shared memory, fetched two panels ahead of the DMMAs that use them.
// Per step: assemble -> panels 0..c-1 (2 DMMA per tile and panel, operands one panel ahead) -> 8x8
// Cholesky + inverse (mma_diag) -> rows below as products with the inverse (2 DMMA per tile) -> store.
// Fixed order of accumulation => bit-identical results for any grid, shard or GPU count.
#pragma once
#include "factor_mma.cuh"

namespace ccgp {

// the same DMMA as mma884 (factor_mma.cuh), as a volatile statement: the compiler may not move it
__device__ __forceinline__ void mma884v(double& c0, double& c1, double a, double b) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                 : "+d"(c0), "+d"(c1)
                 : "d"(a), "d"(b));
}

constexpr int PACK_LD = 16;                         // leading dimension of the slot table (tile rows)
constexpr int PACK_MAXNR = 14;

// Slot table: off[J * PACK_LD + r] = element offset (slot * 64) of tile (r, J), r > J; 0 elsewhere (a valid
// address: the software-pipelined loops read one or two panels past the end and discard the values).
// Tile (r, c) is allocated when column c is assembled and released when tile row r has been applied
// (end of step r); rows >= keep_row are never released.  Returns the number of slots.
inline int pack_plan(int NR, int NJ, int keep_row, uint32_t* off /* [PACK_LD * PACK_LD] */) {
    int slot[PACK_LD][PACK_LD];
    int freel[PACK_LD * PACK_LD], nfree = 0, next = 0;
    for (int i = 0; i < PACK_LD * PACK_LD; ++i) off[i] = 0;
    for (int c = 0; c < NJ; ++c) {
        for (int r = c + 1; r < NR; ++r) {
            slot[r][c] = nfree ? freel[--nfree] : next++;
            off[c * PACK_LD + r] = (uint32_t)(slot[r][c] * 64);
        }
        if (c < keep_row)
            for (int J = 0; J < c; ++J) freel[nfree++] = slot[c][J];
    }
    return next;
}

// shared bytes of one candidate (warp): slots | diagonal scratch tile | inverse | staged rows | Prm
inline size_t pack_warp_smem_bytes(int nslots) {
    size_t dbl = (size_t)nslots * 64 + 64 + 64 + 2 * (MAXD + 2);
    return (dbl * 8 + sizeof(Prm) + 15) / 16 * 16;
}
// shared bytes of the CTA-wide part: exp table | slot table | design | response
inline size_t pack_cta_smem_bytes(const Layout& l, int d) {
    return (size_t)(128 + 128 + d * l.npx + l.npx) * 8;
}

// ---- assemble column c: the diagonal tile into registers, the tiles below into their slots ----
// Arithmetic as mma_build (direct differences, table-driven exp); two tiles per iteration so that eight
// exponentials are in flight per lane.
template <int DT, bool CLAMP>
__device__ __forceinline__ void pack_build_column(const FactorArgs& A, double* Lw, const uint32_t* tabc, const double* Xs,
                                                  const double* ys, const Prm* prm, const double* T, int c, int lane,
                                                  double2& dtile) {
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
    // tiles that need the fix-up tests: the diagonal tile and tile rows reaching beyond the design (r >= n / 8);
    // the interior tiles run a loop without them (they are most of the matrix)
    auto general_pair = [&](int r) {
        double v0, v1, u0, u1;
        entry(r, v0, v1);
        entry(r + 1, u0, u1);
        if (r == c || 8 * r + 7 >= n) fixup(r, v0, v1);            // warp-uniform
        if (8 * r + 15 >= n) fixup(r + 1, u0, u1);
        if (r == c) dtile = make_double2(v0, v1);
        else st2(Lw + tabc[r], v0, v1);
        st2(Lw + tabc[r + 1], u0, u1);
    };
    const int rint = min(n >> 3, NR);                              // interior tile rows: c < r < rint
    int r = c;
    if (r + 1 < NR) { general_pair(r); r += 2; }
#pragma unroll 1
    for (; r + 1 < rint; r += 2) {
        double v0, v1, u0, u1;
        entry(r, v0, v1);
        entry(r + 1, u0, u1);
        st2(Lw + tabc[r], v0, v1);
        st2(Lw + tabc[r + 1], u0, u1);
    }
#pragma unroll 1
    for (; r + 1 < NR; r += 2) general_pair(r);
    if (r < NR) {
        double v0, v1;
        entry(r, v0, v1);
        if (r == c || 8 * r + 7 >= n) fixup(r, v0, v1);
        if (r == c) dtile = make_double2(v0, v1);
        else st2(Lw + tabc[r], v0, v1);
    }
}

// cur[t] -= L(c+t, J) L(c, J)' for panels J = 0..npan-1.  `tp` = table + c: tile (c+t, J) sits at element
// offset tp[J * PACK_LD + t] of `Lw` (the warp's slots, already advanced by this lane's 2*lane).  Operands
// are loaded one panel ahead, their offsets two panels ahead; the reads past the last panel hit offset-0 or
// current-column slots (valid memory, values unused).  SPLIT (short columns): separate accumulators for
// the two k-halves so that consecutive DMMAs are independent (DMMA latency 26 clk, issue interval 16).
template <int NT, int MAXT>
__device__ __forceinline__ void pack_panels(double2 (&cur)[MAXT], const double* Lw, const uint32_t* tp, int npan) {
    constexpr bool SPLIT = NT <= 2;
    double2 a[NT], an[NT], alt[SPLIT ? NT : 1];
    int o[NT];
#pragma unroll
    for (int t = 0; t < NT; ++t) a[t] = ld2(Lw + tp[t]);
#pragma unroll
    for (int t = 0; t < NT; ++t) o[t] = tp[PACK_LD + t];
    if (SPLIT) {
#pragma unroll
        for (int t = 0; t < NT; ++t) alt[t] = make_double2(0.0, 0.0);
    }
    for (int J = 0; J < npan; ++J) {
        tp += PACK_LD;
#pragma unroll
        for (int t = 0; t < NT; ++t) an[t] = ld2(Lw + o[t]);
#pragma unroll
        for (int t = 0; t < NT; ++t) o[t] = tp[PACK_LD + t];
        const double bx = negd(a[0].x), by = negd(a[0].y);
#pragma unroll
        for (int t = 0; t < NT; ++t) mma884v(cur[t].x, cur[t].y, a[t].x, bx);
#pragma unroll
        for (int t = 0; t < NT; ++t) {
            if (SPLIT) mma884v(alt[t].x, alt[t].y, a[t].y, by);
            else mma884v(cur[t].x, cur[t].y, a[t].y, by);
        }
#pragma unroll
        for (int t = 0; t < NT; ++t) a[t] = an[t];
    }
    if (SPLIT) {
#pragma unroll
        for (int t = 0; t < NT; ++t) { cur[t].x += alt[t].x; cur[t].y += alt[t].y; }
    }
}

// rows below the diagonal: L(c+t, c) = cur[t] inv(L_cc)', t = 1..NT-1 (accumulators are the A fragments), into their slots
template <int NT, int MAXT>
__device__ __forceinline__ void pack_solve(const double2 (&cur)[MAXT], double2 li, double* Lw, const uint32_t* tabc) {
    double2 x[NT > 1 ? NT : 2];
#pragma unroll
    for (int t = 1; t < NT; ++t) { x[t] = make_double2(0.0, 0.0); mma884v(x[t].x, x[t].y, cur[t].x, li.x); }
#pragma unroll
    for (int t = 1; t < NT; ++t) mma884v(x[t].x, x[t].y, cur[t].y, li.y);
#pragma unroll
    for (int t = 1; t < NT; ++t) st2(Lw + tabc[t], x[t].x, x[t].y);
}

#define CCGP_PACK_NT_CASES(F)                                                                      \
    F(1) F(2) F(3) F(4) F(5) F(6) F(7) F(8) F(9) F(10) F(11) F(12) F(13) F(14)

// phase timing (debug, tools/phase_timing_pack.py): warp 0 of block 0
#define CCGP_PT(slot) do { if (A.dbg && blockIdx.x == 0 && threadIdx.x == 0) { \
        long long t1_ = clock64(); A.dbg[slot] += t1_ - t_ph; t_ph = t1_; } } while (0)

// MAXT: >= tiles of the first block column (npad / 8); blockDim.x / 32 <= MAXW independent candidates per CTA
template <int MAXT, int DT, int MAXW>
__global__ void __launch_bounds__(MAXW * 32, 1) factor_pack_kernel(const FactorArgs A) {
    static_assert(MAXT <= PACK_MAXNR, "MAXT");
    constexpr int RAWLD = MAXD + 2;
    extern __shared__ __align__(16) double smem_all[];
    const Layout& lay = A.lay;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarp = blockDim.x >> 5;
    const int n = lay.n, NJ = lay.NJ, NR = lay.npad >> 3;
    double* etab = smem_all;
    uint32_t* tab = reinterpret_cast<uint32_t*>(smem_all + 128);
    double* Xs = smem_all + 128 + 128;
    double* ys = Xs + A.d * lay.npx;
    double* Ls = ys + lay.npx + (size_t)warp * (A.team_smem_bytes / 8);
    double* dscr = Ls + A.pack_slots * 64;
    double* linv = dscr + 64;
    double* raw = linv + 64;
    Prm* prm = reinterpret_cast<Prm*>(raw + 2 * RAWLD);

    for (int e = threadIdx.x; e < 128; e += blockDim.x) etab[e] = CCGP_EXP2_TAB[e];
    for (int e = threadIdx.x; e < PACK_LD * PACK_LD; e += blockDim.x) tab[e] = A.pack_off[e];
    for (int e = threadIdx.x; e < n * A.d; e += blockDim.x) {
        int k = e / n, i = e - k * n;
        Xs[k * lay.npx + i] = A.X[e];
    }
    if (lay.naug) for (int i = threadIdx.x; i < n; i += blockDim.x) ys[i] = A.y[i];
    const int64_t w0 = (int64_t)blockIdx.x * nwarp + warp, wstride = (int64_t)gridDim.x * nwarp;
    const int nprm = A.nparams;
    // parameter rows are staged one candidate ahead with cp.async (hides the HBM read)
    if (lane < nprm && w0 < A.W) {
        const int64_t pi0 = (A.n_params == 1) ? 0 : w0 / A.n_designs;
        cp_async8(raw + lane, A.cand + pi0 + (int64_t)lane * A.ldc);
    }
    cp_async_wait_all();
    __syncthreads();
    int buf = 0;
    double* Lw = Ls + 2 * lane;

    for (int64_t w = w0; w < A.W; w += wstride) {
        long long t_ph = (A.dbg && blockIdx.x == 0) ? clock64() : 0;
        if (lane == 0) load_params_from(A, raw + buf * RAWLD, 1, prm);
        {
            const int64_t wn = w + wstride;
            if (lane < nprm && wn < A.W) {
                const int64_t pin = (A.n_params == 1) ? 0 : wn / A.n_designs;
                cp_async8(raw + (buf ^ 1) * RAWLD + lane, A.cand + pin + (int64_t)lane * A.ldc);
            }
        }
        __syncwarp();
        CCGP_PT(0);
        const bool clamp = prm->clamp != 0;

        FactorResult res;
        res.mant_all = 1.0; res.mant_tail = 1.0; res.es_all = 0; res.es_tail = 0; res.bad = 0;

        for (int c = 0; c < NJ; ++c) {
            const int nt = NR - c;                          // tiles (c+t, c), t < nt; t = 0 is the diagonal tile
            double2 dtile;
            if (clamp) pack_build_column<DT, true>(A, Lw, tab + c * PACK_LD, Xs, ys, prm, etab, c, lane, dtile);
            else pack_build_column<DT, false>(A, Lw, tab + c * PACK_LD, Xs, ys, prm, etab, c, lane, dtile);
            CCGP_PT(1);
            // the column's tiles into registers (the diagonal tile is already there)
            const uint32_t* tabc = tab + c * PACK_LD + c;      // tabc[t]: tile (c+t, c)
            double2 cur[MAXT];
            cur[0] = dtile;
#pragma unroll
            for (int t = 1; t < MAXT; ++t) cur[t] = (t < nt) ? ld2(Lw + tabc[t]) : make_double2(0.0, 0.0);
            if (c > 0) {
                switch (nt) {
#define CCGP_F(NTv) case NTv: if constexpr (NTv <= MAXT) pack_panels<NTv, MAXT>(cur, Lw, tab + c, c); break;
                    CCGP_PACK_NT_CASES(CCGP_F)
#undef CCGP_F
                    default: break;
                }
            }
            // diagonal tile: through shared memory into every lane, factor + inverse (one copy of the routine
            // for all column heights: it is ~1 k instructions, and eight warps at different steps share the
            // instruction cache)
            st2(dscr + 2 * lane, cur[0].x, cur[0].y);
            __syncwarp();
            mma_diag(A, dscr, linv, c, lane, res);
            __syncwarp();
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
            CCGP_PT(2);
        }

        // ---------------- scalars (res is valid in every lane: mma_diag keeps the bookkeeping redundantly) ----------------
        if (A.out_mode == OUT_NLL) {
            // element (row, k) of the solved rows z_y (row n) and z_1 (row n+1): a stored tile, or the last
            // diagonal tile (written back to the scratch tile by mma_diag)
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
        cp_async_wait_all();
        __syncwarp();                                       // candidate consumed; staged parameters visible
        buf ^= 1;
        CCGP_PT(3);
        if (A.dbg && blockIdx.x == 0 && threadIdx.x == 0) A.dbg[6] += 1;
    }
}

}  // namespace ccgp
