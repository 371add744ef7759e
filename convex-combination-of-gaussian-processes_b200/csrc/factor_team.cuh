// factor_team.cuh -- NW warps per candidate, all on ONE SM sub-partition, FP64 tensor path: the
// production NLL kernel for designs up to n ~ 110 (every design the reference ships).
//
// Contract as factor_kernel (factor_engine.cuh): per candidate the body of `logpost` up to
// `log.like` ([A]:444-455) or `cond.like` ([V]:564-575); in determinant mode `Entropy`
// ([M]:856-861) / subset log-dets.  Shared-memory layout and DMMA fragment identities: factor_mma.cuh.
//
// A CTA holds 4 teams; team t is warps t, t+4, .., t+4(NW-1), which the hardware places on
// sub-partition t (measured: warp w runs on sub-partition w mod 4, tools/ubench_smsp.cu).  Each
// candidate therefore owns one FP64 pipe and one issue port: no cross-candidate contention, and
// the team's warps fill each other's latency gaps on it.  Roles inside a team, per block column c:
//   warp A (role 0)  : the serial chain.  After its 8x8 Cholesky + inverse of tile (c,c) (mma_diag) it meets
//                      the B warps at ONE barrier, then -- without waiting for their solves -- solves its own
//                      copy of tile (c+1,c) (B_0 left it, updated but unsolved, in a scratch tile), applies it
//                      to the pre-accumulated diagonal tile (c+1,c+1) straight from registers (4 DMMA) and
//                      starts the next factorisation.  At the end: the candidate's scalars.
//   warps B_u (1..NU): tiles (c+t, c) with t = 1+u, 1+u+NU, .. live in registers: last panel in (2 DMMA/tile);
//                      LOOKAHEAD -- their tiles of column c+1 through panel c-1 (the last B warp also
//                      pre-accumulates the next diagonal tile into shared memory); barrier with A; solve against inv(L_cc)
//                      (2 DMMA/tile, the inverse is double-buffered); store; a B-only barrier.
// So the chain (A) and the updates (B) of the same candidate overlap fully: the step time is max(A, B), not
// their sum.  While A reduces candidate w, B_0 transforms the parameters of w+1 (rows staged one candidate
// ahead with cp.async).  Fixed ownership and order => bit-identical results for any grid, shard or GPU count.
#pragma once
#include "factor_mma.cuh"

namespace ccgp {

constexpr int TEAMS_PER_CTA = 4;
constexpr size_t TEAM_CTA_EXTRA = 128 * 8;      // 2^(j/128) table, shared by the CTA
// shared bytes of one team: L | Xs[d*npx] | ys[npx] | linv[2][64] | usv[64] | raw[2*(MAXD+2)] | Prm[2]
inline size_t team_smem_bytes(const Layout& l, int d) {
    size_t dbl = (size_t)l.total + (size_t)d * l.npx + l.npx + 192 + 2 * (MAXD + 2);
    return (dbl * 8 + 2 * sizeof(Prm) + 15) / 16 * 16;
}

// acc[i] -= T(ap + 64 S i) T(bp)' for `npan` consecutive panels (T(p): the tile this lane's slot p
// points into; the next panel is `inc` doubles further, inc shrinking by 64 per panel).
// DIAG: dg/dg2 additionally accumulate -T(bp) T(bp)' (two chains, one per k-half).
// Software-pipelined: operands of panel J+1 are loaded before the DMMAs of panel J issue; the
// load past the last panel reads valid shared memory and is discarded.
// With at most two accumulator chains the k-halves get separate accumulators (DMMA latency 26 clk
// vs 16 clk issue interval).
template <int NT, int MAXT, int S, bool DIAG>
__device__ __forceinline__ void team_panels(double2 (&acc)[MAXT], double2& dg, double2& dg2, const double* bp,
                                            const double* ap, int inc, int npan) {
    constexpr int NA = NT > 0 ? NT : 1;
    constexpr bool SPLIT = !DIAG && NT <= 2;
    double2 a[NA], an[NA], alt[NA];
    double2 b = ld2(bp), bn;
#pragma unroll
    for (int i = 0; i < NT; ++i) { a[i] = ld2(ap + 64 * S * i); alt[i] = make_double2(0.0, 0.0); }
    for (int J = 0; J < npan; ++J) {
        bp += inc; ap += inc; inc -= 64;
        bn = ld2(bp);
#pragma unroll
        for (int i = 0; i < NT; ++i) an[i] = ld2(ap + 64 * S * i);
        const double bx = negd(b.x), by = negd(b.y);
#pragma unroll
        for (int i = 0; i < NT; ++i) mma884(acc[i].x, acc[i].y, a[i].x, bx);
        if (DIAG) mma884(dg.x, dg.y, b.x, bx);
#pragma unroll
        for (int i = 0; i < NT; ++i) {
            if (SPLIT) mma884(alt[i].x, alt[i].y, a[i].y, by);
            else mma884(acc[i].x, acc[i].y, a[i].y, by);
        }
        if (DIAG) mma884(dg2.x, dg2.y, b.y, by);
#pragma unroll
        for (int i = 0; i < NT; ++i) a[i] = an[i];
        b = bn;
    }
    if (SPLIT) {
#pragma unroll
        for (int i = 0; i < NT; ++i) { acc[i].x += alt[i].x; acc[i].y += alt[i].y; }
    }
}
template <int K, int MAXT, int S, bool DIAG>
__device__ __forceinline__ void team_panels_nt(int nt, double2 (&acc)[MAXT], double2& dg, double2& dg2, const double* bp,
                                               const double* ap, int inc, int npan) {
    if constexpr (K <= MAXT) {
        if (nt == K) team_panels<K, MAXT, S, DIAG>(acc, dg, dg2, bp, ap, inc, npan);
        else team_panels_nt<K + 1, MAXT, S, DIAG>(nt, acc, dg, dg2, bp, ap, inc, npan);
    }
}
// L(row_i, c) = acc[i] inv(L_cc)'  -- the accumulator registers are the A fragments
template <int NT, int MAXT, int S>
__device__ __forceinline__ void team_solve(const double2 (&acc)[MAXT], double2 li, double* xp) {
    double2 x[NT];
#pragma unroll
    for (int i = 0; i < NT; ++i) { x[i] = make_double2(0.0, 0.0); mma884(x[i].x, x[i].y, acc[i].x, li.x); }
#pragma unroll
    for (int i = 0; i < NT; ++i) mma884(x[i].x, x[i].y, acc[i].y, li.y);
#pragma unroll
    for (int i = 0; i < NT; ++i) st2(xp + 64 * S * i, x[i].x, x[i].y);
}
template <int K, int MAXT, int S>
__device__ __forceinline__ void team_solve_nt(int nt, const double2 (&acc)[MAXT], double2 li, double* xp) {
    if constexpr (K <= MAXT) {
        if (nt == K) team_solve<K, MAXT, S>(acc, li, xp);
        else team_solve_nt<K + 1, MAXT, S>(nt, acc, li, xp);
    }
}

// phase timing (debug, tools/phase_timing_pair.py, CCGP_KERNEL=3 CCGP_TEAM_NW=2): team 0 of block 0; warp A slots 0.., warp B_0 slots 16..
#define CCGP_TT(slot) do { if (A.dbg && blockIdx.x == 0 && team == 0 && lane == 0 && role < 2) { \
        long long t1_ = clock64(); A.dbg[role * 16 + (slot)] += t1_ - t_ph; t_ph = t1_; } } while (0)

template <int NW, int MAXT, int DT, int MINB>
__global__ void __launch_bounds__(TEAMS_PER_CTA * NW * 32, MINB) factor_team_kernel(const FactorArgs A) {
    constexpr int NU = NW - 1;
    constexpr int TT = NW * 32;                                           // threads of a team
    constexpr int RAWLD = MAXD + 2;
    extern __shared__ __align__(16) double smem_all[];
    const Layout& lay = A.lay;
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    // warp -> (team, role).  Warp w runs on sub-partition w mod 4 (tools/ubench_smsp.cu):
    //   map 0: team = w mod 4, role = w / 4   -- a team's warps share ONE sub-partition
    //   map 1: role = w / 4, team = (w mod 4 - role) mod 4 -- every sub-partition hosts the chain warp of one
    //          team and update warps of two OTHER teams (their stalls are uncorrelated)
    //   map 2: team = w / NW, role = w mod NW -- (NW = 4) all chain warps on sub-partition 0, update warps on 1..3
    int team, role;
    if (A.team_map == 1) { role = wid / TEAMS_PER_CTA; team = ((wid & 3) - role) & 3; }
    else if (A.team_map == 2) { team = wid / NW; role = wid - team * NW; }
    else { team = wid & (TEAMS_PER_CTA - 1); role = wid / TEAMS_PER_CTA; }
    double* etab = smem_all;
    double* Ls = smem_all + 128 + (size_t)team * (A.team_smem_bytes / 8);
    double* Xs = Ls + lay.total;
    double* ys = Xs + A.d * lay.npx;
    double* linv = ys + lay.npx;                                          // two 8x8 inverses (steps alternate)
    double* usv = linv + 128;                                             // tile (c+1, c), updated but unsolved, for warp A
    double* raw = usv + 64;
    Prm* prm2 = reinterpret_cast<Prm*>(raw + 2 * RAWLD);                  // parameter blocks: current / next
    const int n = lay.n, npad = lay.npad, NJ = lay.NJ, NR = npad >> 3;
    const int bar_pub = 1 + team, bar_step = 1 + TEAMS_PER_CTA + team;    // named barriers of this team
    const int bar_b = 1 + 2 * TEAMS_PER_CTA + team;                       // the B warps among themselves

    for (int e = threadIdx.x; e < 128; e += TEAMS_PER_CTA * TT) etab[e] = CCGP_EXP2_TAB[e];
    const int tl = role * 32 + lane;                                      // thread index within the team
    if (A.design_mode == DESIGN_SHARED) {
        for (int e = tl; e < n * A.d; e += TT) {
            int k = e / n, i = e - k * n;
            Xs[k * lay.npx + i] = A.X[e];
        }
        if (lay.naug) for (int i = tl; i < n; i += TT) ys[i] = A.y[i];
    }
    const int64_t w0 = (int64_t)blockIdx.x * TEAMS_PER_CTA + team, wstride = (int64_t)gridDim.x * TEAMS_PER_CTA;
    const int nprm = A.nparams;
    if (role == 1 && lane < nprm && w0 < A.W) {
        const int64_t pi0 = (A.n_params == 1) ? 0 : w0 / A.n_designs;
        cp_async8(raw + lane, A.cand + pi0 + (int64_t)lane * A.ldc);
    }
    cp_async_wait_all();
    __syncthreads();
    int buf = 0;
    if (role == 1 && lane == 0 && w0 < A.W) load_params_from(A, raw, 1, prm2);
    const double* Ll = Ls + 2 * lane;

    for (int64_t w = w0; w < A.W; w += wstride) {
        const int64_t dsg = w % A.n_designs;
        const Prm* prm = prm2 + buf;
        const int64_t wn = w + wstride;
        if (role == 1 && lane < nprm && wn < A.W) {                       // stage the next parameter row
            const int64_t pin = (A.n_params == 1) ? 0 : wn / A.n_designs;
            cp_async8(raw + (buf ^ 1) * RAWLD + lane, A.cand + pin + (int64_t)lane * A.ldc);
        }
        if (A.design_mode != DESIGN_SHARED) stage_design<TT>(A, dsg, Xs, tl);
        long long t_ph = (A.dbg && blockIdx.x == 0) ? clock64() : 0;
        named_sync(bar_step, TT);                                        // parameters (and the design) visible
        CCGP_TT(0);

        if (prm->clamp) mma_build<DT, true>(A, Ls, Xs, ys, prm, etab, role, NW, lane);
        else mma_build<DT, false>(A, Ls, Xs, ys, prm, etab, role, NW, lane);
        CCGP_TT(1);
        named_sync(bar_step, TT);
        CCGP_TT(2);

        FactorResult res;
        res.mant_all = 1.0; res.mant_tail = 1.0; res.es_all = 0; res.es_tail = 0; res.bad = 0;

        if (role == 0) {
            // ---------------- warp A: the serial chain ----------------
            mma_diag(A, Ls, linv, 0, lane, res);
            for (int c = 0; c < NJ; ++c) {
                __threadfence_block();
                CCGP_TT(3);
                named_sync(bar_pub, TT);                                 // inv(L_cc) out; B's lookahead products in
                CCGP_TT(4);
                if (c + 1 < NJ) {
                    // own copy of L(c+1, c) = (unsolved tile) inv(L_cc)', then tile(c+1,c+1) -= L(c+1,c) L(c+1,c)'
                    // with the solved tile fed back from the accumulator registers: no wait for B's solve
                    const double2 li = ld2(linv + 64 * (c & 1) + 2 * lane);
                    const double2 uv = ld2(usv + 2 * lane);
                    double* blk = Ls + tile_off(c + 1, c + 1, npad);
                    double2 t = ld2(blk + 2 * lane);                     // pre-accumulated through panel c-1 by the last B warp
                    double2 x = make_double2(0.0, 0.0);
                    mma884(x.x, x.y, uv.x, li.x);
                    mma884(x.x, x.y, uv.y, li.y);
                    mma884(t.x, t.y, x.x, negd(x.x));
                    mma884(t.x, t.y, x.y, negd(x.y));
                    st2(blk + 2 * lane, t.x, t.y);
                    __syncwarp();
                    CCGP_TT(5);
                    mma_diag(A, blk, linv + 64 * ((c + 1) & 1), c + 1, lane, res);
                }
            }
            named_sync(bar_step, TT);                                    // B's last solves (the z rows) are stored
        } else {
            // ---------------- warps B_u: everything below the diagonal ----------------
            const int u = role - 1;
            double2 cur[MAXT], nxt[MAXT];
            double2 dg = make_double2(0.0, 0.0), dg2 = make_double2(0.0, 0.0);
            {
                const int own0 = (NR - 2 - u + NU) / NU;                 // own tiles of column 0
#pragma unroll
                for (int i = 0; i < MAXT; ++i) cur[i] = (i < own0) ? ld2(Ll + 64 * (1 + u + NU * i)) : make_double2(0.0, 0.0);
            }
            for (int c = 0; c < NJ; ++c) {
                const int nt = NR - c;                                   // tiles (c+t, c), t < nt; t = 0 is A's
                const int own = (nt - 2 - u + NU) / NU;                  // own: t = 1+u+NU i, i < own
                // last panel into the column
                if (c > 0 && own > 0) {
                    const double* bp = Ll + tile_off(c, c - 1, npad);
                    team_panels_nt<1, MAXT, NU, false>(own, cur, dg, dg2, bp, bp + 64 * (1 + u), 0, 1);
                }
                if (u == 0 && nt > 1) st2(usv + 2 * lane, cur[0].x, cur[0].y);   // tile (c+1, c) for warp A's own solve
                // lookahead: own tiles of column c+1 through panel c-1
                if (c + 1 < NJ) {
                    const int own1 = (nt - 3 - u + NU) / NU;
                    const double* nb = Ll + tile_off(c + 1, c + 1, npad);
#pragma unroll
                    for (int i = 0; i < MAXT; ++i) nxt[i] = (i < own1) ? ld2(nb + 64 * (1 + u + NU * i)) : make_double2(0.0, 0.0);
                    if (c > 0) {
                        const double* bp = Ll + 64 * (c + 1);            // tile (c+1, 0)
                        const int inc = 8 * npad - 64;
                        if (u == NU - 1) {                               // the last B warp (fewest tiles) also pre-accumulates tile (c+1, c+1)
                            dg = make_double2(0.0, 0.0); dg2 = make_double2(0.0, 0.0);
                            team_panels_nt<0, MAXT, NU, true>(own1, nxt, dg, dg2, bp, bp + 64 * (1 + u), inc, c);
                            double* dp = Ls + tile_off(c + 1, c + 1, npad) + 2 * lane;
                            const double2 t0 = ld2(dp);
                            st2(dp, t0.x + (dg.x + dg2.x), t0.y + (dg.y + dg2.y));
                        } else if (own1 > 0) {
                            team_panels_nt<1, MAXT, NU, false>(own1, nxt, dg, dg2, bp, bp + 64 * (1 + u), inc, c);
                        }
                    }
                }
                // meet A: its inverse is out, our lookahead products are in
                __threadfence_block();
                CCGP_TT(3);
                named_sync(bar_pub, TT);
                CCGP_TT(4);
                if (own > 0)
                    team_solve_nt<1, MAXT, NU>(own, cur, ld2(linv + 64 * (c & 1) + 2 * lane), Ls + tile_off(c, c, npad) + 2 * lane + 64 * (1 + u));
#pragma unroll
                for (int i = 0; i < MAXT; ++i) cur[i] = nxt[i];
                __threadfence_block();
                CCGP_TT(5);
                if (NU > 1) named_sync(bar_b, NU * 32);                  // panel c complete (A does not wait for it)
                else __syncwarp();
                CCGP_TT(6);
            }
            named_sync(bar_step, TT);                                    // the z rows are stored: A may reduce
            if (u == 0) {                                                // parameters of the next candidate while A reduces this one
                cp_async_wait_all();
                __syncwarp();
                if (lane == 0 && wn < A.W) load_params_from(A, raw + (buf ^ 1) * RAWLD, 1, prm2 + (buf ^ 1));
            }
        }

        // ---------------- scalars (warp A) ----------------
        if (role == 0) {
            // res is valid in every lane of this warp (mma_diag keeps the bookkeeping redundantly)
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
        }
        __threadfence_block();
        buf ^= 1;
        CCGP_TT(7);
        if (A.dbg && blockIdx.x == 0 && team == 0 && lane == 0 && role == 0) A.dbg[15] += 1;
    }
}

}  // namespace ccgp
