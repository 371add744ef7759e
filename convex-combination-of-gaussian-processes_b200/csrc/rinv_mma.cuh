// rinv_mma.cuh -- logpost's `R.Inv` ([A]:448 `try(solve(R))`, returned at [A]:466 and stored per posterior sample by
// `factors.frame` [A]:587-591) and the number behind its NA rule, rcond_1(R) = 1 / (||R||_1 ||R^-1||_1) ([A]:448-449:
// solve() fails when LAPACK's estimate of it is below .Machine$double.eps), on the FP64 tensor path.
//
// Same contract as rinv_kernel (predict_kernel.cuh), Gaussian component families.  One CTA of 4 warps per candidate:
//   phase 1  build + factor exactly as predict_mma_kernel (every 8x8 inverse inv(L_cc) kept);
//   phase 2  V = L^-T as 8x8 tiles, from the SAME column step that solves any row below the matrix -- the rows of the
//            identity are test "sites" whose correlation vector is e_i:
//              V(I, I) = inv(L_II)',   V(I, c) = -(sum_{J=I}^{c-1} V(I, J) L(c, J)') inv(L_cc)',  c > I
//            (tile row I per warp, round robin; 2 DMMA per term, two accumulation chains);
//   phase 3  R^-1 = V V': tile (I, J), I >= J, = sum_{K >= I} V(I, K) V(J, K)' -- both operands in the fragment layout of
//            factor_mma.cuh -- written over the factor, which is dead by then;
//   phase 4  the n x n inverse (both triangles) to HBM, column 1-norms of R (summed from the assembled tiles before the
//            factorisation; all entries positive) and of R^-1, fixed summation order -> rcond.
// rinv_kernel does the same with one forward and one backward substitution per COLUMN, n dependent steps each, a warp per
// column: ~8x slower at n = 100, and it sits on the default path of every R-side `logpost` call (na.rule = "rcond").
#pragma once
#include "predict_mma.cuh"

namespace ccgp {

constexpr int RM_NW = 4;

// shared doubles: L (later: R^-1) | Xs | ys | linv_all[NJ*64] | red[64] | etab[128] | V[NJ(NJ+1)/2 * 64] | csum[2*npad] | Prm
inline size_t rinv_mma_smem_bytes(const Layout& l, int d) {
    size_t dbl = (size_t)l.total + (size_t)d * l.npx + l.npx + (size_t)l.NJ * 64 + 64 + 128 +
                 (size_t)l.NJ * (l.NJ + 1) / 2 * 64 + 2 * (size_t)l.npad;
    return (dbl * 8 + sizeof(Prm) + 15) / 16 * 16;
}

template <int MAXT, int DT>
__global__ void __launch_bounds__(RM_NW * 32, 2) rinv_mma_kernel(const RinvArgs P) {
    constexpr int NW = RM_NW, TEAM = NW * 32, NU = NW - 1;
    extern __shared__ __align__(16) double smem_all[];
    const FactorArgs& A = P.F;
    const Layout& lay = A.lay;
    const int n = lay.n, npad = lay.npad, NJ = lay.NJ, NR = npad >> 3, d = A.d, npx = lay.npx;
    double* Ls = smem_all;
    double* Xs = Ls + lay.total;
    double* ys = Xs + d * npx;
    double* linv_all = ys + npx;
    double* red = linv_all + NJ * 64;
    double* etab = red + 64;
    double* Vs = etab + 128;
    double* csumR = Vs + (size_t)NJ * (NJ + 1) / 2 * 64;
    double* csumI = csumR + npad;
    Prm* prm = reinterpret_cast<Prm*>(csumI + npad);
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;

    for (int e = tid; e < 128; e += TEAM) etab[e] = CCGP_EXP2_TAB[e];
    for (int e = tid; e < n * d; e += TEAM) {
        int k = e / n, i = e - k * n;
        Xs[k * npx + i] = A.X[e];
    }
    for (int i = tid; i < n; i += TEAM) ys[i] = A.y[i];
    const double* Ll = Ls + 2 * lane;

    for (int64_t s = blockIdx.x; s < A.W; s += gridDim.x) {
        __syncthreads();
        if (tid == 0) load_params(A, s, prm);
        __syncthreads();
        // ---------------- phase 1: build + factor (predict_mma_kernel's phase 1), all inverses kept ----------------
        if (prm->clamp) mma_build<DT, true>(A, Ls, Xs, ys, prm, etab, warp, NW, lane);
        else mma_build<DT, false>(A, Ls, Xs, ys, prm, etab, warp, NW, lane);
        __syncthreads();
        if (P.out_rcond) {                                  // ||R||_1: column sums of the assembled matrix (unit diagonal)
            for (int t = tid; t < n; t += TEAM) {
                double cs = 1.0;
                for (int i = 0; i < t; ++i) cs += Ls[elem_off_rm(t, i, npad)];
                for (int i = t + 1; i < n; ++i) cs += Ls[elem_off_rm(i, t, npad)];
                csumR[t] = cs;
            }
            __syncthreads();
        }
        FactorResult res;
        res.mant_all = 1.0; res.mant_tail = 1.0; res.es_all = 0; res.es_tail = 0; res.bad = 0;
        if (warp == 0) {
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
                mma_diag(A, blk, linv_all + 64 * c, c, lane, res);
                __threadfence_block();
                named_arrive(1, TEAM);
                __syncthreads();
            }
            if (lane == 0) red[62] = res.bad ? 1.0 : 0.0;
        } else {
            const int uw = warp - 1;
            double2 cur[MAXT], nxt[MAXT];
            double2 dg = make_double2(0.0, 0.0), dg2 = make_double2(0.0, 0.0);
#pragma unroll
            for (int t = 0; t < MAXT; ++t) {
                const int r = 1 + uw + t * NU;
                cur[t] = (r < NR) ? ld2(Ls + tile_off(r, 0, npad) + 2 * lane) : make_double2(0.0, 0.0);
            }
            for (int c = 0; c < NJ; ++c) {
                const int r0 = c + 1 + uw;
                const int nt0 = (NR - r0 + NU - 1) / NU;
                if (c > 0 && nt0 > 0)
                    mma_panels_nt<1, MAXT, NU, false>(nt0, cur, dg, dg2, Ll + tile_off(r0, c - 1, npad),
                                                      Ll + tile_off(c, c - 1, npad), 0, 1);
                if (c + 1 < NJ) {
                    const int r1 = r0 + 1;
                    const int nt1 = max((NR - r1 + NU - 1) / NU, 0);
#pragma unroll
                    for (int t = 0; t < MAXT; ++t)
                        nxt[t] = (t < nt1) ? ld2(Ll + tile_off(r1 + t * NU, c + 1, npad)) : make_double2(0.0, 0.0);
                    if (c > 0) {
                        const double* ap = Ll + 64 * r1;
                        const double* bp = Ll + 64 * (c + 1);
                        const int inc = 8 * npad - 64;
                        if (uw == NU - 1) {
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
                named_sync(1, TEAM);
                if (nt0 > 0) mma_solve_nt<1, MAXT, NU>(nt0, cur, ld2(linv_all + 64 * c + 2 * lane), Ls + tile_off(r0, c, npad) + 2 * lane);
#pragma unroll
                for (int t = 0; t < MAXT; ++t) cur[t] = nxt[t];
                __syncthreads();
            }
        }
        // ---------------- scalars: beta = z_1.z_y / z_1.z_1 ----------------
        double s11 = 0.0, s1y = 0.0;
        for (int k = tid; k < n; k += TEAM) {
            const int off = elem_off_rm(n, k, npad);
            const double zy = Ls[off], z1 = Ls[off + 8];
            s11 = fma(z1, z1, s11);
            s1y = fma(z1, zy, s1y);
        }
        team_sum2<TEAM>(s11, s1y, red);
        const bool bad = red[62] != 0.0;
        const double nanv = __longlong_as_double(0x7ff8000000000000LL);
        if (tid == 0) {
            if (P.out_beta) P.out_beta[s] = bad ? nanv : s1y / s11;
            if (P.status) P.status[s] = bad ? 1 : 0;
        }
        // ---------------- phase 2: V = L^-T, tile row I per warp ----------------
        for (int I = warp; I < NJ; I += NW) {
            {
                const double* li = linv_all + 64 * I;
                const int row = lane >> 2, col = 2 * (lane & 3);
                st2(Vs + ((size_t)I * (I + 1) / 2 + I) * 64 + 2 * lane, li[col * 8 + row], li[(col + 1) * 8 + row]);
            }
            __syncwarp();
            for (int c = I + 1; c < NJ; ++c) {
                double2 acc = make_double2(0.0, 0.0), alt = make_double2(0.0, 0.0);
                for (int J = I; J < c; ++J) {
                    const double2 a = ld2(Vs + ((size_t)J * (J + 1) / 2 + I) * 64 + 2 * lane);
                    const double2 b = ld2(Ll + tile_off(c, J, npad));
                    mma884(acc.x, acc.y, a.x, negd(b.x));
                    mma884(alt.x, alt.y, a.y, negd(b.y));
                }
                acc.x += alt.x; acc.y += alt.y;
                const double2 li = ld2(linv_all + 64 * c + 2 * lane);
                double2 x = make_double2(0.0, 0.0);
                mma884(x.x, x.y, acc.x, li.x);
                mma884(x.x, x.y, acc.y, li.y);
                st2(Vs + ((size_t)c * (c + 1) / 2 + I) * 64 + 2 * lane, x.x, x.y);
                __syncwarp();
            }
        }
        __syncthreads();
        // ---------------- phase 3: R^-1 = V V', lower tiles, over the (dead) factor ----------------
        {
            int I = 0, J = warp;                            // tile index t = I (I + 1) / 2 + J, t = warp, warp + NW, ...
            while (J > I) { J -= I + 1; ++I; }
            while (I < NJ) {
                double2 acc = make_double2(0.0, 0.0), alt = make_double2(0.0, 0.0);
                for (int K = I; K < NJ; ++K) {
                    const double2 a = ld2(Vs + ((size_t)K * (K + 1) / 2 + I) * 64 + 2 * lane);
                    const double2 b = ld2(Vs + ((size_t)K * (K + 1) / 2 + J) * 64 + 2 * lane);
                    mma884(acc.x, acc.y, a.x, b.x);
                    mma884(alt.x, alt.y, a.y, b.y);
                }
                st2(Ls + tile_off(I, J, npad) + 2 * lane, acc.x + alt.x, acc.y + alt.y);
                J += NW;
                while (J > I) { J -= I + 1; ++I; }
            }
        }
        __syncthreads();
        // ---------------- phase 4: outputs ----------------
        if (P.out_rinv) {
            double* out = P.out_rinv + s * (int64_t)n * n;
            for (int e = tid; e < n * n; e += TEAM) {
                const int t = e / n, i = e - t * n;
                const double v = Ls[elem_off_rm(max(i, t), min(i, t), npad)];
                out[e] = bad ? nanv : v;
            }
        }
        if (P.out_rcond) {
            for (int t = tid; t < n; t += TEAM) {
                double cs = 0.0;
                for (int i = 0; i < t; ++i) cs += fabs(Ls[elem_off_rm(t, i, npad)]);
                for (int i = t; i < n; ++i) cs += fabs(Ls[elem_off_rm(i, t, npad)]);
                csumI[t] = cs;
            }
            __syncthreads();
            if (tid == 0) {
                double a1 = 0.0, b1 = 0.0;
                for (int t = 0; t < n; ++t) { a1 = fmax(a1, csumI[t]); b1 = fmax(b1, csumR[t]); }
                P.out_rcond[s] = bad ? 0.0 : 1.0 / (a1 * b1);
            }
        }
    }
}

}  // namespace ccgp
