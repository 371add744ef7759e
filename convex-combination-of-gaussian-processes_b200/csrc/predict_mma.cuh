// predict_mma.cuh -- the predictive table on the FP64 tensor path.
//
// Replaces the S*T calls of `predict.post` ([A]:604-623) made by `prediction` ([A]:637-654) and the per-sample
// `factors` ([A]:550-559), like predict_kernel (predict_kernel.cuh), for the Gaussian component families:
//   mean = beta + (z_y - beta z_1).v,  var = sigma2 (1 - v.v + (1 - z_1.v)^2 / z_1.z_1),  v = L^-1 r(x)
// (sigma2 without the (p^2+(1-p)^2) factor, exactly as [A]:619; the correlation vector may use its own
// parameter row -- quirk Q2 of [V]:672.)
//
// Idea: a test site is one more row below the matrix, like y' and 1'.  Its solved row v' = r' L^-T comes out of the
// same left-looking column step as every other row: tile (site group, c) = r-tile - sum_{J<c} V'(sg,J) L(c,J)',
// then times inv(L_cc)' -- all DMMA, operands in the fragment layout of factor_mma.cuh.  So:
//   phase 1 (one CTA of 4 warps per posterior row): build + factor as factor_mma_kernel, but every 8x8 inverse
//            inv(L_cc) is kept (NJ x 64 doubles) next to the factor;
//   phase 2: each warp streams groups of 16 sites through the NJ column steps (V' tiles of the group in a private
//            shared buffer, 13 KB at n = 100), accumulating v.v, z_1.v, z_y.v as the columns are solved.
// Per posterior row at n = 100, T = 625: 125 k exponentials + 14.5 k DMMA, against 2 n^2 T dependent FMAs spread over
// lanes and shuffles in predict_kernel.
#pragma once
#include "factor_mma.cuh"
#include "predict_kernel.cuh"

namespace ccgp {

constexpr int PM_NW = 4;          // warps per CTA (1 diagonal + 3 update warps in phase 1)
constexpr int PM_G = 2;           // site groups (of 8 sites) a warp carries through the column steps at once

// shared doubles: L | Xs | ys | linv_all[NJ*64] | red[64] | etab[128] | vbuf[PM_NW][PM_G*NJ*64] | Prm x2
inline size_t predict_mma_smem_bytes(const Layout& l, int d) {
    size_t dbl = (size_t)l.total + (size_t)d * l.npx + l.npx + (size_t)l.NJ * 64 + 64 + 128 + (size_t)PM_NW * PM_G * l.NJ * 64;
    return (dbl * 8 + 2 * sizeof(Prm) + 15) / 16 * 16;
}

template <int MAXT, int DT>
__global__ void __launch_bounds__(PM_NW * 32, 2) predict_mma_kernel(const PredictArgs P) {
    constexpr int NW = PM_NW, TEAM = NW * 32, NU = NW - 1;
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
    double* vbuf_all = etab + 128;
    Prm* prm = reinterpret_cast<Prm*>(vbuf_all + (size_t)NW * PM_G * NJ * 64);
    Prm* prmv = prm + 1;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int g8 = lane >> 2, m4 = lane & 3;

    for (int e = tid; e < 128; e += TEAM) etab[e] = CCGP_EXP2_TAB[e];
    for (int e = tid; e < n * d; e += TEAM) {
        int k = e / n, i = e - k * n;
        Xs[k * npx + i] = A.X[e];
    }
    for (int i = tid; i < n; i += TEAM) ys[i] = A.y[i];
    const double* Ll = Ls + 2 * lane;

    const int nchunk = (P.fac_mode == 2 && P.t_chunks > 1) ? P.t_chunks : 1;
    const int64_t rec_l = lay.total, rec_i = (int64_t)NJ * 64;
    for (int64_t item = blockIdx.x; item < A.W * nchunk; item += gridDim.x) {
        const int64_t s = item / nchunk;
        const int chunk = (int)(item - s * nchunk);
        __syncthreads();
        if (tid == 0) load_params(A, s, prm);
        if (tid == 32) {
            if (P.candv) {
                load_params_fam(A, P.vec_family, P.candv + s, P.ldcv, prmv);     // (no local copy of the 2 KB argument block)
            } else {
                load_params(A, s, prmv);
            }
        }
        __syncthreads();
        if (P.fac_mode == 2) {
            // the stored factor of row s: L, the inverse diagonal tiles, the bad flag (ccgp_factors_create wrote them)
            const double2* src = reinterpret_cast<const double2*>(P.fac + s * P.fac_ld);
            double2* dl = reinterpret_cast<double2*>(Ls);
            double2* di = reinterpret_cast<double2*>(linv_all);
            for (int64_t e = tid; e < rec_l / 2; e += TEAM) dl[e] = src[e];
            for (int64_t e = tid; e < rec_i / 2; e += TEAM) di[e] = src[rec_l / 2 + e];
            if (tid == 0) red[62] = P.fac[s * P.fac_ld + rec_l + rec_i];
            __syncthreads();
        } else {
        // ---------------- phase 1: build + factor (factor_mma_kernel's schedule), all inverses kept ----------------
        if (prm->clamp) mma_build<DT, true>(A, Ls, Xs, ys, prm, etab, warp, NW, lane);
        else mma_build<DT, false>(A, Ls, Xs, ys, prm, etab, warp, NW, lane);
        __syncthreads();
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
        if (P.fac_mode == 1) {                   // (every step of the factor loop ends in __syncthreads: L and the inverses are final)
            double2* dst = reinterpret_cast<double2*>(P.fac + s * P.fac_ld);
            const double2* sl = reinterpret_cast<const double2*>(Ls);
            const double2* si = reinterpret_cast<const double2*>(linv_all);
            for (int64_t e = tid; e < rec_l / 2; e += TEAM) dst[e] = sl[e];
            for (int64_t e = tid; e < rec_i / 2; e += TEAM) dst[rec_l / 2 + e] = si[e];
            if (tid == 0) { P.fac[s * P.fac_ld + rec_l + rec_i] = red[62]; P.fac[s * P.fac_ld + rec_l + rec_i + 1] = 0.0; }
        }
        }
        // ---------------- scalars: s11 = z_1.z_1, beta = z_1.z_y / s11 ----------------
        double s11 = 0.0, s1y = 0.0;
        for (int k = tid; k < n; k += TEAM) {
            const int off = elem_off_rm(n, k, npad);
            const double zy = Ls[off], z1 = Ls[off + 8];
            s11 = fma(z1, z1, s11);
            s1y = fma(z1, zy, s1y);
        }
        team_sum2<TEAM>(s11, s1y, red);
        const double beta = s1y / s11;
        const bool bad = red[62] != 0.0;
        if (tid == 0 && P.status && chunk == 0) P.status[s] = bad ? 1 : 0;

        // ---------------- phase 2: sites, 16 per warp and pass ----------------
        double* vb = vbuf_all + (size_t)warp * PM_G * NJ * 64;
        const bool clampv = prmv->clamp != 0 || true;      // sites may lie anywhere: always the clamped exponential
        const double rho = prmv->rho, ca = prmv->a, cb = prmv->b;
        // this CTA's share of the sites: groups of 8 * PM_G, chunk `chunk` of `nchunk`
        const int64_t ngrp = (P.T + 8 * PM_G - 1) / (8 * PM_G);
        const int64_t g_lo = ngrp * chunk / nchunk, g_hi = ngrp * (chunk + 1) / nchunk;
        for (int64_t t0 = (g_lo + warp) * (8 * PM_G); t0 < g_hi * (8 * PM_G); t0 += (int64_t)NW * 8 * PM_G) {
            double xs[PM_G][DT > 0 ? DT : MAXD];
#pragma unroll
            for (int g = 0; g < PM_G; ++g) {
                const int64_t t = min(t0 + 8 * g + g8, P.T - 1);
                if (DT > 0) {
#pragma unroll
                    for (int k = 0; k < DT; ++k) xs[g][k] = P.Xnew[t + P.T * k];
                } else {
                    for (int k = 0; k < d; ++k) xs[g][k] = P.Xnew[t + P.T * k];
                }
            }
            double qq[PM_G], u1[PM_G], uy[PM_G];
#pragma unroll
            for (int g = 0; g < PM_G; ++g) { qq[g] = 0.0; u1[g] = 0.0; uy[g] = 0.0; }
            for (int c = 0; c < NJ; ++c) {
                const int j0 = 8 * c + 2 * m4, j1 = j0 + 1;
                const int jc0 = min(j0, n - 1), jc1 = min(j1, n - 1);
                double2 cur[PM_G];
#pragma unroll
                for (int g = 0; g < PM_G; ++g) {
                    double s0 = 0.0, s1 = 0.0;
                    if (DT > 0) {
#pragma unroll
                        for (int k = 0; k < DT; ++k) {
                            const double wk = prmv->wts[k];
                            const double d0 = xs[g][k] - Xs[k * npx + jc0], d1 = xs[g][k] - Xs[k * npx + jc1];
                            s0 = fma(wk * d0, d0, s0);
                            s1 = fma(wk * d1, d1, s1);
                        }
                    } else {
                        for (int k = 0; k < d; ++k) {
                            const double wk = prmv->wts[k];
                            const double d0 = xs[g][k] - Xs[k * npx + jc0], d1 = xs[g][k] - Xs[k * npx + jc1];
                            s0 = fma(wk * d0, d0, s0);
                            s1 = fma(wk * d1, d1, s1);
                        }
                    }
                    double v0 = fma(cb, dexp_neg_tab_dev<true>(rho * s0, etab), ca * dexp_neg_tab_dev<true>(s0, etab));
                    double v1 = fma(cb, dexp_neg_tab_dev<true>(rho * s1, etab), ca * dexp_neg_tab_dev<true>(s1, etab));
                    if (j0 >= n) v0 = 0.0;
                    if (j1 >= n) v1 = 0.0;
                    cur[g] = make_double2(v0, v1);
                }
                // cur[g] -= V'(g, J) L(c, J)' for the panels already solved (operands of panel J+1 are loaded before the
                // DMMAs of panel J issue; the load past the last panel reads valid shared memory and is discarded)
                if (c > 0) {
                    const double* bp = Ll + 64 * c;                    // tile (c, 0)
                    int inc = 8 * npad - 64;
                    double2 b = ld2(bp), a[PM_G], an[PM_G];
#pragma unroll
                    for (int g = 0; g < PM_G; ++g) a[g] = ld2(vb + (g * NJ) * 64 + 2 * lane);
                    for (int J = 0; J < c; ++J) {
                        bp += inc; inc -= 64;
                        const double2 bn = ld2(bp);
#pragma unroll
                        for (int g = 0; g < PM_G; ++g) an[g] = ld2(vb + (g * NJ + J + 1) * 64 + 2 * lane);
                        const double bx = negd(b.x), by = negd(b.y);
#pragma unroll
                        for (int g = 0; g < PM_G; ++g) mma884(cur[g].x, cur[g].y, a[g].x, bx);
#pragma unroll
                        for (int g = 0; g < PM_G; ++g) mma884(cur[g].x, cur[g].y, a[g].y, by);
                        b = bn;
#pragma unroll
                        for (int g = 0; g < PM_G; ++g) a[g] = an[g];
                    }
                }
                const double2 li = ld2(linv_all + 64 * c + 2 * lane);
                const int zoff = elem_off_rm(n, min(j0, npad - 2), npad);
                const double2 zy = ld2(Ls + zoff), z1 = ld2(Ls + zoff + 8);      // z_y, z_1 at columns j0, j0+1
#pragma unroll
                for (int g = 0; g < PM_G; ++g) {
                    double2 x = make_double2(0.0, 0.0);
                    mma884(x.x, x.y, cur[g].x, li.x);
                    mma884(x.x, x.y, cur[g].y, li.y);
                    st2(vb + (g * NJ + c) * 64 + 2 * lane, x.x, x.y);
                    if (j0 < n) { qq[g] = fma(x.x, x.x, qq[g]); u1[g] = fma(z1.x, x.x, u1[g]); uy[g] = fma(fma(-beta, z1.x, zy.x), x.x, uy[g]); }
                    if (j1 < n) { qq[g] = fma(x.y, x.y, qq[g]); u1[g] = fma(z1.y, x.y, u1[g]); uy[g] = fma(fma(-beta, z1.y, zy.y), x.y, uy[g]); }
                }
                __syncwarp();
            }
#pragma unroll
            for (int g = 0; g < PM_G; ++g) {
#pragma unroll
                for (int o = 1; o <= 2; o <<= 1) {
                    qq[g] += __shfl_xor_sync(0xffffffffu, qq[g], o);
                    u1[g] += __shfl_xor_sync(0xffffffffu, u1[g], o);
                    uy[g] += __shfl_xor_sync(0xffffffffu, uy[g], o);
                }
                const int64_t t = t0 + 8 * g + g8;
                if (m4 == 0 && t < P.T) {
                    const double nanv = __longlong_as_double(0x7ff8000000000000LL);
                    const double om = 1.0 - u1[g];
                    P.out_mean[t + P.T * s] = bad ? nanv : beta + uy[g];
                    P.out_var[t + P.T * s] = bad ? nanv : A.sigma2 * (1.0 - qq[g] + om * om / s11);
                }
            }
            __syncwarp();
        }
        (void)clampv;
    }
}

}  // namespace ccgp
