// predict_kernel.cuh -- predictive mean/variance tables and explicit R^-1.
//
// predict_kernel replaces the S*T calls of `predict.post` ([A]:604-623) made by
// `prediction` ([A]:637-654) plus the per-sample `factors` ([A]:550-559): one CTA
// per posterior sample refactors R once (factor_engine.cuh), then each warp
// takes groups of TP test sites and runs the forward substitution v = L^-1 r(x)
// with the vector distributed over lanes (row i lives in lane i%32):
//   r'R^-1 r            = v.v
//   var.factor1' r      = z_1.v            (1'R^-1 r)
//   mean.factor' r      = (z_y - beta z_1).v
//   var.factor2         = z_1.z_1
//   mean = beta + mean.factor'r;  var = sigma2 (1 - v.v + (1 - z_1.v)^2 / z_1.z_1)
// (sigma2 without the (p^2+(1-p)^2) factor, exactly as [A]:619.)
//
// rinv_kernel produces logpost's `R.Inv` ([A]:448) column by column as
// L^-T (L^-1 e_t); it is meant for the handful of candidates a caller wants the
// explicit inverse for, not for sweeps (n^2*8 bytes per candidate turn the path
// HBM-bound).
#pragma once
#include "factor_engine.cuh"

namespace ccgp {

struct PredictArgs {
    FactorArgs F;          // design, matrix parameters (natural scale), sigma2
    const double* Xnew;    // T x d column-major
    int64_t T;
    const double* candv;   // parameters of the correlation vector (NULL: same as F.cand)
    int64_t ldcv;
    int vec_family;
    int vec_unnormalised;  // 1: r(x) = p^2 r1 + (1-p)^2 r2 without the division ([D2]:479 returns early)
    double* out_mean;      // T x S column-major
    double* out_var;
    int32_t* status;       // S
    // factors kept on the device (ccgp_factors_*, predict_mma_kernel only): per posterior row the factor L (fragment
    // layout, lay.total doubles), the NJ inverse diagonal tiles (NJ * 64) and {bad flag, pad}
    double* fac;           // row s at fac + s * fac_ld
    int64_t fac_ld;
    int fac_mode;          // 0: factor and predict; 1: factor, store, predict (T may be 0); 2: load the stored factor, predict
    int t_chunks;          // mode 2: the sites of one row are split over this many CTAs (>= 1)
};

// extra shared doubles after the factor engine's block:
//   rinvd[npad] | z1[npad] | zr[npad] | rv[warps][TP][npad] | Prm(vec)
template <int TEAM, int TP>
inline size_t predict_smem_bytes(const Layout& l, int d) {
    return smem_bytes(l, d) + (size_t)(3 * l.npad + (TEAM / 32) * TP * l.npad) * 8 + sizeof(Prm) + 16 + MAXD * 8 * (TEAM / 32) * TP;
}

template <int TEAM, int TR, int TC, int DT, int MR, int TP, int MINB>
__global__ void __launch_bounds__(TEAM, MINB) predict_kernel(const PredictArgs P) {
    extern __shared__ __align__(16) double smem[];
    const FactorArgs& A = P.F;
    const Layout& lay = A.lay;
    const SmemPtrs sp = carve_smem(smem, lay, A.d);
    double* Ls = sp.Ls; double* Xs = sp.Xs; double* ys = sp.ys; double* rinv_s = sp.rinv_s; double* red = sp.red;
    Prm* prm = sp.prm;
    stage_tiletab(A.tiletab, sp.tab, lay.NJ, TEAM);
    double* extra = reinterpret_cast<double*>(sp.end);
    double* rinvd = extra;
    double* z1s = rinvd + lay.npad;
    double* zrs = z1s + lay.npad;
    double* rv = zrs + lay.npad;                                   // [warp][TP][npad]
    double* xn = rv + (TEAM / 32) * TP * lay.npad;                 // [warp][TP][MAXD]
    Prm* prmv = reinterpret_cast<Prm*>(xn + (TEAM / 32) * TP * MAXD);
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int n = lay.n, npad = lay.npad, d = A.d, npx = lay.npx;

    for (int e = tid; e < n * d; e += TEAM) {
        int k = e / n, i = e - k * n;
        Xs[k * npx + i] = A.X[e];
    }
    for (int i = tid; i < n; i += TEAM) ys[i] = A.y[i];

    for (int64_t s = blockIdx.x; s < A.W; s += gridDim.x) {
        __syncthreads();
        if (tid == 0) load_params(A, s, prm);
        if (tid == 32 % TEAM) {
            if (P.candv) {
                load_params_fam(A, P.vec_family, P.candv + s, P.ldcv, prmv);     // (no local copy of the 2 KB argument block)
            } else {
                load_params(A, s, prmv);
            }
        }
        __syncthreads();
        FactorResult res = factor_candidate<TEAM, TR, TC, DT>(A, Ls, Xs, ys, rinv_s, prm, sp.ctr, sp.tab, 0);
        if (tid == 0) red[62] = res.bad ? 1.0 : 0.0;

        double s11 = 0.0, s1y = 0.0;
        for (int k = tid; k < n; k += TEAM) {
            int off = elem_off(n, k, npad);
            double zy = Ls[off], z1 = Ls[off + 1];
            s11 = fma(z1, z1, s11);
            s1y = fma(z1, zy, s1y);
        }
        team_sum2<TEAM>(s11, s1y, red);
        const double beta = s1y / s11;
        for (int k = tid; k < n; k += TEAM) {
            int off = elem_off(n, k, npad);
            double zy = Ls[off], z1 = Ls[off + 1];
            z1s[k] = z1;
            zrs[k] = fma(-beta, z1, zy);
            rinvd[k] = 1.0 / Ls[elem_off(k, k, npad)];
        }
        __syncthreads();
        const bool bad = red[62] != 0.0;
        if (tid == 0 && P.status) P.status[s] = bad ? 1 : 0;

            double* myxn = xn + warp * TP * MAXD;
        const int64_t ngroups = (P.T + TP - 1) / TP;
        for (int64_t tg = warp; tg < ngroups; tg += TEAM / 32) {
            const int64_t t0 = tg * TP;
            // stage the TP new sites, then their correlation vectors r(x)
            for (int e = lane; e < TP * d; e += 32) {
                int tp = e / d, k = e - tp * d;
                int64_t t = t0 + tp;
                myxn[tp * MAXD + k] = (t < P.T) ? P.Xnew[t + P.T * k] : 0.0;
            }
            __syncwarp();
            double rr[MR][TP];
#pragma unroll
            for (int m = 0; m < MR; ++m) {
                const int i = lane + 32 * m;
#pragma unroll
                for (int tp = 0; tp < TP; ++tp) {
                    double v = 0.0;
                    if (i < n) {
                        if (prmv->kind != 0) {
                            v = corr1d(prmv, fabs(myxn[tp * MAXD] - Xs[i]));
                            if (P.vec_unnormalised) v *= prmv->w;
                        } else {
                            double s1 = 0.0;
                            for (int k = 0; k < d; ++k) {
                                double df = myxn[tp * MAXD + k] - Xs[k * npx + i];
                                s1 = fma(prmv->wts[k] * df, df, s1);
                            }
                            v = fma(prmv->b, dexp_neg(prmv->rho * s1), prmv->a * dexp_neg(s1));
                        }
                    }
                    rr[m][tp] = v;
                }
            }
            double q[TP], u1[TP], uy[TP];
#pragma unroll
            for (int tp = 0; tp < TP; ++tp) { q[tp] = 0.0; u1[tp] = 0.0; uy[tp] = 0.0; }
#pragma unroll
            for (int m0 = 0; m0 < MR; ++m0) {
                const int kend = min(32, n - 32 * m0);
                for (int kk = 0; kk < kend; ++kk) {
                    const int k = 32 * m0 + kk;
                    const double rk = rinvd[k], z1k = z1s[k], zrk = zrs[k];
                    double vk[TP];
#pragma unroll
                    for (int tp = 0; tp < TP; ++tp) {
                        vk[tp] = __shfl_sync(0xffffffffu, rr[m0][tp], kk) * rk;
                        q[tp] = fma(vk[tp], vk[tp], q[tp]);
                        u1[tp] = fma(z1k, vk[tp], u1[tp]);
                        uy[tp] = fma(zrk, vk[tp], uy[tp]);
                    }
                    const double* colp = Ls + elem_off(0, k, npad);  // (i,k) at colp + i, valid for i >= 8*(k/8)
#pragma unroll
                    for (int m = m0; m < MR; ++m) {
                        const int i = lane + 32 * m;
                        if (i > k && i < n) {
                            const double lik = colp[i];
#pragma unroll
                            for (int tp = 0; tp < TP; ++tp) rr[m][tp] = fma(-lik, vk[tp], rr[m][tp]);
                        }
                    }
                }
            }
            if (lane < TP) {
                double qq = 0, uu1 = 0, uuy = 0;
#pragma unroll
                for (int tp = 0; tp < TP; ++tp)
                    if (lane == tp) { qq = q[tp]; uu1 = u1[tp]; uuy = uy[tp]; }
                const int64_t t = t0 + lane;
                if (t < P.T) {
                    const double nanv = __longlong_as_double(0x7ff8000000000000LL);
                    const double om = 1.0 - uu1;
                    P.out_mean[t + P.T * s] = bad ? nanv : beta + uuy;
                    P.out_var[t + P.T * s] = bad ? nanv : A.sigma2 * (1.0 - qq + om * om / s11);
                }
            }
            __syncwarp();
        }
    }
}

struct RinvArgs {
    FactorArgs F;
    double* out_rinv;  // n*n per candidate, column-major (may be NULL: only beta / rcond wanted)
    double* out_beta;
    double* out_rcond; // 1 / (||R||_1 ||R^-1||_1), both norms exact: what base R's solve() tests against
                       // .Machine$double.eps ([A]:448-449; LAPACK dgecon ESTIMATES the same quantity) (may be NULL)
    int32_t* status;
};

template <int TEAM>
inline size_t rinv_smem_bytes(const Layout& l, int d) {
    return smem_bytes(l, d) + (size_t)(l.npad + (TEAM / 32) * l.npad) * 8 + 32;
}

// One CTA per candidate; warp w produces columns t = w, w+W, ... of R^-1.
template <int TEAM, int TR, int TC, int MINB>
__global__ void __launch_bounds__(TEAM, MINB) rinv_kernel(const RinvArgs P) {
    extern __shared__ __align__(16) double smem[];
    const FactorArgs& A = P.F;
    const Layout& lay = A.lay;
    const SmemPtrs sp = carve_smem(smem, lay, A.d);
    double* Ls = sp.Ls; double* Xs = sp.Xs; double* ys = sp.ys; double* rinv_s = sp.rinv_s; double* red = sp.red;
    Prm* prm = sp.prm;
    stage_tiletab(A.tiletab, sp.tab, lay.NJ, TEAM);
    double* extra = reinterpret_cast<double*>(sp.end);
    double* rinvd = extra;
    double* vbuf = rinvd + lay.npad;  // [warp][npad]
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int n = lay.n, npad = lay.npad, d = A.d, npx = lay.npx;

    for (int e = tid; e < n * d; e += TEAM) {
        int k = e / n, i = e - k * n;
        Xs[k * npx + i] = A.X[e];
    }
    for (int i = tid; i < n; i += TEAM) ys[i] = A.y[i];

    for (int64_t s = blockIdx.x; s < A.W; s += gridDim.x) {
        __syncthreads();
        if (tid == 0) load_params(A, s, prm);
        __syncthreads();
        FactorResult res = factor_candidate<TEAM, TR, TC, 0>(A, Ls, Xs, ys, rinv_s, prm, sp.ctr, sp.tab, 0);
        if (tid == 0) red[62] = res.bad ? 1.0 : 0.0;
        double s11 = 0.0, s1y = 0.0;
        for (int k = tid; k < n; k += TEAM) {
            int off = elem_off(n, k, npad);
            double zy = Ls[off], z1 = Ls[off + 1];
            s11 = fma(z1, z1, s11);
            s1y = fma(z1, zy, s1y);
            rinvd[k] = 1.0 / Ls[elem_off(k, k, npad)];
        }
        team_sum2<TEAM>(s11, s1y, red);
        __syncthreads();
        const bool bad = red[62] != 0.0;
        const double nanv = __longlong_as_double(0x7ff8000000000000LL);
        if (tid == 0) {
            if (P.out_beta) P.out_beta[s] = bad ? nanv : s1y / s11;
            if (P.status) P.status[s] = bad ? 1 : 0;
        }
        double* v = vbuf + warp * npad;
        double* out = P.out_rinv ? P.out_rinv + s * (int64_t)n * n : nullptr;
        double inv1 = 0.0, r1 = 0.0;                 // this warp's largest column 1-norms of R^-1 and of R
        for (int t = warp; t < n; t += TEAM / 32) {
            // forward: v = L^-1 e_t  (zero above t)
            for (int i = lane; i < n; i += 32) v[i] = (i == t) ? 1.0 : 0.0;
            __syncwarp();
            for (int k = t; k < n; ++k) {
                const double vk = v[k] * rinvd[k];
                __syncwarp();
                const double* colp = Ls + elem_off(0, k, npad);
                for (int i = k + 1 + lane; i < n; i += 32) v[i] = fma(-colp[i], vk, v[i]);
                if (lane == 0) v[k] = vk;
                __syncwarp();
            }
            // backward: x = L^-T v
            for (int k = n - 1; k >= 0; --k) {
                const double* colp = Ls + elem_off(0, k, npad);
                double acc = 0.0;
                for (int i = k + 1 + lane; i < n; i += 32) acc = fma(colp[i], v[i], acc);
#pragma unroll
                for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
                const double xk = (v[k] - acc) * rinvd[k];
                __syncwarp();
                if (lane == 0) v[k] = xk;
                __syncwarp();
            }
            if (out) for (int i = lane; i < n; i += 32) out[i + (int64_t)n * t] = bad ? nanv : v[i];
            if (P.out_rcond) {
                double ca = 0.0, cr = 0.0;
                for (int i = lane; i < n; i += 32) {
                    ca += fabs(v[i]);
                    double rv = 1.0;
                    if (i != t) {
                        if (prm->kind != 0) {
                            rv = corr1d(prm, fabs(Xs[i] - Xs[t]));
                        } else {
                            double s1 = 0.0;
                            for (int k = 0; k < d; ++k) { const double df = Xs[k * npx + i] - Xs[k * npx + t]; s1 = fma(prm->wts[k] * df, df, s1); }
                            rv = fma(prm->b, dexp_neg(prm->rho * s1), prm->a * dexp_neg(s1));
                        }
                    }
                    cr += fabs(rv);
                }
#pragma unroll
                for (int o = 16; o > 0; o >>= 1) { ca += __shfl_xor_sync(0xffffffffu, ca, o); cr += __shfl_xor_sync(0xffffffffu, cr, o); }
                inv1 = fmax(inv1, ca); r1 = fmax(r1, cr);
            }
            __syncwarp();
        }
        if (P.out_rcond) {
            __syncthreads();
            if (lane == 0) { red[2 * warp] = inv1; red[2 * warp + 1] = r1; }
            __syncthreads();
            if (tid == 0) {
                double a1 = 0.0, b1 = 0.0;
                for (int w = 0; w < TEAM / 32; ++w) { a1 = fmax(a1, red[2 * w]); b1 = fmax(b1, red[2 * w + 1]); }
                P.out_rcond[s] = bad ? 0.0 : 1.0 / (a1 * b1);
            }
        }
    }
}

}  // namespace ccgp
