// factor_engine.cuh -- fused covariance-build + Cholesky + forward solves, one
// CTA ("team") per candidate, the whole FP64 factor resident in shared memory.
//
// Replaces, per candidate, the body of the reference's `logpost` up to
// `log.like` ([A]:444-455: Mixed.corr.matrix -> solve(R) -> beta.MLE -> dmnorm)
// and of `cond.like` ([V]:564-575), and -- in determinant mode -- `Entropy`
// ([M]:856-861) / `Augmented.Mixed.Entropy` ([M]:869-877) / subset log-dets.
//
// Algorithm (SURVEY Appendix B "minimal NLL"):
//   A = [ R ; y' ; 1' ]  (n+2 rows, n columns; R by direct differences, unit diag)
//   left-looking blocked Cholesky, panel width 8: the two extra rows come out as
//   z_y = L^-1 y and z_1 = L^-1 1, so the triangular solves are free.
//   beta = z_1.z_y / z_1.z_1,  Q_R = |z_y - beta z_1|^2,  log det R = sum log piv.
//
// Schedule: the mixed correlation matrix is assembled first (all threads, 2 exp per entry);
// then a RIGHT-LOOKING blocked Cholesky with one-panel lookahead, 3 team barriers per panel:
//   [U1] every thread: apply factored panel J to the tiles of block column J+1;
//   [LA] warp 0 factors the 8x8 diagonal block of panel J+1 (lane r <-> row r; one shuffle, one
//        rsqrt, one multiply and one FMA on the per-column critical chain) WHILE the other warps
//        apply panel J to the rest of the trailing matrix, tiles handed out from a shared counter
//        (warp 0 joins when its serial part is done);
//   [TS] one thread per remaining row of panel J+1: triangular solve against the diagonal block.
// Every tile is updated by exactly one thread per step, so results do not depend on which warp
// picked it up (bit-reproducible).
//
// Shared-memory layout of L ("block-column trapezoid"): block column J (8 wide)
// stores rows 8J..npad-1 column-major with height H_J = npad-8J, block columns
// back to back.  Every column start is 16-byte aligned so rows are read as
// double2; a column of L is contiguous in its row index, so the K-loop row
// loads of consecutive lanes are consecutive 16-byte words (conflict-free) and
// the panel-row values are warp-wide broadcasts.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <type_traits>
#include "ccgp_math.h"

namespace ccgp {

constexpr int MAXD = 16;
// A Cholesky pivot of the unit-diagonal R at or below 64 eps means kappa(R) >~ 1e14, where
// no digit of the likelihood survives: the analogue of base R's `solve` refusing
// rcond < .Machine$double.eps ([A]:448-449 -> NA).  (Exactly duplicated design points give
// pivots of a few eps of either sign; R refuses those through dgecon.)
constexpr double PIVOT_MIN = 64 * 2.220446049250313e-16;
constexpr double LOG2PI = 1.8378770664093454835606594728112;
constexpr double LN2 = 0.69314718055994530941723212145818;

enum { FAM_ISO = 0, FAM_ANISO = 1, FAM_ISO_RAW2 = 2, FAM_MATERN1D = 3, FAM_MATERN_SPLINE1D = 4 };
enum { OUT_NLL = 0, OUT_DET = 1 };
enum { DESIGN_SHARED = 0, DESIGN_OLD_PLUS_NEW = 1, DESIGN_GATHER = 2 };

struct Prm {  // per-candidate parameters, one copy in shared memory
    double wts[MAXD];  // component-1 per-dimension scales theta_k
    double rho;        // component-2 exponent = rho * component-1 exponent
    double a, b;       // p^2/w, (1-p)^2/w
    double c;          // w * sigma2
    double p;
    int clamp;         // 1: exponents may exceed 1e8 -> use the argument-clamping exp
    int kind;          // 0 Gaussian components; 1 Matern + Matern; 2 Matern + cubic spline (1-D scripts)
    double c1, c2;     // 1-D kinds: distance scales (2 sqrt(nu)/theta1; 2 sqrt(nu)/theta2 or 1/theta2)
    double mnorm;      // 1 / (Gamma(nu) 2^(nu-1))
    double w;          // p^2 + (1-p)^2
    int twonu;         // 2 nu
    int pad_;
};

// mixed correlation of two 1-D sites at distance h for the Matern / spline kinds
__device__ __forceinline__ double corr1d(const Prm* prm, double h) {
    const double m1 = ccgp_matern(prm->c1 * h, prm->twonu, prm->mnorm);
    const double m2 = (prm->kind == 1) ? ccgp_matern(prm->c2 * h, prm->twonu, prm->mnorm) : ccgp_spline(prm->c2 * h);
    return fma(prm->b, m2, prm->a * m1);
}

struct Layout {
    int n;      // design points
    int naug;   // 0 (determinant mode) or 2 (rows y and 1)
    int npad;   // n + naug rounded up to 8
    int NJ;     // block columns = ceil(n / 8)
    int total;  // doubles of L storage
    int npx;    // leading dimension of the staged design (n rounded up to 2)
};

__host__ __device__ __forceinline__ int blk_base(int J, int npad) { return 8 * (J * npad - 4 * J * (J - 1)); }

inline Layout make_layout(int n, int naug) {
    Layout l;
    l.n = n;
    l.naug = naug;
    l.npad = (n + naug + 7) / 8 * 8;
    l.NJ = (n + 7) / 8;
    l.total = blk_base(l.NJ, l.npad);
    l.npx = (n + 1) / 2 * 2;
    return l;
}

// shared bytes: L | Xs[d*npx] | ys[npx] | rinv[8] | red[64] | Prm | ints[8] | tile table
// (the table has NJ+1 + total/(TR*TC) <= NJ+1 + total/16 entries)
__host__ __device__ __forceinline__ size_t tiletab_bytes(const Layout& l) { return ((size_t)l.NJ + 1 + l.total / 16 + 3) / 4 * 16; }
inline size_t smem_bytes(const Layout& l, int d) {
    size_t dbl = (size_t)l.total + (size_t)d * l.npx + l.npx + 8 + 64;
    return dbl * 8 + sizeof(Prm) + 32 + tiletab_bytes(l);
}

struct SmemPtrs {
    double *Ls, *Xs, *ys, *rinv_s, *red;
    Prm* prm;
    int* ctr;
    uint32_t* tab;
    char* end;     // first byte after the factor engine's block (kernels with extra buffers start here)
};
__device__ __forceinline__ SmemPtrs carve_smem(double* smem, const Layout& lay, int d) {
    SmemPtrs p;
    p.Ls = smem;
    p.Xs = p.Ls + lay.total;
    p.ys = p.Xs + d * lay.npx;
    p.rinv_s = p.ys + lay.npx;
    p.red = p.rinv_s + 8;
    p.prm = reinterpret_cast<Prm*>(p.red + 64);
    p.ctr = reinterpret_cast<int*>(reinterpret_cast<char*>(p.prm) + sizeof(Prm));
    p.tab = reinterpret_cast<uint32_t*>(p.ctr + 8);
    p.end = reinterpret_cast<char*>(p.tab) + tiletab_bytes(lay);
    return p;
}
// copy the tile table (global -> shared) once per CTA; the caller synchronises before use
__device__ __forceinline__ void stage_tiletab(const uint32_t* g, uint32_t* s_tab, int NJ, int nthreads,
                                              int tid = threadIdx.x) {
    const int ntiles = (int)__ldg(g + NJ);
    for (int e = tid; e < NJ + 1 + ntiles; e += nthreads) s_tab[e] = __ldg(g + e);
}

struct FactorArgs {
    Layout lay;
    int d;
    int design_mode;
    const double* X;      // SHARED: n x d (ld n); OLD_PLUS_NEW: D_old n_old x d (ld n_old); GATHER: pool (ld ldpool)
    const double* y;      // SHARED + naug
    const double* Dnew;   // OLD_PLUS_NEW: design c at Dnew + c*n_new*d, n_new x d column-major
    const int32_t* idx;   // GATHER: idx[c + ldi*r]
    int64_t ldi;
    int64_t ldpool;
    int n_old;            // OLD_PLUS_NEW: rows of D_old
    int tail0;            // pivots j >= tail0 form the "tail" determinant
    int64_t n_designs;
    const double* cand;   // n_params x k column-major, ld ldc
    int64_t ldc;
    int64_t n_params;
    int family;
    int logscale;
    double sigma2;
    int mean_mode;
    double tau;
    int64_t W;            // work items: w -> design w % n_designs, parameter row w / n_designs
    int twonu;            // 1-D Matern kinds: 2 nu (integer), and 1/(Gamma(nu) 2^(nu-1))
    double mnorm;
    double span2[MAXD];   // squared coordinate ranges of the design (bounds the exponents)
    int force_clamp;      // 1 when span2 is unknown (per-candidate designs)
    int num_sm;               // CTAs co-resident on one SM are blockIdx = s, s+num_sm, ..: each takes a
                              // different warp for the serial diagonal block (spreads it over the SMSPs)
    const uint32_t* tiletab;  // [NJ+1] first-tile index per block column, then one packed
                              // (first row | row-pair stride << 10 | first column << 20) per tile
    int64_t team_smem_bytes;  // shared bytes of one team (several one-warp teams may share a CTA)
    int debug_stop;       // debug: 1 = stop after the build phase (tools/occupancy_probe.py)
    long long* dbg;       // optional phase-timing buffer (tools/phase_timing.py); NULL in production
    int* sm_slots;        // DMMA kernel: per-SM arrival counters (zeroed before the launch) -> CTA slot on its SM
    int nparams;          // parameters per candidate row (ccgp_num_params)
    int shared_row;       // large-n path: 1 = every candidate uses parameter row 0 (subset log-dets)
    int team_map;         // team kernel: warp -> (team, role) mapping, see factor_team.cuh (CCGP_TEAM_MAP)
    int pack_slots;       // packed-residency kernel (factor_pack.cuh): tile slots per candidate, and the slot table
    uint32_t pack_off[256];   // off[J * 16 + r] = element offset of tile (r, J) in the candidate's slots
    uint32_t pack_raw[256];   // producer/consumer kernel (factor_pc.cuh): element offset of the RAW tile (r, c)
    double* out0;         // NLL: nll          DET: log det (all pivots)
    double* out1;         // NLL: beta         DET: log det (tail pivots)
    double* out2;         // DET: -det(tail) (negated determinant, the ME criterion value)
    int32_t* status;
    int out_mode;
};

// phase timing (debug): block 0, lane 0 of warps 0 and 1 accumulate clock deltas per phase:
// slot = warp*16 + phase*2 + {0: work until the barrier, 1: wait inside the barrier}
#define CCGP_T0() long long t_ph = (A.dbg && blockIdx.x == 0) ? clock64() : 0
#define CCGP_TW(ph) do { if (A.dbg && blockIdx.x == 0 && (threadIdx.x & 31) == 0 && (threadIdx.x >> 5) < 2) { \
        long long t1_ = clock64(); A.dbg[(threadIdx.x >> 5) * 16 + (ph) * 2] += t1_ - t_ph; t_ph = t1_; } } while (0)
#define CCGP_TB(ph) do { if (A.dbg && blockIdx.x == 0 && (threadIdx.x & 31) == 0 && (threadIdx.x >> 5) < 2) { \
        long long t1_ = clock64(); A.dbg[(threadIdx.x >> 5) * 16 + (ph) * 2 + 1] += t1_ - t_ph; t_ph = t1_; } } while (0)

template <int TEAM>
__device__ __forceinline__ void team_sync() {
    if (TEAM == 32) __syncwarp(); else __syncthreads();
}

__device__ __forceinline__ double2 ld2(const double* p) { return *reinterpret_cast<const double2*>(p); }

// running product of pivots kept as mantissa in [1,2) and an integer exponent
__device__ __forceinline__ void prod_accum(double& mant, int& es, double piv) {
    mant *= piv;
    int hi = __double2hiint(mant);
    int e = ((hi >> 20) & 0x7ff) - 1023;
    es += e;
    mant = __hiloint2double(hi - (e << 20), __double2loint(mant));
}

// `c`/`ld`: the candidate's parameter row (element k at c[k*ld]) -- global memory, or a staged copy
// `family`: the component family of THIS row (the predictive kernels read a second row of another family, quirk Q2)
__device__ inline void load_params_fam(const FactorArgs& A, const int family, const double* c, const int64_t ld, Prm* prm) {
    const int d = A.d;
    double p, rho;
    if (family == FAM_ANISO) {
        double lam;
        if (A.logscale) {
            for (int k = 0; k < d; ++k) prm->wts[k] = exp(c[k * ld]);
            p = 1.0 / (1.0 + exp(-c[d * ld]));
            lam = exp(c[(d + 1) * ld]);
        } else {
            p = c[0];
            for (int k = 0; k < d; ++k) prm->wts[k] = c[(k + 1) * ld];
            lam = c[(d + 1) * ld];
        }
        rho = 1.0 + lam;
    } else {
        double t1, t2;
        if (A.logscale) {
            t1 = exp(c[0]);
            t2 = exp(c[ld]);
            p = 1.0 / (1.0 + exp(-c[2 * ld]));
        } else {
            p = c[0];
            t1 = c[ld];
            t2 = c[2 * ld];
        }
        for (int k = 0; k < d; ++k) prm->wts[k] = t1;
        rho = t2 / t1;
    }
    double w = p * p + (1.0 - p) * (1.0 - p);
    prm->kind = 0;
    prm->w = w;
    if (family >= FAM_MATERN1D) {
        const double sn = sqrt(2.0 * A.twonu);                  // 2 sqrt(nu) = sqrt(4 nu) = sqrt(2 * twonu)
        prm->kind = (family == FAM_MATERN1D) ? 1 : 2;
        prm->c1 = sn / prm->wts[0];
        prm->c2 = (prm->kind == 1) ? sn / (rho * prm->wts[0]) : 1.0 / (rho * prm->wts[0]);
        prm->twonu = A.twonu;
        prm->mnorm = A.mnorm;
    }
    prm->rho = rho;
    prm->a = p * p / w;
    prm->b = (1.0 - p) * (1.0 - p) / w;
    prm->c = w * A.sigma2;
    prm->p = p;
    double smax = 0.0;
    for (int k = 0; k < d; ++k) smax += prm->wts[k] * A.span2[k];
    smax *= fmax(rho, 1.0);
    prm->clamp = (A.force_clamp || !(smax < 1e6)) ? 1 : 0;   // 1e6: the table-driven exp's int32 range
}
__device__ inline void load_params_from(const FactorArgs& A, const double* c, const int64_t ld, Prm* prm) {
    load_params_fam(A, A.family, c, ld, prm);
}
__device__ inline void load_params(const FactorArgs& A, int64_t pi, Prm* prm) {
    load_params_from(A, A.cand + pi, A.ldc, prm);
}

struct FactorResult {  // valid in thread 0 of the team after factor_candidate()
    double mant_all, mant_tail;
    int es_all, es_tail;
    int bad;
};

// ---- assemble the entries of panel J (rows 8J..npad-1, 8 columns) -----------------------
// `gt`/`gsz`: index and size (threads, multiple of 32) of the group of threads doing the work.
// Warps take columns, lanes take rows: the row coordinates are per-lane, the column's are
// broadcast.  Every slot of the panel is written (zeros above the diagonal / in dead columns,
// y' and 1' in the two extra rows).  Rows >= n and the 8x8 diagonal corner are fixed up on
// rarely-taken branches so the common entry costs the two exponentials and little else.
template <int DT, bool CLAMP, bool TWOROWS>
__device__ __forceinline__ void build_panel(const FactorArgs& A, double* Ls, const double* Xs, const double* ys,
                                            const Prm* prm, int J, int gt, int gsz) {
    const int n = A.lay.n, npad = A.lay.npad, naug = A.lay.naug, npx = A.lay.npx, d = A.d;
    const int H = npad - 8 * J;
    double* pan = Ls + blk_base(J, npad);
    const int gw = gt >> 5, gnw = gsz >> 5, lane = gt & 31;
    if (prm->kind != 0) {
        // 1-D Matern / spline components ([D1]:368-374, [D2]:454-462): tiny designs, plain loop
        for (int c = gw; c < 8; c += gnw) {
            const int j = 8 * J + c;
            for (int r0 = lane; r0 < H; r0 += 32) {
                const int i = 8 * J + r0;
                double v = 0.0;
                if (j < n) {
                    if (i < n) v = (i == j) ? 1.0 : (i > j ? corr1d(prm, fabs(Xs[i] - Xs[j])) : 0.0);
                    else v = (naug && i == n) ? ys[j] : ((naug && i == n + 1) ? 1.0 : 0.0);
                }
                pan[c * H + r0] = v;
            }
        }
        return;
    }
    const double rho = prm->rho, a = prm->a, b = prm->b;
    double wts[DT > 0 ? DT : 1];
    if (DT > 0) {
#pragma unroll
        for (int k = 0; k < DT; ++k) wts[k] = prm->wts[k];
    }
    // a warp takes two adjacent columns at a time: the row point is loaded once and four
    // exponentials (2 components x 2 columns) are in flight per lane
    for (int c = 2 * gw; c < 8; c += 2 * gnw) {
        const int j = 8 * J + c;
        double* col0 = pan + c * H;
        double* col1 = col0 + H;
        if (j >= n) {                                    // both columns dead (warp-uniform)
            for (int r0 = lane; r0 < H; r0 += 32) { col0[r0] = 0.0; col1[r0] = 0.0; }
            continue;
        }
        const bool live1 = (j + 1) < n;
        const int j1 = live1 ? j + 1 : j;
        double xj0[DT > 0 ? DT : 1], xj1[DT > 0 ? DT : 1];
        if (DT > 0) {
#pragma unroll
            for (int k = 0; k < DT; ++k) { xj0[k] = Xs[k * npx + j]; xj1[k] = Xs[k * npx + j1]; }
        }
        // entries (row r0, columns j and j+1) of the panel
        auto entry2 = [&](int r0, double& v0, double& v1) {
            const int i = 8 * J + r0;
            const int ic = min(i, n - 1);                // rows >= n are fixed up below
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
                    const double d0 = xi - Xs[k * npx + j], d1 = xi - Xs[k * npx + j1];
                    s0 = fma(wk * d0, d0, s0);
                    s1 = fma(wk * d1, d1, s1);
                }
            }
            v0 = fma(b, dexp_neg_dev<CLAMP>(rho * s0), a * dexp_neg_dev<CLAMP>(s0));
            v1 = fma(b, dexp_neg_dev<CLAMP>(rho * s1), a * dexp_neg_dev<CLAMP>(s1));
            if (r0 <= c + 1) {                           // diagonal corner: unit diagonal, unused upper part
                if (r0 <= c) v0 = (r0 == c) ? 1.0 : 0.0;
                v1 = (r0 == c + 1) ? 1.0 : 0.0;
            }
            if (i >= n) {                                // rows y' and 1', zero padding
                v0 = (naug && i == n) ? ys[j] : ((naug && i == n + 1) ? 1.0 : 0.0);
                v1 = (naug && i == n) ? ys[j1] : ((naug && i == n + 1) ? 1.0 : 0.0);
            }
            if (!live1) v1 = 0.0;
        };
        int r0 = lane;
        if (TWOROWS) {
            // a lone warp has nobody to hide its latencies: two full row passes per iteration keep
            // eight exponentials in flight (only complete 32-row passes are paired: nothing is wasted)
            const int Pf = H >> 5;
            for (int u = 0; u + 1 < Pf; u += 2, r0 += 64) {
                double a0, a1, b0, b1;
                entry2(r0, a0, a1);
                entry2(r0 + 32, b0, b1);
                col0[r0] = a0; col1[r0] = a1;
                col0[r0 + 32] = b0; col1[r0 + 32] = b1;
            }
        }
        for (; r0 < H; r0 += 32) {
            double v0, v1;
            entry2(r0, v0, v1);
            col0[r0] = v0;
            col1[r0] = v1;
        }
    }
}

// 1/sqrt(x) for a normal positive x: MUFU.RSQ64H seed (2^-22) + one cubic correction, no
// special-case path (pivots are screened against PIVOT_MIN separately)
__device__ __forceinline__ double fast_rsqrt(double x) {
    double y0;
    asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(y0) : "d"(x));
    const double t = y0 * y0;
    const double e = fma(x, -t, 1.0);
    const double p = fma(e, 0.375, 0.5);
    const double ye = y0 * e;
    return fma(p, ye, y0);
}

// trailing update of one TR x TC tile with the 8 columns of factored panel J:
//   C -= L(rows, 8J..8J+7) L(j0.., 8J..8J+7)'
// The tile's TR rows are TR/2 row PAIRS (ia + m*rs, +1): consecutive tiles of a column group
// take consecutive pairs, so the double2 loads/stores of consecutive lanes are consecutive
// 16-byte words (bank-conflict free); the TC panel-row values are broadcasts.
template <int TR, int TC, int PF>
__device__ __forceinline__ void tile_update(double* Ls, int npad, int J, int ia, int rs, int j0) {
    static_assert(PF == 1 || PF == 2 || PF == 4, "PF");
    const int HJ = npad - 8 * J;
    const double* pJ = Ls + blk_base(J, npad) - 8 * J;             // L(i, 8J+k) at pJ + k*HJ + i
    const int Jc = j0 >> 3, HC = npad - 8 * Jc;
    double* cb = Ls + blk_base(Jc, npad) + (j0 - 8 * Jc) * HC + (ia - 8 * Jc);   // C(ia, j0); next column + HC
    // operands are fetched PF k-steps at a time, one group ahead of the FMAs that use them, so the
    // shared-memory latency of group g+1 hides behind the arithmetic of group g
    constexpr int NG = 8 / PF;
    double2 lr[2][PF][TR / 2], lc[2][PF][TC / 2];
    auto fetch = [&](int g, int buf) {
#pragma unroll
        for (int u = 0; u < PF; ++u) {
            const double* cp = pJ + (g * PF + u) * HJ;
#pragma unroll
            for (int m = 0; m < TR / 2; ++m) lr[buf][u][m] = ld2(cp + ia + m * rs);
#pragma unroll
            for (int m = 0; m < TC / 2; ++m) lc[buf][u][m] = ld2(cp + j0 + 2 * m);
        }
    };
    fetch(0, 0);
    double acc[TR][TC];
#pragma unroll
    for (int cc = 0; cc < TC; ++cc)
#pragma unroll
        for (int m = 0; m < TR / 2; ++m) {
            double2 v = ld2(cb + cc * HC + m * rs);
            acc[2 * m][cc] = v.x; acc[2 * m + 1][cc] = v.y;
        }
#pragma unroll
    for (int g = 0; g < NG; ++g) {
        if (g + 1 < NG) fetch(g + 1, (g + 1) & 1);
#pragma unroll
        for (int u = 0; u < PF; ++u)
#pragma unroll
            for (int m = 0; m < TR / 2; ++m)
#pragma unroll
                for (int mc = 0; mc < TC / 2; ++mc) {
                    const double2 a = lr[g & 1][u][m], b = lc[g & 1][u][mc];
                    acc[2 * m][2 * mc] = fma(-a.x, b.x, acc[2 * m][2 * mc]);
                    acc[2 * m][2 * mc + 1] = fma(-a.x, b.y, acc[2 * m][2 * mc + 1]);
                    acc[2 * m + 1][2 * mc] = fma(-a.y, b.x, acc[2 * m + 1][2 * mc]);
                    acc[2 * m + 1][2 * mc + 1] = fma(-a.y, b.y, acc[2 * m + 1][2 * mc + 1]);
                }
    }
#pragma unroll
    for (int cc = 0; cc < TC; ++cc)
#pragma unroll
        for (int m = 0; m < TR / 2; ++m)
            *reinterpret_cast<double2*>(cb + cc * HC + m * rs) = make_double2(acc[2 * m][cc], acc[2 * m + 1][cc]);
}

// ---- 8x8 diagonal block of panel J, one warp ---------------------------------------------------
// Every lane redundantly factors the whole block in registers (36 entries, fully unrolled): no
// shuffles, ~200 FP64 instructions in the warp's instruction stream instead of ~560 with a
// lane-per-row layout -- under 4-CTA contention the serial part is bound by how many instructions
// the warp must issue, not only by the rsqrt/FMA chain.  Per column the chain is
// rsqrt -> multiply -> FMA (next pivot); everything else fills the issue slots in between.
__device__ __forceinline__ void diag_block(const FactorArgs& A, double* Ls, double* rinv_s, int J, int lane,
                                           FactorResult& res) {
    const int n = A.lay.n, npad = A.lay.npad;
    const int H = npad - 8 * J;
    double* blk = Ls + blk_base(J, npad);          // entry (r, c) at blk[c*H + r]
    double a[8][8];
#pragma unroll
    for (int c = 0; c < 8; ++c) {
#pragma unroll
        for (int r2 = c & ~1; r2 < 8; r2 += 2) {   // rows in aligned pairs; (c-1, c) only adds an unused upper entry
            double2 v = ld2(blk + c * H + r2);
            a[r2][c] = v.x; a[r2 + 1][c] = v.y;
        }
    }
    double pv_own = 1.0;                     // lane 8+c keeps pivot c for the determinant
#pragma unroll
    for (int c = 0; c < 8; ++c) {
        const double piv = a[c][c];
        const bool live = (8 * J + c) < n;
        const double ri = live ? fast_rsqrt(piv) : 0.0;
        if (live && !(piv > PIVOT_MIN)) res.bad = 1;
        if (live && lane == c + 8) pv_own = piv;
        a[c][c] = piv * ri;
#pragma unroll
        for (int r = c + 1; r < 8; ++r) a[r][c] *= ri;
#pragma unroll
        for (int c2 = c + 1; c2 < 8; ++c2)
#pragma unroll
            for (int r = c2; r < 8; ++r) a[r][c2] = fma(-a[r][c], a[c2][c], a[r][c2]);
        if (lane == 0) rinv_s[c] = ri;
    }
    if (lane == 0) {                         // every lane holds the same block; one writes it back
#pragma unroll
        for (int c = 0; c < 8; ++c)
#pragma unroll
            for (int r = c; r < 8; ++r) blk[c * H + r] = a[r][c];
    }
    // determinant bookkeeping off the critical chain: lane 8+c folds pivot c into its own running
    // (mantissa, exponent) pair; the 8 partial products are combined once per candidate
    if (lane >= 8 && lane < 16) {
        prod_accum(res.mant_all, res.es_all, pv_own);
        if (8 * J + (lane - 8) >= A.tail0) prod_accum(res.mant_tail, res.es_tail, pv_own);
    }
}

// ---- rows below the diagonal block of panel J: one thread per row, x L_JJ' = a -----------------
template <int TEAM>
__device__ __forceinline__ void panel_trsm(double* Ls, const double* rinv_s, int npad, int J, int tid) {
    const int H = npad - 8 * J, pbase = blk_base(J, npad);
    const double* dj = Ls + pbase;            // L_JJ(c, c1) at dj[c1*H + c] (warp-wide broadcast loads)
    const int nrows = H - 8;
    // NR rows per thread at a time: their substitution chains are independent, so they interleave
    constexpr int NR = (TEAM == 32) ? 3 : 1;
    for (int t = tid; t < nrows; t += TEAM * NR) {
        double x[NR][8];
        double* p = Ls + pbase + 8 + t;
#pragma unroll
        for (int q = 0; q < NR; ++q) {
            const int tq = min(t + q * TEAM, nrows - 1);     // clamped rows recompute the last row (discarded)
#pragma unroll
            for (int c = 0; c < 8; ++c) x[q][c] = Ls[pbase + 8 + tq + c * H];
        }
#pragma unroll
        for (int c = 0; c < 8; ++c) {
#pragma unroll
            for (int c1 = 0; c1 < c; ++c1) {
                const double l = dj[c1 * H + c];
#pragma unroll
                for (int q = 0; q < NR; ++q) x[q][c] = fma(-x[q][c1], l, x[q][c]);
            }
            const double ri = rinv_s[c];
#pragma unroll
            for (int q = 0; q < NR; ++q) x[q][c] *= ri;
        }
#pragma unroll
        for (int q = 0; q < NR; ++q) {
            if (t + q * TEAM < nrows) {
#pragma unroll
                for (int c = 0; c < 8; ++c) p[q * TEAM + c * H] = x[q][c];
            }
        }
    }
}

// Build A, then right-looking blocked Cholesky with one-panel lookahead.  All threads of the team
// call this.  `ctr` is one shared int (dynamic tile counter).
template <int TEAM, int TR, int TC, int DT>
__device__ __forceinline__ FactorResult factor_candidate(const FactorArgs& A, double* Ls, const double* Xs,
                                                         const double* ys, double* rinv_s, const Prm* prm, int* ctr,
                                                         const uint32_t* tab, int fw, int tid = threadIdx.x) {
    static_assert(TR == 4 || TR == 8, "TR");
    static_assert(TC == 4 || TC == 8, "TC");
    constexpr int W = TEAM / 32;
    const int lane = tid & 31;
    const int warp = tid >> 5;
    const int npad = A.lay.npad, NJ = A.lay.NJ;
    const uint32_t* first = tab;                     // first[Jc]: index of the first tile of block column Jc
    const uint32_t* tiles = tab + NJ + 1;            // (shared-memory copy of the table)
    const int ntiles = (int)first[NJ];

    CCGP_T0();
    if (tid == 0) {
        ctr[2] = 0;                                           // "diagonal block of panel ctr[2] is published"
        if (NJ > 1) { ctr[0] = (int)tab[1] + min((npad - 8) / TR, 4) * (8 / TC); ctr[3] = 0; }
    }
    const bool clampx = prm->clamp != 0;
    for (int J = 0; J < NJ && A.debug_stop != 2; ++J) {
        if (clampx) build_panel<DT, true, (TEAM == 32)>(A, Ls, Xs, ys, prm, J, tid, TEAM);
        else build_panel<DT, false, (TEAM == 32)>(A, Ls, Xs, ys, prm, J, tid, TEAM);
    }
    CCGP_TW(0);
    team_sync<TEAM>();
    CCGP_TB(0);

    FactorResult res;
    res.mant_all = 1.0; res.mant_tail = 1.0; res.es_all = 0; res.es_tail = 0; res.bad = 0;
    if (A.debug_stop >= 1) return res;

    if (warp == fw) diag_block(A, Ls, rinv_s, 0, lane, res);
    team_sync<TEAM>();
    panel_trsm<TEAM>(Ls, rinv_s, npad, 0, tid);
    CCGP_TW(3);
    team_sync<TEAM>();
    CCGP_TB(3);

    // ---- one team barrier per step ----------------------------------------------------------
    // step J applies factored panel J to the trailing matrix.  The serial warp first updates the
    // tiles that hold the diagonal block of panel J+1 (they lead block column J+1 in the table),
    // factors that block and publishes it (flag); meanwhile the other warps take the remaining
    // tiles in chunks from a shared counter (the serial warp joins when done).  Panel J+1's rows
    // are solved as soon as the flag is up and every tile of block column J+1 is finished
    // (completion counter).  Counters alternate between two slots so the next step's can be armed
    // while this step still uses its own.
    constexpr int PFK = (TEAM == 32 ? 4 : 2);
    constexpr int CGN = 8 / TC;
    volatile int* flag = reinterpret_cast<volatile int*>(ctr + 2);
    for (int J = 0; J + 1 < NJ; ++J) {
        const int t1 = (int)first[J + 1], t2 = (int)first[J + 2];
        const int nd = min((npad - 8 * (J + 1)) / TR, 4) * CGN;    // leading tiles = diagonal block of panel J+1
        const int need = t2 - t1 - nd;
        int* cnt = ctr + (J & 1);
        int* dn = ctr + 3 + (J & 1);
        if (W > 1 && tid == 0 && J + 2 < NJ) {                     // arm the next step's slots
            const int ndn = min((npad - 8 * (J + 2)) / TR, 4) * CGN;
            ctr[(J + 1) & 1] = t2 + ndn;
            ctr[3 + ((J + 1) & 1)] = 0;
        }
        if (warp == fw) {
            if (lane < nd) {
                const uint32_t e = tiles[t1 + lane];
                tile_update<TR, TC, PFK>(Ls, npad, J, e & 0x3ff, (e >> 10) & 0x3ff, e >> 20);
            }
            __syncwarp();
            diag_block(A, Ls, rinv_s, J + 1, lane, res);
            __syncwarp();
            if (W > 1 && lane == 0) { __threadfence_block(); *flag = J + 1; }
            CCGP_TW(4);
        }
        if (W == 1) {
            for (int t = t1 + nd + lane; t < ntiles; t += 32) {
                const uint32_t e = tiles[t];
                tile_update<TR, TC, PFK>(Ls, npad, J, e & 0x3ff, (e >> 10) & 0x3ff, e >> 20);
            }
            __syncwarp();
        } else {
            // the NEXT chunk is grabbed before this one is computed: the atomic's latency hides
            int nxt = 0;
            if (lane == 0) nxt = atomicAdd(cnt, 32);
            nxt = __shfl_sync(0xffffffffu, nxt, 0);
            while (nxt < ntiles) {
                const int base = nxt;
                int grab = 0;
                if (lane == 0) grab = atomicAdd(cnt, 32);
                const int t = base + lane;
                if (t < ntiles) {
                    const uint32_t e = tiles[t];
                    tile_update<TR, TC, PFK>(Ls, npad, J, e & 0x3ff, (e >> 10) & 0x3ff, e >> 20);
                }
                const int incol = min(base + 32, t2) - base;       // tiles of block column J+1 in this chunk
                __syncwarp();
                if (incol > 0 && lane == 0) { __threadfence_block(); atomicAdd(dn, incol); }
                nxt = __shfl_sync(0xffffffffu, grab, 0);
            }
            CCGP_TW(2);
            while (*flag < J + 1 || *reinterpret_cast<volatile int*>(dn) < need) __nanosleep(64);
            __threadfence_block();
            __syncwarp();
            CCGP_TB(2);
        }
        // ---- rows of panel J+1 below its diagonal block ---------------------------------------
        panel_trsm<TEAM>(Ls, rinv_s, npad, J + 1, tid);
        CCGP_TW(3);
        team_sync<TEAM>();
        CCGP_TB(3);
    }
    if (warp == fw) {
        // fold the 8 per-column partial products (lanes 8..15) into lane 0, in a fixed order
        res.bad = __any_sync(0xffffffffu, res.bad) ? 1 : 0;
        double ma = 1.0, mt = 1.0;
        int ea = 0, et = 0;
#pragma unroll
        for (int c = 0; c < 8; ++c) {
            const double m1 = __shfl_sync(0xffffffffu, res.mant_all, 8 + c);
            const double m2 = __shfl_sync(0xffffffffu, res.mant_tail, 8 + c);
            const int e1 = __shfl_sync(0xffffffffu, res.es_all, 8 + c);
            const int e2 = __shfl_sync(0xffffffffu, res.es_tail, 8 + c);
            prod_accum(ma, ea, m1); ea += e1;
            prod_accum(mt, et, m2); et += e2;
        }
        res.mant_all = ma; res.es_all = ea; res.mant_tail = mt; res.es_tail = et;
    }
    return res;
}

__device__ __forceinline__ int elem_off(int i, int k, int npad) {
    int J = k >> 3;
    return blk_base(J, npad) + (k & 7) * (npad - 8 * J) + i - 8 * J;
}

// deterministic team-wide sum of two values; result returned to every thread
template <int TEAM>
__device__ __forceinline__ void team_sum2(double& u, double& v, double* red) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        u += __shfl_xor_sync(0xffffffffu, u, o);
        v += __shfl_xor_sync(0xffffffffu, v, o);
    }
    if (TEAM > 32) {
        const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
        __syncthreads();
        if (lane == 0) { red[2 * warp] = u; red[2 * warp + 1] = v; }
        __syncthreads();
        u = 0.0; v = 0.0;
#pragma unroll
        for (int w = 0; w < TEAM / 32; ++w) { u += red[2 * w]; v += red[2 * w + 1]; }
    }
}

template <int TEAM>
__device__ __forceinline__ void stage_design(const FactorArgs& A, int64_t dsg, double* Xs, int tid = threadIdx.x) {
    const int n = A.lay.n, npx = A.lay.npx, d = A.d;
    if (A.design_mode == DESIGN_OLD_PLUS_NEW) {
        const int n_old = A.n_old, n_new = n - n_old;
        const double* Dn = A.Dnew + dsg * (int64_t)(n_new * d);
        for (int e = tid; e < n * d; e += TEAM) {
            int k = e / n, i = e - k * n;
            Xs[k * npx + i] = (i < n_old) ? A.X[k * n_old + i] : Dn[k * n_new + (i - n_old)];
        }
    } else if (A.design_mode == DESIGN_GATHER) {
        for (int e = tid; e < n * d; e += TEAM) {
            int k = e / n, i = e - k * n;
            int64_t src = A.idx[dsg + A.ldi * i];
            Xs[k * npx + i] = A.X[k * A.ldpool + src];
        }
    }
}

// TPC > 1 (TEAM == 32 only): a CTA hosts TPC independent one-warp teams, each with its own slice
// of shared memory and its own candidates.  Hardware spreads the warps of ONE CTA over the four
// SM sub-partitions, whereas the single warps of four separate CTAs were observed to pile up on
// one sub-partition (FP64 pipe active 27 % = one quarter) -- same residency, 4x the issue ports.
template <int TEAM, int TR, int TC, int DT, int MINB, int TPC>
__global__ void __launch_bounds__(TEAM * TPC, MINB) factor_kernel(const FactorArgs A) {
    static_assert(TPC == 1 || TEAM == 32, "several teams per CTA only for one-warp teams");
    extern __shared__ __align__(16) double smem_all[];
    const Layout& lay = A.lay;
    const int team = (TPC == 1) ? 0 : (int)(threadIdx.x >> 5);
    double* smem = smem_all + (size_t)team * (A.team_smem_bytes / 8);
    const SmemPtrs sp = carve_smem(smem, lay, A.d);
    double* Ls = sp.Ls; double* Xs = sp.Xs; double* ys = sp.ys; double* rinv_s = sp.rinv_s; double* red = sp.red;
    Prm* prm = sp.prm;
    int* ctr = sp.ctr;
    const int tid = (TPC == 1) ? (int)threadIdx.x : (int)(threadIdx.x & 31);
    stage_tiletab(A.tiletab, sp.tab, lay.NJ, TEAM, tid);
    const int n = lay.n, npad = lay.npad;
    // warp of the serial part: num_sm > 0 rotates it over co-resident CTAs, num_sm < 0 fixes it to warp (-num_sm - 1)
    const int fw = (A.num_sm > 0 ? (int)(blockIdx.x / A.num_sm) : (A.num_sm < 0 ? -A.num_sm - 1 : 0)) % (TEAM / 32);
    const int64_t w0 = (int64_t)blockIdx.x * TPC + team, wstride = (int64_t)gridDim.x * TPC;
    const int otid = fw * 32;                                                        // its lane 0 writes the outputs

    if (A.design_mode == DESIGN_SHARED) {
        for (int e = tid; e < n * A.d; e += TEAM) {
            int k = e / n, i = e - k * n;
            Xs[k * lay.npx + i] = A.X[e];
        }
        if (lay.naug) for (int i = tid; i < n; i += TEAM) ys[i] = A.y[i];
    }

    for (int64_t w = w0; w < A.W; w += wstride) {
        team_sync<TEAM>();  // previous candidate fully consumed
        const int64_t dsg = w % A.n_designs;
        const int64_t pi = (A.n_params == 1) ? 0 : w / A.n_designs;
        if (tid == 0) load_params(A, pi, prm);
        if (A.design_mode != DESIGN_SHARED) stage_design<TEAM>(A, dsg, Xs, tid);
        team_sync<TEAM>();

        FactorResult res = factor_candidate<TEAM, TR, TC, DT>(A, Ls, Xs, ys, rinv_s, prm, ctr, sp.tab, fw, tid);

        if (A.out_mode == OUT_NLL) {
            double s11 = 0.0, s1y = 0.0;
            for (int k = tid; k < n; k += TEAM) {
                int off = elem_off(n, k, npad);
                double zy = Ls[off], z1 = Ls[off + 1];
                s11 = fma(z1, z1, s11);
                s1y = fma(z1, zy, s1y);
            }
            team_sum2<TEAM>(s11, s1y, red);
            const double beta = s1y / s11;
            double qr = 0.0, dummy = 0.0;
            for (int k = tid; k < n; k += TEAM) {
                int off = elem_off(n, k, npad);
                double rz = fma(-beta, Ls[off + 1], Ls[off]);
                qr = fma(rz, rz, qr);
            }
            team_sum2<TEAM>(qr, dummy, red);
            if (tid == otid) {
                const double c = prm->c;
                const double logdet = log(res.mant_all) + res.es_all * LN2;
                double nll;
                if (A.mean_mode == 0) {
                    nll = 0.5 * (qr / c + n * LOG2PI + n * log(c) + logdet);
                } else {
                    const double g = 1.0 + A.tau * A.tau * s11 / c;
                    const double quad = qr / c + s1y * s1y / (c * s11 * g);
                    nll = 0.5 * (quad + n * LOG2PI + n * log(c) + logdet + log(g));
                }
                const bool bad = res.bad || !(nll == nll);
                const double nanv = __longlong_as_double(0x7ff8000000000000LL);
                A.out0[w] = bad ? nanv : nll;
                if (A.out1) A.out1[w] = bad ? nanv : beta;
                if (A.status) A.status[w] = bad ? 1 : 0;
            }
        } else {
            if (tid == otid) {
                const double nanv = __longlong_as_double(0x7ff8000000000000LL);
                const bool bad = res.bad != 0;
                if (A.out0) A.out0[w] = bad ? nanv : log(res.mant_all) + res.es_all * LN2;
                if (A.out1) A.out1[w] = bad ? nanv : log(res.mant_tail) + res.es_tail * LN2;
                if (A.out2) A.out2[w] = bad ? nanv : -scalbn(res.mant_tail, res.es_tail);
                if (A.status) A.status[w] = bad ? 1 : 0;
            }
        }
    }
}

}  // namespace ccgp
