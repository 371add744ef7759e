// factor_mma_launch.cu -- variant table and launcher of the DMMA factor kernel (factor_mma.cuh)
#include "ccgp_ctx.h"
#include "factor_mma.cuh"
#include "factor_warp.cuh"
#include "factor_team.cuh"

// ---- DMMA variants (factor_mma.cuh): NW warps per candidate, MAXT tiles per update warp ----
struct MmaVariant { int nw, maxt; factor_fn fn_d0, fn_d2; };
#define M(NWv, MT) {NWv, MT, factor_mma_kernel<NWv, MT, 0>, factor_mma_kernel<NWv, MT, 2>},
static const MmaVariant g_mma_variants[] = {M(4, 6) M(4, 9) M(2, 12)};
#undef M
static const int g_num_mma_variants = sizeof(g_mma_variants) / sizeof(g_mma_variants[0]);

// ---- one-warp-per-candidate DMMA kernel (factor_warp.cuh): MAXT = tiles of the first block column, even ----
struct WarpVariant { int maxt, minb; factor_fn fn_d0, fn_d2; };
#define WV(MT, MB) {MT, MB, factor_warp_kernel<MT, 0, MB>, factor_warp_kernel<MT, 2, MB>},
static const WarpVariant g_warp_variants[] = {WV(4, 3) WV(8, 2) WV(10, 2) WV(14, 1)};
#undef WV

// ---- team kernel (factor_team.cuh): NW warps per candidate on one sub-partition; nrmax = most tile rows ----
struct TeamVariant { int nw, nrmax, maxt; factor_fn fn_d0, fn_d2; };
#define TV(NWv, NRM, MT, MB) {NWv, NRM, MT, factor_team_kernel<NWv, MT, 0, MB>, factor_team_kernel<NWv, MT, 2, MB>},
static const TeamVariant g_team_variants[] = {TV(2, 14, 13, 1) TV(3, 14, 7, 1) TV(4, 14, 5, 1)};
#undef TV

static int launch_factor_team(ccgp_ctx* ctx, FactorArgs& A, int* launched) {
    *launched = 0;
    const Layout& l = A.lay;
    const int NR = l.npad / 8;
    const int nw = env_int("CCGP_TEAM_NW", 3);
    const TeamVariant* var = nullptr;
    for (const TeamVariant& v : g_team_variants)
        if (v.nw == nw && NR <= v.nrmax) { var = &v; break; }
    if (!var) return 0;
    const size_t tsm = team_smem_bytes(l, A.d);
    const size_t smem = tsm * TEAMS_PER_CTA + TEAM_CTA_EXTRA;
    if (smem > (size_t)ctx->max_smem_optin) return 0;
    const int threads = TEAMS_PER_CTA * var->nw * 32;
    factor_fn fn = (A.d == 2) ? var->fn_d2 : var->fn_d0;
    CK(cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    int nb = 0;
    CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, fn, threads, smem));
    if (nb < 1) return 0;
    { int cap = env_int("CCGP_CTAS_PER_SM", 0); if (cap > 0 && cap < nb) nb = cap; }
    int64_t grid = (int64_t)nb * ctx->num_sm;
    if (grid * TEAMS_PER_CTA > A.W) grid = (A.W + TEAMS_PER_CTA - 1) / TEAMS_PER_CTA;
    if (grid < 1) { *launched = 1; return 0; }
    A.team_smem_bytes = (int64_t)tsm;
    A.dbg = ctx->dbg;
    A.nparams = (A.family == FAM_ANISO) ? A.d + 2 : 3;
    A.team_map = env_int("CCGP_TEAM_MAP", 0);
    if (A.team_map == 2 && var->nw != 4) A.team_map = 0;
    fn<<<(unsigned)grid, threads, smem, ctx->stream>>>(A);
    CK(cudaGetLastError());
    ctx->launches++;
    ctx->last_team = var->nw * 32; ctx->last_smem = (int)smem; ctx->last_ctas = nb; ctx->last_variant = 400 + var->nw * 10;
    *launched = 1;
    return 0;
}

static int launch_factor_warp(ccgp_ctx* ctx, FactorArgs& A, int* launched) {
    *launched = 0;
    const Layout& l = A.lay;
    const int NR = l.npad / 8;
    const WarpVariant* var = nullptr;
    for (const WarpVariant& v : g_warp_variants)
        if (NR <= v.maxt) { var = &v; break; }
    if (!var) return 0;
    const size_t team_smem = warp_team_smem_bytes(l, A.d);
    const size_t smem = team_smem * WARP_TEAMS + WARP_CTA_EXTRA;
    if (smem > (size_t)ctx->max_smem_optin) return 0;
    factor_fn fn = (A.d == 2) ? var->fn_d2 : var->fn_d0;
    CK(cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    int nb = 0;
    CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, fn, WARP_TEAMS * 32, smem));
    if (nb < 1) return 0;
    { int cap = env_int("CCGP_CTAS_PER_SM", 0); if (cap > 0 && cap < nb) nb = cap; }
    int64_t grid = (int64_t)nb * ctx->num_sm;
    if (grid * WARP_TEAMS > A.W) grid = (A.W + WARP_TEAMS - 1) / WARP_TEAMS;
    if (grid < 1) { *launched = 1; return 0; }
    A.team_smem_bytes = (int64_t)team_smem;
    A.dbg = ctx->dbg;
    A.nparams = (A.family == FAM_ANISO) ? A.d + 2 : 3;
    fn<<<(unsigned)grid, WARP_TEAMS * 32, smem, ctx->stream>>>(A);
    CK(cudaGetLastError());
    ctx->launches++;
    ctx->last_team = 32; ctx->last_smem = (int)smem; ctx->last_ctas = nb; ctx->last_variant = 200 + var->maxt;
    *launched = 1;
    return 0;
}

// returns 1 when the candidate batch was launched on a DMMA kernel, 0 when none applies.
// Choice by tile rows NR = npad/8, measured on B200 (profiles/r01_tune_kernels.txt):
//   NR <= 10 (n <= ~78): one warp per candidate -- shared memory lets 2-3 candidates share a sub-partition
//   NR 11..13 and d = 2: the packed-residency kernel (factor_pack.cuh), eight one-warp candidates per SM
//   NR <= 14 (n <= ~110): three warps per candidate on one sub-partition (shared memory caps residency at 4/SM)
//   larger: the CTA-per-candidate DMMA kernel (factor_mma.cuh), then the DFMA kernel / HBM path
int launch_factor_mma(ccgp_ctx* ctx, FactorArgs& A, int* launched) {
    *launched = 0;
    const Layout& l = A.lay;
    if (A.family >= FAM_MATERN1D || env_int("CCGP_NO_MMA", 0)) return 0;
    const int NR = l.npad / 8;
    const int kern = env_int("CCGP_KERNEL", 0);      // 0 auto, 1 warp, 3 team, 4 cta, 5 packed (factor_pack.cuh), 6 producer/consumer packed (factor_pc.cuh)
    if (kern == 5) {
        RC(launch_factor_pack(ctx, A, launched));
        if (*launched) return 0;
    }
    if (kern == 6) {
        RC(launch_factor_pc(ctx, A, launched));
        if (*launched) return 0;
    }
    if ((kern == 0 && NR <= 10) || kern == 1) {
        RC(launch_factor_warp(ctx, A, launched));
        if (*launched) return 0;
    }
    // 2-D designs of 11..13 tile rows (n = 79..102): eight packed candidates per SM fit and beat four teams by 4-6 %
    // (profiles/r02_pack_vs_team.txt: n = 84/88/96/100 -> +6/+5/+4/+4 %; with d >= 3 the packed kernel's generic
    // distance loop loses 12 %, and from n = 103 only four candidates fit)
    if (kern == 0 && A.d == 2 && NR >= 11 && NR <= 13 && !env_int("CCGP_NO_PACK", 0)) {
        RC(launch_factor_pack(ctx, A, launched, 8));
        if (*launched) return 0;
    }
    if ((kern == 0 && NR <= 14) || kern == 3) {
        RC(launch_factor_team(ctx, A, launched));
        if (*launched) return 0;
        if (kern == 0) {                               // four teams no longer fit shared memory (n ~ 105..110): one warp each does
            RC(launch_factor_warp(ctx, A, launched));
            if (*launched) return 0;
        }
    }
    if (l.npad < env_int("CCGP_MMA_MIN_NPAD", 40)) return 0;
    const int nw = env_int("CCGP_MMA_NW", 4);
    const MmaVariant* var = nullptr;
    for (int i = 0; i < g_num_mma_variants; ++i) {
        const MmaVariant& v = g_mma_variants[i];
        if (v.nw == nw && (NR - 1 + (v.nw - 2)) / (v.nw - 1) <= v.maxt) { var = &v; break; }
    }
    if (!var) return 0;
    const size_t smem = mma_smem_bytes(l, A.d);
    if (smem > (size_t)ctx->max_smem_optin) return 0;
    factor_fn fn = (A.d == 2) ? var->fn_d2 : var->fn_d0;
    CK(cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    int nb = 0;
    CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, fn, var->nw * 32, smem));
    if (nb < 1) return 0;
    { int cap = env_int("CCGP_CTAS_PER_SM", 0); if (cap > 0 && cap < nb) nb = cap; }
    int64_t grid = (int64_t)nb * ctx->num_sm;
    if (grid > A.W) grid = A.W;
    if (grid < 1) { *launched = 1; return 0; }
    A.team_smem_bytes = (int64_t)smem;
    A.dbg = ctx->dbg;
    A.nparams = (A.family == FAM_ANISO) ? A.d + 2 : 3;
    if (!ctx->sm_slots) CK(cudaMalloc(&ctx->sm_slots, 1024 * sizeof(int)));
    CK(cudaMemsetAsync(ctx->sm_slots, 0, 1024 * sizeof(int), ctx->stream));
    A.sm_slots = env_int("CCGP_MMA_NOROTATE", 0) ? nullptr : ctx->sm_slots;
    fn<<<(unsigned)grid, var->nw * 32, smem, ctx->stream>>>(A);
    CK(cudaGetLastError());
    ctx->launches++;
    ctx->last_team = var->nw * 32; ctx->last_smem = (int)smem; ctx->last_ctas = nb; ctx->last_variant = 100 + var->nw * 10;
    *launched = 1;
    return 0;
}

