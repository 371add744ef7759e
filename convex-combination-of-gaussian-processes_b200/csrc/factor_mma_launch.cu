// factor_mma_launch.cu -- variant table and launcher of the DMMA factor kernel (factor_mma.cuh)
#include "ccgp_ctx.h"
#include "factor_mma.cuh"

// ---- DMMA variants (factor_mma.cuh): NW warps per candidate, MAXT tiles per update warp ----
struct MmaVariant { int nw, maxt; factor_fn fn_d0, fn_d2; };
#define M(NWv, MT) {NWv, MT, factor_mma_kernel<NWv, MT, 0>, factor_mma_kernel<NWv, MT, 2>},
static const MmaVariant g_mma_variants[] = {M(4, 2) M(4, 4) M(4, 6) M(4, 9) M(3, 3) M(3, 6) M(3, 9) M(2, 6) M(2, 12)};
#undef M
static const int g_num_mma_variants = sizeof(g_mma_variants) / sizeof(g_mma_variants[0]);

// returns 1 when the candidate batch was launched on the DMMA kernel, 0 when it does not apply
int launch_factor_mma(ccgp_ctx* ctx, FactorArgs& A, int* launched) {
    *launched = 0;
    const Layout& l = A.lay;
    if (A.family >= FAM_MATERN1D || env_int("CCGP_NO_MMA", 0)) return 0;
    if (l.npad < env_int("CCGP_MMA_MIN_NPAD", 40)) return 0;
    const int nw = env_int("CCGP_MMA_NW", 4);
    const int NR = l.npad / 8;
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

