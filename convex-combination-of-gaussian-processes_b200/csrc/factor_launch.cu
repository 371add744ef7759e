// factor_launch.cu -- kernel-variant table and launcher of the DFMA factor kernel (factor_engine.cuh)
#include "ccgp_ctx.h"

// ------------------------------------------------------------------ variants
struct Variant { int team, tr, tc, minb, tpc; factor_fn fn_d0, fn_d2; };

// The DFMA kernel now only serves the 1-D Matern/spline families, designs the tensor-path kernels do not cover
// and the CCGP_NO_MMA cross-check: the tuned variants it needs (the round-1 sweep of 21 is in profiles/r01_tune_*).
#define CCGP_VARIANTS(X)                                                                           \
    X(32, 4, 4, 16) X(128, 4, 4, 4) X(256, 4, 4, 2) X(64, 8, 4, 8) Y(4, 4, 4)

#define X(T, R, K, M) {T, R, K, M, 1, factor_kernel<T, R, K, 0, M, 1>, factor_kernel<T, R, K, 2, M, 1>},
#define Y(R, K, P) {32, R, K, 1, P, factor_kernel<32, R, K, 0, 1, P>, factor_kernel<32, R, K, 2, 1, P>},
static const Variant g_variants[] = {CCGP_VARIANTS(X)};
#undef X
#undef Y
static const int g_num_variants = sizeof(g_variants) / sizeof(g_variants[0]);

static int default_variant(const Layout& l) {
    // measured on B200 (profiles/r01_tune_variants_v3.json): one warp per candidate wins while many
    // CTAs fit per SM; 4 warps once shared memory caps residency at ~4 candidates per SM
    if (l.npad <= 72) return 0;    // 32 threads, 4x4 tiles
    if (l.npad <= 136) return 1;   // 128 threads, 4x4 tiles
    return 2;                      // 256 threads, 4x4 tiles (1-2 candidates resident per SM)
}

// choose variant + grid and launch the factor kernel
int launch_factor(ccgp_ctx* ctx, FactorArgs& A) {
    const Layout& l = A.lay;
    {
        int done = 0;
        RC(launch_factor_mma(ctx, A, &done));
        if (done) return 0;
    }
    int v = env_int("CCGP_VARIANT", -1);
    if (v < 0 || v >= g_num_variants) v = default_variant(l);
    const size_t team_smem = (smem_bytes(l, A.d) + 15) / 16 * 16;
    if (team_smem > (size_t)ctx->max_smem_optin) {
        snprintf(ctx->err, sizeof(ctx->err), "n=%d needs %zu B shared memory (> %d)", l.n, team_smem, ctx->max_smem_optin);
        return CCGP_ERR_UNSUPPORTED;
    }
    if (g_variants[v].tpc > 1 && team_smem * g_variants[v].tpc > (size_t)ctx->max_smem_optin) v = default_variant(l);
    const Variant& var = g_variants[v];
    const size_t smem = team_smem * var.tpc;
    A.team_smem_bytes = (int64_t)team_smem;
    factor_fn fn = (A.d == 2) ? var.fn_d2 : var.fn_d0;
    if (env_int("CCGP_NO_DT", 0)) fn = var.fn_d0;
    CK(cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    int nb = 0;
    CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, fn, var.team * var.tpc, smem));
    if (nb < 1) { snprintf(ctx->err, sizeof(ctx->err), "kernel variant %d does not fit", v); return CCGP_ERR_UNSUPPORTED; }
    { int cap = env_int("CCGP_CTAS_PER_SM", 0); if (cap > 0 && cap < nb) nb = cap; }
    int64_t grid = (int64_t)nb * ctx->num_sm;
    if (grid * var.tpc > A.W) grid = (A.W + var.tpc - 1) / var.tpc;
    if (grid < 1) return 0;
    A.dbg = ctx->dbg;
    A.num_sm = env_int("CCGP_ROTATE", 0) ? ctx->num_sm : 0;    // rotation measured no better than warp 0 (profiles/)
    if (env_int("CCGP_FW", -1) >= 0) A.num_sm = -(env_int("CCGP_FW", 0) + 1);
    A.debug_stop = env_int("CCGP_DEBUG_STOP", 0);
    RC(get_tiletab(ctx, l, var.tr, var.tc, &A.tiletab));
    fn<<<(unsigned)grid, var.team * var.tpc, smem, ctx->stream>>>(A);
    CK(cudaGetLastError());
    ctx->launches++;
    ctx->last_team = var.team; ctx->last_smem = (int)smem; ctx->last_ctas = nb; ctx->last_variant = v;
    return 0;
}

