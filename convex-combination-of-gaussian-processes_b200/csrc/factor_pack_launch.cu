// factor_pack_launch.cu -- variant table and launcher of the packed-residency NLL kernel (factor_pack.cuh)
#include "ccgp_ctx.h"
#include "factor_pack.cuh"

// ---- packed-residency kernel (factor_pack.cuh): one warp per candidate, recycled tile slots, up to MAXW warps per CTA ----
struct PackVariant { int maxt, maxw; factor_fn fn_d0, fn_d2; };
#define PV(MT, MW) {MT, MW, factor_pack_kernel<MT, 0, MW>, factor_pack_kernel<MT, 2, MW>},
static const PackVariant g_pack_variants[] = {PV(14, 8)};
#undef PV

int launch_factor_pack(ccgp_ctx* ctx, FactorArgs& A, int* launched, int min_warps) {
    *launched = 0;
    const Layout& l = A.lay;
    const int NR = l.npad / 8;
    if (A.design_mode != DESIGN_SHARED || NR > PACK_MAXNR || NR < 2) return 0;
    A.pack_slots = pack_plan(NR, l.NJ, l.naug ? l.n >> 3 : NR, A.pack_off);
    const size_t wsm = pack_warp_smem_bytes(A.pack_slots), csm = pack_cta_smem_bytes(l, A.d);
    if (csm + wsm > (size_t)ctx->max_smem_optin) return 0;
    int fit = (int)(((size_t)ctx->max_smem_optin - csm) / wsm);
    const int want = env_int("CCGP_PACK_WARPS", 0);
    const PackVariant* var = nullptr;
    int nw = 0;
    for (const PackVariant& v : g_pack_variants) {
        if (NR > v.maxt) continue;
        const int w = std::min(std::min(fit, v.maxw), want > 0 ? want : 64);
        if (w > nw) { nw = w; var = &v; }
    }
    if (!var || nw < 1 || nw < min_warps) return 0;
    if (env_int("CCGP_PACK_EVEN", 1) && nw > 4) nw = nw / 4 * 4;     // the same number of candidates on every sub-partition
    const size_t smem = csm + wsm * nw;
    factor_fn fn = (A.d == 2) ? var->fn_d2 : var->fn_d0;
    CK(cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    int64_t grid = ctx->num_sm;
    if (grid * nw > A.W) grid = (A.W + nw - 1) / nw;
    if (grid < 1) { *launched = 1; return 0; }
    A.team_smem_bytes = (int64_t)wsm;
    A.dbg = ctx->dbg;
    A.debug_stop = 0;
    A.nparams = (A.family == FAM_ANISO) ? A.d + 2 : 3;
    fn<<<(unsigned)grid, nw * 32, smem, ctx->stream>>>(A);
    CK(cudaGetLastError());
    ctx->launches++;
    ctx->last_team = 32; ctx->last_smem = (int)smem; ctx->last_ctas = nw; ctx->last_variant = 500 + var->maxt;
    *launched = 1;
    return 0;
}


