// factor_pc_launch.cu -- launcher of the producer/consumer packed NLL kernel (factor_pc.cuh)
#include "ccgp_ctx.h"
#include "factor_pc.cuh"

int launch_factor_pc(ccgp_ctx* ctx, FactorArgs& A, int* launched) {
    *launched = 0;
    const Layout& l = A.lay;
    const int NR = l.npad / 8;
    if (A.design_mode != DESIGN_SHARED || NR > PACK_MAXNR || NR < 2) return 0;
    A.pack_slots = pc_plan(NR, l.NJ, l.naug ? l.n >> 3 : NR, A.pack_off, A.pack_raw);
    if (A.pack_slots < 0) return 0;
    const size_t wsm = pc_warp_smem_bytes(A.pack_slots), csm = pc_cta_smem_bytes(l, A.d);
    const size_t smem = csm + wsm * PC_NC;
    if (smem > (size_t)ctx->max_smem_optin) return 0;
    factor_fn fn = (A.d == 2) ? factor_pc_kernel<14, 2> : factor_pc_kernel<14, 0>;
    CK(cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    int64_t grid = ctx->num_sm;
    if (grid * PC_NC > A.W) grid = (A.W + PC_NC - 1) / PC_NC;
    if (grid < 1) { *launched = 1; return 0; }
    A.team_smem_bytes = (int64_t)wsm;
    A.dbg = ctx->dbg;
    A.debug_stop = 0;
    A.team_map = std::min(std::max(env_int("CCGP_PC_SPLIT", 1), 0), 2);     // tiles per column the consumer assembles itself
    A.nparams = (A.family == FAM_ANISO) ? A.d + 2 : 3;
    fn<<<(unsigned)grid, (PC_NC + PC_NP) * 32, smem, ctx->stream>>>(A);
    CK(cudaGetLastError());
    ctx->launches++;
    ctx->last_team = 32; ctx->last_smem = (int)smem; ctx->last_ctas = PC_NC; ctx->last_variant = 614;
    *launched = 1;
    return 0;
}
