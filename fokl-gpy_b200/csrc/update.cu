// update.cu -- `fitupdate` (update=True, src/FoKL/FoKLRoutines.py:1850-2583): the draw loops of the three-case sampler
// gibbs_Xin_update on the device.  The math is in update_math.cuh (shared with the host emulation build); the spectral
// preparation of a call (eigendecompositions through fokl_candidates_eval, projections by plain FP64 GEMMs) is driven by
// FoKL/_update.py.  The chain is D strictly sequential draws of O(p) (cases 1, 2) or O(po pn + po^2) (case 3) work: one
// CTA, the coupling matrices read through L1 / L2 (they are a few hundred KB at most), three CTA barriers per draw.
#include "fokl_ctx.cuh"
#include "update_math.cuh"

namespace {

__global__ void update_chain_kernel(fokl::UpdModel m, fokl::UpdArrays A, fokl::UpdVariates V, double *gam_o, double *gam_n,
                                    double *sigs, double *taus, double *lik, int32_t *info)
{
    extern __shared__ __align__(16) double upd_sh[];
    fokl::Team t;
    t.tid = threadIdx.x; t.nthr = blockDim.x;
    t.lane = threadIdx.x & 31; t.nlane = 32;
    t.warp = threadIdx.x >> 5; t.nwarp = blockDim.x >> 5;
    const int bad = fokl::update_chain(t, m, A, V, gam_o, gam_n, sigs, taus, lik, upd_sh);
    if (threadIdx.x == 0 && info) *info = bad;
}

}  // namespace

extern "C" int fokl_update_chain(fokl_ctx *ctx, const fokl_update_model *mdl, const double *lam_o, const double *c_o,
                                 const double *t_o, const double *m_o, const double *lam_n, const double *c_n,
                                 const double *M, const double *Mt, const double *K, const double *W, int rng_mode,
                                 uint64_t seed, uint64_t stream_id, const double *variates, double *gam_o, double *gam_n,
                                 double *sigs, double *taus, double *lik, int32_t *info)
{
    FOKL_CHECK_CTX(ctx);
    if (!mdl || !sigs || !taus || !lik) FOKL_FAIL(ctx, FOKL_EINVAL, "update_chain: null argument");
    const int mode = mdl->mode, po = mdl->po, pn = mdl->pn;
    if (mode < 1 || mode > 3 || po < 0 || pn < 0 || mdl->draws < 1 || mdl->n < 1)
        FOKL_FAIL(ctx, FOKL_EINVAL, "update_chain: bad model");
    if (mode == 1 && (po != 0 || pn < 1 || !lam_n || !c_n || !gam_n))
        FOKL_FAIL(ctx, FOKL_EINVAL, "update_chain: case 1 needs lam_n, c_n, gam_n and po = 0");
    if (mode == 2 && (pn != 0 || po < 1 || !lam_o || !c_o || !m_o || !gam_o))
        FOKL_FAIL(ctx, FOKL_EINVAL, "update_chain: case 2 needs lam_o, c_o, m_o, gam_o and pn = 0");
    if (mode == 3 && (po < 1 || pn < 1 || !lam_o || !c_o || !t_o || !m_o || !lam_n || !c_n || !M || !Mt || !K || !W ||
                      !gam_o || !gam_n))
        FOKL_FAIL(ctx, FOKL_EINVAL, "update_chain: case 3 needs every array");
    if (rng_mode != FOKL_RNG_INJECTED && rng_mode != FOKL_RNG_PHILOX)
        FOKL_FAIL(ctx, FOKL_EINVAL, "update_chain: rng_mode must be injected or philox");
    if (rng_mode == FOKL_RNG_INJECTED && !variates) FOKL_FAIL(ctx, FOKL_EINVAL, "update_chain: injected mode without variates");
    int rc = fokl_bind_device(ctx);
    if (rc) return rc;

    fokl::UpdModel m;
    m.mode = mode; m.po = po; m.pn = pn; m.draws = mdl->draws;
    m.astar = mdl->a_star; m.atau_star = mdl->atau_star;
    m.b = mdl->b; m.btau = mdl->btau; m.sigsqd0 = mdl->sigsqd0; m.yty = mdl->yty; m.squerr = mdl->squerr;
    m.n = (double)mdl->n;
    fokl::UpdArrays A;
    A.lam_o = lam_o; A.c_o = c_o; A.t_o = t_o; A.m_o = m_o; A.lam_n = lam_n; A.c_n = c_n;
    A.M = M; A.Mt = Mt; A.K = K; A.W = W;
    fokl::UpdVariates V;
    V.table = rng_mode == FOKL_RNG_INJECTED ? variates : nullptr;
    V.g.k0 = (uint32_t)seed; V.g.k1 = (uint32_t)(seed >> 32);
    V.s_lo = (uint32_t)stream_id; V.s_hi = (uint32_t)(stream_id >> 32) & 0x7fffffffu;
    V.w = po + pn; V.astar = m.astar; V.atau_star = m.atau_star;

    const int threads = (po + pn) > 96 ? 512 : 256;
    const size_t smem = (size_t)fokl::update_scratch_doubles(po, pn, threads / 32) * sizeof(double);
    if (smem > 48 * 1024) {
        if (smem > (ctx->smem_optin ? ctx->smem_optin : 48 * 1024)) FOKL_FAIL(ctx, FOKL_EINVAL, "update_chain: model too wide");
        FOKL_CUDA(ctx, cudaFuncSetAttribute(update_chain_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    }
    update_chain_kernel<<<1, threads, smem, ctx->stream>>>(m, A, V, gam_o, gam_n, sigs, taus, lik, info);
    FOKL_LAUNCH_CHECK(ctx);
    return FOKL_OK;
}
