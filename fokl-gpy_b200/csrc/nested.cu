// nested.cu -- chains of NESTED candidate models without one eigensolver per model.
//
// The kill loop of a substage (FR:1669-1690) accepts deletions one after the other, and the chain of every accepted
// model can matter to a later decision (the threshold of FR:1671 uses the intercept's posterior mean of the most recent
// accepted model, FR:1690).  For the 3-way substages of a 16-input problem that is ~800 models of 700 ... 2000 columns,
// each a principal sub-matrix of its predecessor with ONE row / column removed.  A cold eigendecomposition per model
// (csrc/eigbig.cuh: 6 p^3 flops per Jacobi sweep, 13 - 16 sweeps) costs ~70 ms at p = 2000; here the spectral
// decomposition is carried from one model to the next instead:
//
//   A = Q diag(lam) Q', delete variable m, u = row m of Q (u_j = component m of eigenvector j, |u| = 1).  The eigenvalues
//   mu_1 < ... < mu_{p-1} of the compressed matrix are the roots of the secular equation
//        f(mu) = sum_j u_j^2 / (lam_j - mu) = 0,        lam_i < mu_i < lam_{i+1}   (Cauchy interlacing),
//   its eigenvectors, in the old eigenbasis, z_i ~ (diag(lam) - mu_i)^-1 u, so Q_new = Q Z with row m dropped.
//
// fokl_secular_step does one such step in three kernels:
//   (1) roots: one warp per root; the unknown is the offset delta from the nearer pole, not mu itself (lam_j - mu_i is
//       then accurate to a few ulp however close root and pole are).  Bisection on the IEEE bit pattern of |delta|
//       (26 halvings settle the exponent and 15 bits whatever the magnitude), then safeguarded Newton steps;
//   (2) Gu / Eisenstat weights: uhat_j^2 = prod_i (mu_i - lam_j) / prod_{l != j} (lam_l - lam_j) -- the vector for which
//       the computed roots are the EXACT roots; eigenvectors built from uhat are orthogonal to working precision
//       (built from u they are not when roots crowd a pole);
//   (3) Z' (rows = new eigenvectors in the old eigenbasis), normalised.
// The caller (Engine.nested_chains) applies Q_new = Z' Q with a plain FP64 GEMM and keeps only lam, Q' X'y and the
// intercept's row of every model for fokl_chain_icpt, which runs the eigenbasis draw loop (FR:1519-1548)
// of all models side by side and returns the one statistic the selection loop reads from these chains: the mean of
// the intercept's draws over rows from0.. (FR:1671).  Philox variates are generated in the kernel with the same
// (stream, draw, element) keys as cand_variates_kernel, i.e. these are the chains the cold path would run.
#include "fokl_ctx.cuh"
#include "cand_math.cuh"
#include <math.h>
#include <stdlib.h>
#include <algorithm>
#include <string.h>
#include <vector>

namespace {

constexpr int kSecThreads = 256;
constexpr int kSecWarps = kSecThreads / 32;

__device__ __forceinline__ double warp_sum(double v)
{
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
__device__ __forceinline__ double warp_prod(double v)
{
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v *= __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

// f(anchor + delta) = sum_j u2_j / ((lam_j - lam_a) - delta), summed by one warp (every lane returns the sum)
__device__ __forceinline__ double secular_f(const double *lam, const double *u2, int p, double la, double delta, int lane)
{
    double s = 0.0;
    for (int j = lane; j < p; j += 32) s += u2[j] / ((lam[j] - la) - delta);
    return warp_sum(s);
}

// status bits: 1 = two equal eigenvalues (no interval for a root), 2 = non-finite input
__global__ void __launch_bounds__(kSecThreads) secular_roots_kernel(const double *__restrict__ lam, const double *__restrict__ u,
                                                                    int p, double *__restrict__ u2, double *__restrict__ delta,
                                                                    int32_t *__restrict__ anchor, int32_t *status, int n_bisect)
{
    // u2 is filled by the first pass of every warp's own use below?  No: a separate tiny loop keeps the kernel simple --
    // every CTA squares the whole vector into shared memory.
    extern __shared__ double sh[];
    double *lam_s = sh, *u2_s = sh + p;
    for (int j = threadIdx.x; j < p; j += blockDim.x) {
        const double v = u[j];
        // an exactly vanishing component would put a root ON its pole: floor it (the eigenpair then comes out unchanged)
        const double a = fabs(v) < 1e-140 ? 1e-140 : fabs(v);
        lam_s[j] = lam[j];
        u2_s[j] = a * a;
        if (blockIdx.x == 0) u2[j] = a * a;
        if (!(fabs(v) <= 1.0e300) || !(fabs(lam[j]) <= 1.0e300)) atomicOr(status, 2);
    }
    __syncthreads();
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    for (int i = blockIdx.x * kSecWarps + warp; i < p - 1; i += gridDim.x * kSecWarps) {
        const double lo = lam_s[i], hi = lam_s[i + 1];
        const double g = hi - lo;
        if (!(g > 0.0)) {
            if (lane == 0) { delta[i] = 0.0; anchor[i] = i; atomicOr(status, 1); }
            continue;
        }
        const double half = 0.5 * g;
        const double fm = secular_f(lam_s, u2_s, p, lo, half, lane);
        // root in the lower half: anchor at lam_i, delta in (0, half]; else anchor at lam_{i+1}, delta in [half - g, 0)
        const bool left = fm > 0.0;
        const int a = left ? i : i + 1;
        const double la = lam_s[a];
        // bisection on |delta| in (0, bound]: f is increasing in mu.  left: f(0+) = -inf, f(bound) > 0; right
        // (delta = -t): f(t = 0+) = +inf, f(t = bound) <= 0.
        const double bound = left ? half : g - half;
        // 26 halvings of the bit pattern settle the exponent (<= 11 steps) and >= 15 leading bits of |delta|; safeguarded
        // Newton steps on the bracket then reach the rounding floor of f quadratically (3 - 4 evaluations instead of 36
        // more halvings; a step that leaves the bracket is replaced by its midpoint)
        unsigned long long blo = 0ull, bhi = (unsigned long long)__double_as_longlong(bound);
        for (int it = 0; it < n_bisect && bhi - blo > 1ull; ++it) {
            const unsigned long long bm = blo + ((bhi - blo) >> 1);
            const double t = __longlong_as_double((long long)bm);
            const double f = secular_f(lam_s, u2_s, p, la, left ? t : -t, lane);
            const bool root_below = left ? (f > 0.0) : (f < 0.0);       // the root's |delta| is below t
            if (root_below) bhi = bm; else blo = bm;
        }
        double tlo = __longlong_as_double((long long)blo), thi = __longlong_as_double((long long)bhi);
        double t = thi;
        if (bhi - blo > 1ull) {
            t = 0.5 * (tlo + thi);
            for (int it = 0; it < 12; ++it) {
                const double dl = left ? t : -t;
                double f = 0.0, fp = 0.0;
                for (int j = lane; j < p; j += 32) {
                    const double r = 1.0 / ((lam_s[j] - la) - dl);
                    const double ur = u2_s[j] * r;
                    f += ur;
                    fp = fma(ur, r, fp);
                }
                f = warp_sum(f);
                fp = warp_sum(fp);                                        // d f / d delta > 0
                const bool root_below = left ? (f > 0.0) : (f < 0.0);
                if (root_below) thi = t; else tlo = t;
                double tn = left ? t - f / fp : t + f / fp;               // |delta| = t: d f / d t = +-fp
                if (fabs(tn - t) <= 4.5e-16 * t) { t = tn < tlo ? tlo : (tn > thi ? thi : tn); break; }   // converged
                if (!(tn >= tlo && tn <= thi)) tn = 0.5 * (tlo + thi);
                t = tn;
                if (!(thi > tlo)) break;
            }
        }
        if (lane == 0) {
            delta[i] = left ? t : -t;
            anchor[i] = a;
        }
    }
}

// uhat_j = sign(u_j) sqrt( prod_i (mu_i - lam_j) / (lam_{i'} - lam_j) ),  i' = i (i < j) or i + 1 (i >= j)
__global__ void __launch_bounds__(kSecThreads) secular_uhat_kernel(const double *__restrict__ lam, const double *__restrict__ u,
                                                                   const double *__restrict__ delta,
                                                                   const int32_t *__restrict__ anchor, int p,
                                                                   double *__restrict__ uhat)
{
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    for (int j = blockIdx.x * kSecWarps + warp; j < p; j += gridDim.x * kSecWarps) {
        const double lj = lam[j];
        double pr = 1.0;
        for (int i = lane; i < p - 1; i += 32) {
            const int a = anchor[i];
            const double num = (a == j) ? delta[i] : (lam[a] - lj) + delta[i];
            const double den = lam[i < j ? i : i + 1] - lj;
            pr *= num / den;
        }
        pr = warp_prod(pr);
        if (lane == 0) uhat[j] = copysign(sqrt(fabs(pr)), u[j]);
    }
}

// zt[i][j] = uhat_j / ((lam_j - lam_a(i)) - delta_i), rows normalised; mu_i = lam_a(i) + delta_i
__global__ void __launch_bounds__(kSecThreads) secular_zt_kernel(const double *__restrict__ lam, const double *__restrict__ uhat,
                                                                 const double *__restrict__ delta,
                                                                 const int32_t *__restrict__ anchor, int p,
                                                                 double *__restrict__ mu, double *__restrict__ zt, int64_t ldz)
{
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    for (int i = blockIdx.x * kSecWarps + warp; i < p - 1; i += gridDim.x * kSecWarps) {
        const int a = anchor[i];
        const double la = lam[a], d = delta[i];
        double *row = zt + (int64_t)i * ldz;
        double s = 0.0;
        for (int j = lane; j < p; j += 32) {
            const double v = uhat[j] / ((lam[j] - la) - d);
            row[j] = v;
            s = fma(v, v, s);
        }
        s = warp_sum(s);
        const double inv = rsqrt(s);
        for (int j = lane; j < p; j += 32) row[j] *= inv;
        if (lane == 0) mu[i] = la + d;
    }
}

// ---- intercept-only chains ------------------------------------------------------------------------------------------
struct IcptParams {
    const int32_t *p;            // per model
    const int64_t *off;          // per model: offset into lam / ct / q0
    const uint64_t *stream_id;
    const double *lam, *ct, *q0;
    fokl::CandConst k;
    uint64_t seed;
    double *mean0;               // per model: mean over draws from0 .. D - 1 of beta_0 = sum_j q0_j gamma_j
    int32_t *info;               // per model: 1 = bstar < 0 seen
};

constexpr int kIcptThreads = 256;

// The draw loop of fokl::gibbs_chain (same operations, canonical orientation: every eigenvector is turned so that its
// projection on X'y is non-negative), with the variates generated in place and a fourth sum -- the intercept's draw --
// instead of the stored eigenbasis coefficients.
__global__ void __launch_bounds__(kIcptThreads) chain_icpt_kernel(const IcptParams P)
{
    __shared__ double red[2][4][kIcptThreads / 32];
    __shared__ double gsh[2][2];
    const int c = blockIdx.x;
    const int p = P.p[c];
    const int64_t off = P.off[c];
    const double *lamb = P.lam + off, *ct = P.ct + off, *q0 = P.q0 + off;
    const fokl::CandConst &k = P.k;
    const int D = k.draws;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    constexpr int nwarp = kIcptThreads / 32;
    fokl::Philox g;
    g.k0 = (uint32_t)P.seed;
    g.k1 = (uint32_t)(P.seed >> 32);
    const uint32_t slo = (uint32_t)P.stream_id[c], shi = (uint32_t)(P.stream_id[c] >> 32) & 0x7fffffffu;
    const double astar = fokl::chain_astar(k, p), atau_star = fokl::chain_atau_star(k, p);
    double ssig = sqrt(k.sigsqd0), itau = 1.0 / k.tausqd0;
    int bad = 0;
    double acc0 = 0.0;
    for (int d = 0; d < D; ++d) {
        // the two gamma variates of this draw (rejection loops): one thread each, read after the reduction barrier
        if (tid == kIcptThreads - 1) gsh[d & 1][0] = fokl::philox_gamma(g, slo, shi, (uint32_t)d, 0u, astar);
        if (tid == kIcptThreads - 33) gsh[d & 1][1] = fokl::philox_gamma(g, slo, shi, (uint32_t)d, 1024u, atau_star);
        double s1 = 0.0, s2 = 0.0, s3 = 0.0, s4 = 0.0;
        for (int e = tid; e < p; e += kIcptThreads) {
            double z = fokl::philox_normal(g, slo, shi, (uint32_t)d, (uint32_t)e);
            const double l = lamb[e];
            double cc = ct[e], qq = q0[e];
            if (cc < 0.0) { z = -z; }                         // canon: z_j *= sign(ct_j) -- ct itself keeps its sign
            const double rs = FOKL_RSQRT(l + itau);
            const double gm = (rs * rs) * cc + (ssig * rs) * z;
            s1 += l * gm * gm; s2 += gm * cc; s3 += gm * gm; s4 += qq * gm;
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            s1 += __shfl_xor_sync(0xffffffffu, s1, o);
            s2 += __shfl_xor_sync(0xffffffffu, s2, o);
            s3 += __shfl_xor_sync(0xffffffffu, s3, o);
            s4 += __shfl_xor_sync(0xffffffffu, s4, o);
        }
        if (lane == 0) { red[d & 1][0][warp] = s1; red[d & 1][1][warp] = s2; red[d & 1][2][warp] = s3; red[d & 1][3][warp] = s4; }
        __syncthreads();
        s1 = s2 = s3 = s4 = 0.0;
#pragma unroll
        for (int w = 0; w < nwarp; ++w) { s1 += red[d & 1][0][w]; s2 += red[d & 1][1][w]; s3 += red[d & 1][2][w]; s4 += red[d & 1][3][w]; }
        const double g1 = gsh[d & 1][0], g2 = gsh[d & 1][1];
        const double bstar = k.b + 0.5 * (s1 - 2.0 * s2 + k.yty + s3 * itau);
        double sig, rb;
        if (bstar < 0.0) { sig = nan(""); rb = sig; bad = 1; }
        else { sig = bstar * (1.0 / g1); rb = FOKL_RCP(bstar); }
        const double btau_star = (0.5 * g1 * rb) * s3 + k.btau;
        itau = g2 * FOKL_RCP(btau_star);
        ssig = sqrt(sig);
        if (d >= k.from0) acc0 += s4;
    }
    if (tid == 0) {
        P.mean0[c] = acc0 / (double)(D - k.from0);
        P.info[c] = bad;
    }
}

}  // namespace

extern "C" int fokl_secular_step(fokl_ctx *ctx, const double *lam, const double *u, int p, double *mu, double *zt,
                                 int64_t ldz, double *work, int32_t *status)
{
    FOKL_CHECK_CTX(ctx);
    if (!lam || !u || !mu || !zt || !work || !status || p < 2 || ldz < p)
        FOKL_FAIL(ctx, FOKL_EINVAL, "secular_step: bad argument");
    int rc = fokl_bind_device(ctx);
    if (rc) return rc;
    // work: u2 (p), delta (p), uhat (p), anchor (p ints)
    double *u2 = work, *delta = work + p, *uhat = work + 2 * (size_t)p;
    int32_t *anchor = reinterpret_cast<int32_t *>(work + 3 * (size_t)p);
    const int grid = std::max(1, std::min(4 * ctx->num_sms, (p + kSecWarps - 1) / kSecWarps));
    const size_t smem = 2 * (size_t)p * sizeof(double);
    if (smem > 48 * 1024)
        FOKL_CUDA(ctx, cudaFuncSetAttribute(secular_roots_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    // FOKL_SECULAR_BISECT=64: pure bisection to the last bit (tools / tests)
    int n_bisect = 26;
    if (const char *e = getenv("FOKL_SECULAR_BISECT")) n_bisect = std::max(8, std::min(64, atoi(e)));
    secular_roots_kernel<<<grid, kSecThreads, smem, ctx->stream>>>(lam, u, p, u2, delta, anchor, status, n_bisect);
    FOKL_LAUNCH_CHECK(ctx);
    secular_uhat_kernel<<<grid, kSecThreads, 0, ctx->stream>>>(lam, u, delta, anchor, p, uhat);
    FOKL_LAUNCH_CHECK(ctx);
    secular_zt_kernel<<<grid, kSecThreads, 0, ctx->stream>>>(lam, uhat, delta, anchor, p, mu, zt, ldz);
    FOKL_LAUNCH_CHECK(ctx);
    return FOKL_OK;
}

extern "C" int fokl_chain_icpt(fokl_ctx *ctx, int n_models, const int32_t *p_dev, const int64_t *off_dev,
                               const uint64_t *stream_ids_dev, const double *lam, const double *ct, const double *q0,
                               const fokl_hypers *hyp, uint64_t seed, double *mean0, int32_t *info)
{
    FOKL_CHECK_CTX(ctx);
    if (n_models < 1 || !p_dev || !off_dev || !stream_ids_dev || !lam || !ct || !q0 || !hyp || !mean0 || !info)
        FOKL_FAIL(ctx, FOKL_EINVAL, "chain_icpt: bad argument");
    if (hyp->draws < 1 || hyp->stat_from0 < 0 || hyp->stat_from0 >= hyp->draws)
        FOKL_FAIL(ctx, FOKL_EINVAL, "chain_icpt: statistic window outside the chain");
    int rc = fokl_bind_device(ctx);
    if (rc) return rc;
    IcptParams P;
    P.p = p_dev; P.off = off_dev; P.stream_id = stream_ids_dev; P.lam = lam; P.ct = ct; P.q0 = q0;
    P.k.a = hyp->a; P.k.b = hyp->b; P.k.atau = hyp->atau; P.k.btau = hyp->btau;
    P.k.sigsqd0 = hyp->sigsqd0; P.k.tausqd0 = hyp->tausqd0; P.k.yty = hyp->yty; P.k.sum_y = hyp->sum_y;
    P.k.n = (double)hyp->n; P.k.draws = hyp->draws; P.k.from0 = hyp->stat_from0; P.k.from1 = hyp->stat_from1;
    P.seed = seed; P.mean0 = mean0; P.info = info;
    chain_icpt_kernel<<<n_models, kIcptThreads, 0, ctx->stream>>>(P);
    FOKL_LAUNCH_CHECK(ctx);
    return FOKL_OK;
}
