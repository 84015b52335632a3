// gram.cu -- K2: Gram / projection update on the FP64 tensor pipe (sm_100a, mma.sync DMMA).
//
// Replaces XtX = X'X and Xty = X'y of the reference (src/FoKL/FoKLRoutines.py:1492-1494), which
// recomputes the full P x P product on every `gibbs` call.  Here only the block that is new when C
// columns are appended is formed:   block = [X_old  X_new  y]' X_new     ((P_old + C + 1) x C)
//
// X is column-major, so both MMA operands are contiguous along the reduction (row) index: a CTA streams
// KB-row slabs of 64 "A" columns and 64 "B" columns into shared memory with cp.async (3-stage ring),
// and 8 warps issue mma.sync.m8n8k4.f64 on a 64 x 64 output tile.  The row range is split across
// CTAs (blockIdx.y); partial tiles go to a workspace and are summed in a fixed order (deterministic).
// blockIdx.x (the tile) varies fastest so CTAs of one row slab run together and share it through L2.
#include "fokl_ctx.cuh"
#include <algorithm>

namespace {

constexpr int TM = 64, TN = 64, KB = 32, STRIDE = KB + 4, STAGES = 3;
constexpr int kGramThreads = 256;
constexpr size_t kStageDoubles = (size_t)(TM + TN) * STRIDE;
constexpr size_t kGramSmem = STAGES * kStageDoubles * sizeof(double);

struct GramParams {
    const double *X;
    const double *y;
    int64_t ld, n;
    int p_old, c;          // A side: columns 0 .. p_old + c - 1 of X, then y; B side: columns p_old .. p_old + c - 1
    int tiles_b;
    int64_t rows_per_split;
    double *out;           // nsplit x (p + 1) x c partials (or the final block when nsplit == 1)
};

__device__ __forceinline__ void cp_async16(void *smem, const void *gmem, int src_bytes)
{
    unsigned s = (unsigned)__cvta_generic_to_shared(smem);
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;\n" ::"r"(s), "l"(gmem), "r"(src_bytes));
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;\n" ::); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;\n" ::"n"(N)); }

__device__ __forceinline__ void dmma_m8n8k4(double &c0, double &c1, double a, double b)
{
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
                 : "+d"(c0), "+d"(c1)
                 : "d"(a), "d"(b));
}

__global__ void __launch_bounds__(kGramThreads, 2) gram_kernel(const GramParams P)
{
    extern __shared__ __align__(16) double smem[];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int wm = warp >> 2, wn = warp & 3;
    const int p = P.p_old + P.c;
    const int ta = blockIdx.x / P.tiles_b, tb = blockIdx.x % P.tiles_b;
    const int64_t n_lo = (int64_t)blockIdx.y * P.rows_per_split;
    const int64_t n_hi = (n_lo + P.rows_per_split < P.n) ? n_lo + P.rows_per_split : P.n;
    const int nk = (int)((n_hi - n_lo + KB - 1) / KB);

    // per-thread copy assignments: 8 x 16-byte segments per stage
    const double *src_col[8];
    int dst_off[8], seg_row[8];
#pragma unroll
    for (int q = 0; q < 8; ++q) {
        int idx = q * kGramThreads + tid;
        int col = idx >> 4, seg = idx & 15;          // col 0..127 (A then B), seg 0..15
        const double *base = nullptr;
        if (col < TM) {
            int i = ta * TM + col;
            if (i < p) base = P.X + (int64_t)i * P.ld;
            else if (i == p) base = P.y;
        } else {
            int j = tb * TN + (col - TM);
            if (j < P.c) base = P.X + (int64_t)(P.p_old + j) * P.ld;
        }
        src_col[q] = base;
        dst_off[q] = col * STRIDE + seg * 2;
        seg_row[q] = seg * 2;
    }

    auto load_stage = [&](int kt, int buf) {
        double *dst = smem + (size_t)buf * kStageDoubles;
        const int64_t r0 = n_lo + (int64_t)kt * KB;
#pragma unroll
        for (int q = 0; q < 8; ++q) {
            if (src_col[q] == nullptr) continue;
            int64_t r = r0 + seg_row[q];
            int64_t left = n_hi - r;
            int bytes = left >= 2 ? 16 : (left == 1 ? 8 : 0);
            const double *src = bytes ? src_col[q] + r : src_col[q];
            cp_async16(dst + dst_off[q], src, bytes);
        }
    };

    double acc[4][2][2];
#pragma unroll
    for (int a = 0; a < 4; ++a)
#pragma unroll
        for (int b = 0; b < 2; ++b) acc[a][b][0] = acc[a][b][1] = 0.0;

    // 8 x 8 output fragments this warp really has to compute: inside the (p + 1) x c block and not strictly below
    // the diagonal of the symmetric X_new' X_new part (fokl_gram_scatter mirrors the upper triangle)
    unsigned fmask = 0;
#pragma unroll
    for (int mt = 0; mt < 4; ++mt)
#pragma unroll
        for (int nt = 0; nt < 2; ++nt) {
            const int i0 = ta * TM + wm * 32 + mt * 8, j0 = tb * TN + wn * 16 + nt * 8;
            const bool has_y = (i0 <= p) && (p < i0 + 8);
            const bool below = (i0 - P.p_old > j0 + 7) && !has_y;
            if (i0 <= p && j0 < P.c && !below) fmask |= 1u << (mt * 2 + nt);
        }
    const bool cta_active = __syncthreads_or(fmask != 0);

    if (cta_active) {
#pragma unroll
        for (int s = 0; s < STAGES - 1; ++s) {
            if (s < nk) load_stage(s, s);
            cp_async_commit();
        }
        const int frag_r = lane >> 2, frag_k = lane & 3;
        for (int kt = 0; kt < nk; ++kt) {
            cp_async_wait<STAGES - 2>();
            __syncthreads();
            {
                int nxt = kt + STAGES - 1;
                if (nxt < nk) load_stage(nxt, nxt % STAGES);
                cp_async_commit();
            }
            if (fmask == 0) continue;
            const double *As = smem + (size_t)(kt % STAGES) * kStageDoubles;
            const double *Bs = As + (size_t)TM * STRIDE;
#pragma unroll
            for (int kk = 0; kk < KB / 4; ++kk) {
                double af[4], bf[2];
#pragma unroll
                for (int mt = 0; mt < 4; ++mt)
                    if (fmask & (3u << (mt * 2))) af[mt] = As[(wm * 32 + mt * 8 + frag_r) * STRIDE + kk * 4 + frag_k];
#pragma unroll
                for (int nt = 0; nt < 2; ++nt)
                    if (fmask & (0x55u << nt)) bf[nt] = Bs[(wn * 16 + nt * 8 + frag_r) * STRIDE + kk * 4 + frag_k];
#pragma unroll
                for (int mt = 0; mt < 4; ++mt)
#pragma unroll
                    for (int nt = 0; nt < 2; ++nt)
                        if (fmask & (1u << (mt * 2 + nt))) dmma_m8n8k4(acc[mt][nt][0], acc[mt][nt][1], af[mt], bf[nt]);
            }
        }
        cp_async_wait<0>();
    }

    double *out = P.out + (size_t)blockIdx.y * (size_t)(p + 1) * P.c;
#pragma unroll
    for (int mt = 0; mt < 4; ++mt)
#pragma unroll
        for (int nt = 0; nt < 2; ++nt) {
            int i = ta * TM + wm * 32 + mt * 8 + (lane >> 2);
            int j = tb * TN + wn * 16 + nt * 8 + (lane & 3) * 2;
            if (i <= p) {
                if (j < P.c) out[(size_t)i * P.c + j] = acc[mt][nt][0];
                if (j + 1 < P.c) out[(size_t)i * P.c + j + 1] = acc[mt][nt][1];
            }
        }
}

__global__ void gram_reduce_kernel(const double *__restrict__ part, int nsplit, int64_t elems, double *__restrict__ out)
{
    for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < elems; e += (int64_t)gridDim.x * blockDim.x) {
        double s = 0.0;
        for (int k = 0; k < nsplit; ++k) s += part[(size_t)k * elems + e];
        out[e] = s;
    }
}

}  // namespace

extern "C" int fokl_gram_update(fokl_ctx *ctx, const double *X, int64_t ld, int64_t n, int p_old, int c,
                                const double *y, double *block)
{
    FOKL_CHECK_CTX(ctx);
    if (!X || !y || !block || n < 1 || p_old < 0 || c < 1 || ld < n)
        FOKL_FAIL(ctx, FOKL_EINVAL, "gram_update: bad argument");
    if ((ld % 2) != 0 || ((uintptr_t)X % 16) != 0 || ((uintptr_t)y % 16) != 0)
        FOKL_FAIL(ctx, FOKL_EINVAL, "gram_update: X and y must be 16-byte aligned and ld even");
    int rc = fokl_bind_device(ctx);
    if (rc) return rc;
    const int p = p_old + c;
    const int tiles_a = (p + 1 + TM - 1) / TM, tiles_b = (c + TN - 1) / TN;
    const int tiles = tiles_a * tiles_b;
    const int64_t chunks = (n + KB - 1) / KB;
    int64_t want = ((int64_t)ctx->num_sms * 4 + tiles - 1) / tiles;          // ~2 waves of 2 CTAs/SM
    int64_t max_split = std::max<int64_t>(1, chunks / 8);                    // >= 8 slabs per CTA
    int nsplit = (int)std::max<int64_t>(1, std::min<int64_t>(std::min<int64_t>(want, max_split), 65535));
    int64_t rows_per_split = ((chunks + nsplit - 1) / nsplit) * KB;
    nsplit = (int)((n + rows_per_split - 1) / rows_per_split);
    const int64_t elems = (int64_t)(p + 1) * c;

    GramParams P;
    P.X = X; P.y = y; P.ld = ld; P.n = n; P.p_old = p_old; P.c = c; P.tiles_b = tiles_b;
    P.rows_per_split = rows_per_split;
    if (nsplit == 1) {
        P.out = block;
    } else {
        double *part = (double *)fokl_scratch(ctx, fokl_ctx::B_GRAM, (size_t)nsplit * elems * sizeof(double));
        if (!part) return FOKL_ENOMEM;
        P.out = part;
    }
    FOKL_CUDA(ctx, cudaFuncSetAttribute(gram_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kGramSmem));
    gram_kernel<<<dim3(tiles, nsplit), kGramThreads, kGramSmem, ctx->stream>>>(P);
    FOKL_LAUNCH_CHECK(ctx);
    if (nsplit > 1) {
        int blocks = (int)std::min<int64_t>((elems + 255) / 256, 1024);
        gram_reduce_kernel<<<blocks, 256, 0, ctx->stream>>>(P.out, nsplit, elems, block);
        FOKL_LAUNCH_CHECK(ctx);
    }
    return FOKL_OK;
}
