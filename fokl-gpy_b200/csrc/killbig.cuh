// killbig.cuh -- the kill loop (FR:1669-1690) of a wide model on the whole device.
//
// Same mathematics and the same per-entry arithmetic as fokl::kill_loop (cand_math.cuh): the sweep-operator tableau of
// [G Xty; Xty' yty] (centred), p forward sweeps, then per accepted kill one reverse sweep; every proposal's BIC is O(1)
// from the tableau's diagonal and last row.  The single-CTA kernel keeps the tableau in shared memory (p <= 238) or
// walks it in L2 alone; for the 3-way substages of a 16-input problem (p ~ 2000: 34 MB of tableau, ~2000 forward
// pivots and ~1400 accepted kills, each a rank-1 update of the whole tableau) that takes seconds.  Here the rows of the
// tableau are dealt round-robin to the CTAs of a cooperative launch, one grid barrier per pivot:
//   * forward sweep k: everybody needs pivot row k as it is after sweep k - 1.  Its owner handles row k first in sweep
//     k - 1 and publishes the finished row in a double-buffered broadcast area, so sweep k starts right after the
//     barrier and the owner can overwrite row k in place while the others still read the broadcast copy;
//   * reverse sweep q (a kill): row and column q are dead afterwards -- no later pivot reads them -- so the owner simply
//     never writes row q again and everybody reads it in place; dead rows are skipped from then on;
//   * the proposal scan is done redundantly by every CTA from the diagonal and the last row: identical arithmetic,
//     identical decisions, no broadcast.  Both are kept as contiguous, double-buffered copies (written for the next
//     round while the current round's copy is still being scanned by slower CTAs).
#pragma once
#include "cand_math.cuh"
#include "eigbig.cuh"

namespace killb {

constexpr int kThreads = 512;
constexpr int kWarps = kThreads / 32;

struct Params {
    const double *G;
    int64_t ldg;
    const double *Xty;
    const int32_t *cols, *cand_pos;
    const double *bv0, *bv1;
    int p, vm, ldt;
    fokl::CandConst c;
    fokl::KillLoopIn in;
    double *T;           // [p + 1][ldt] (+ one row: entry [p + 1][0] = 1.0 if pre != 0 and the tableau is usable)
    int pre;             // != 0: T already holds the tableau after the p forward sweeps (kill_tableau_kernel)
    double *dg;          // [2][ldt] diagonal of T (buffer r & 1 is read in kill round r)
    double *last;        // [2][ldt] last row of T, likewise
    double *bcast;       // [2][ldt] forward sweeps: the next pivot row
    unsigned *bar;       // zeroed before the launch
    int32_t *out_i;
    double *out_ev;
};

// one row of a sweep: pivot row copy `rowbuf` (shared), pivot index k, this row j != k
__device__ __forceinline__ void sweep_row(double *Tj, const double *rowbuf, int ld, int k, int j, double sign, double invD,
                                          int lane, double *dg, double *pub)
{
    const double ckj = rowbuf[j];
    const double vk = sign * ckj * invD;
    int i = lane;
    for (; i + 96 < ld; i += 128) {
        const double a0 = __ldcg(Tj + i), a1 = __ldcg(Tj + i + 32), a2 = __ldcg(Tj + i + 64), a3 = __ldcg(Tj + i + 96);
        double r0 = a0 - (rowbuf[i] * ckj) * invD, r1 = a1 - (rowbuf[i + 32] * ckj) * invD;
        double r2 = a2 - (rowbuf[i + 64] * ckj) * invD, r3 = a3 - (rowbuf[i + 96] * ckj) * invD;
        if (i == k) r0 = vk;
        if (i + 32 == k) r1 = vk;
        if (i + 64 == k) r2 = vk;
        if (i + 96 == k) r3 = vk;
        Tj[i] = r0; Tj[i + 32] = r1; Tj[i + 64] = r2; Tj[i + 96] = r3;
        if (pub) { pub[i] = r0; pub[i + 32] = r1; pub[i + 64] = r2; pub[i + 96] = r3; }
        if (i == j) dg[j] = r0;
        if (i + 32 == j) dg[j] = r1;
        if (i + 64 == j) dg[j] = r2;
        if (i + 96 == j) dg[j] = r3;
    }
    for (; i < ld; i += 32) {
        double r0 = __ldcg(Tj + i) - (rowbuf[i] * ckj) * invD;
        if (i == k) r0 = vk;
        Tj[i] = r0;
        if (pub) pub[i] = r0;
        if (i == j) dg[j] = r0;
    }
}

__global__ void __launch_bounds__(kThreads, 1) kill_loop_big_kernel(const Params P)
{
    extern __shared__ __align__(16) double sh[];
    const int p = P.p, ld = p + 1, ldt = P.ldt, vm = P.vm;
    double *rowbuf = sh;                                              // ld doubles (+ pad)
    unsigned char *dead = reinterpret_cast<unsigned char *>(sh + ((ld + 2) & ~1));   // ld bytes
    __shared__ int shi[4];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int nc = gridDim.x, cta = blockIdx.x;
    eigb::TeamBarrier bar;
    bar.ctr = P.bar; bar.target = 0; bar.team = nc;
    const int32_t *idx = P.cols;
    const fokl::CandConst &c = P.c;
    const double ybar = c.sum_y / c.n;
    const int64_t row0 = (int64_t)idx[0] * P.ldg;
    double *T = P.T;

    // ---- tableau [G x; x' yty] (centred like fokl::kill_loop_t), rows dealt round-robin ------------------------------------
    // (pre: the swept tableau was formed from the model's eigendecomposition -- only the diagonal and the last row of
    // the first kill round are copied out, the p pivots with one grid barrier each are not run)
    const bool pre = P.pre != 0 && __ldcg(T + (int64_t)ld * ldt) > 0.5;
    if (pre) {
        for (int i = cta * kThreads + tid; i < ld; i += nc * kThreads) {
            P.dg[i] = __ldcg(T + (int64_t)i * ldt + i);
            P.last[i] = __ldcg(T + (int64_t)p * ldt + i);
        }
    }
    for (int j = cta + nc * warp; j < ld && !pre; j += nc * kWarps) {
        double *Tj = T + (int64_t)j * ldt;
        for (int i = lane; i < ld; i += 32) {
            double v;
            if (i < p && j < p) v = P.G[(int64_t)idx[i] * P.ldg + idx[j]];
            else if (i == p && j == p) v = c.yty - c.n * ybar * ybar;
            else {
                const int q = i < p ? i : j;
                v = P.Xty[idx[q]] - ybar * P.G[row0 + idx[q]];
            }
            Tj[i] = v;
            if (j == 0) P.bcast[i] = v;
        }
    }
    for (int i = tid; i < ld; i += kThreads) dead[i] = 0;
    bar.sync();

    // ---- forward sweeps -----------------------------------------------------------------------------------------------------
    int bad = 0;
    for (int k = 0; k < p && !pre; ++k) {
        const double *src = P.bcast + (size_t)(k & 1) * ldt;
        for (int i = tid; i < ld; i += kThreads) rowbuf[i] = __ldcg(src + i);
        __syncthreads();
        const double d = rowbuf[k];
        if (!(d > 1e-11 * P.G[(int64_t)idx[k] * P.ldg + idx[k]])) { bad = 1; break; }   // same value in every CTA
        const double invD = 1.0 / d;
        // the next pivot row first, published for everybody
        const int kn = k + 1;
        const bool own_next = kn < p && (kn % nc) == cta;
        const int wnext = own_next ? (kn / nc) % kWarps : -1;
        if (warp == wnext)
            sweep_row(T + (int64_t)kn * ldt, rowbuf, ld, k, kn, 1.0, invD, lane, P.dg, P.bcast + (size_t)(kn & 1) * ldt);
        const bool final_sweep = k == p - 1;      // the kill rounds start from this sweep's diagonal and last row
        for (int j = cta + nc * warp; j < ld; j += nc * kWarps) {
            if (j == kn && own_next) continue;
            double *Tj = T + (int64_t)j * ldt;
            if (j == k) {
                for (int i = lane; i < ld; i += 32) {
                    const double v = (i == k) ? -invD : rowbuf[i] * invD;
                    Tj[i] = v;
                    if (i == k) P.dg[k] = v;
                }
                continue;
            }
            sweep_row(Tj, rowbuf, ld, k, j, 1.0, invD, lane, P.dg, (final_sweep && j == p) ? P.last : nullptr);
        }
        bar.sync();
    }

    // ---- kill rounds --------------------------------------------------------------------------------------------------------
    int n_acc = 0, tested = 0, pa = p, cur = P.in.start;
    double evmin = P.in.evmin;
    const double ln_n = log(c.n);
    const double thr = P.in.threshav * P.in.icpt;
    for (int rnd = 0; !bad && cur < vm; ++rnd) {
        const double *dgr = P.dg + (size_t)(rnd & 1) * ldt, *Tp = P.last + (size_t)(rnd & 1) * ldt;
        double *dgw = P.dg + (size_t)((rnd + 1) & 1) * ldt, *lastw = P.last + (size_t)((rnd + 1) & 1) * ldt;
        if (tid == 0) { shi[0] = 0x7fffffff; shi[1] = 0; shi[2] = 0; }
        __syncthreads();
        const double sse = __ldcg(dgr + p);
        for (int i = cur + tid; i < vm; i += kThreads) {
            const bool prop = (P.bv1[i] > P.in.threshstdb) || (P.bv1[i] > P.in.threshstda && P.bv0[i] < thr);
            if (!prop) continue;
            const int q = P.cand_pos[i];
            const double tqq = __ldcg(dgr + q), tqy = __ldcg(Tp + q);
            if (!(tqq < 0.0)) { shi[2] = 1; continue; }
            const double sig = (sse + tqy * tqy / (-tqq)) / c.n;
            const double evt = (double)(pa - 1) * ln_n - 2.0 * (-(c.n / 2.0) * log(sig) - (c.n - 1.0) / 2.0) +
                               P.in.aic_adj * (double)(pa - 1);
            if (evt < evmin) atomicMin(&shi[0], i);
        }
        __syncthreads();
        const int hit = shi[0];
        if (shi[2]) { bad = 2; break; }
        for (int i = cur + tid; i < vm && i <= hit; i += kThreads) {
            const bool prop = (P.bv1[i] > P.in.threshstdb) || (P.bv1[i] > P.in.threshstda && P.bv0[i] < thr);
            if (prop) atomicAdd(&shi[1], 1);
        }
        __syncthreads();
        tested += shi[1];
        if (hit == 0x7fffffff) break;
        const int q = P.cand_pos[hit];
        const double tqq = __ldcg(dgr + q), tqy = __ldcg(Tp + q);
        const double sig = (sse + tqy * tqy / (-tqq)) / c.n;
        evmin = (double)(pa - 1) * ln_n - 2.0 * (-(c.n / 2.0) * log(sig) - (c.n - 1.0) / 2.0) + P.in.aic_adj * (double)(pa - 1);
        // reverse sweep of pivot q: row q is read in place (its owner leaves it alone from now on)
        const double *src = T + (int64_t)q * ldt;
        for (int i = tid; i < ld; i += kThreads) rowbuf[i] = __ldcg(src + i);
        if (tid == 0) dead[q] = 1;
        __syncthreads();
        const double invD = 1.0 / rowbuf[q];
        for (int j = cta + nc * warp; j < ld; j += nc * kWarps) {
            if (dead[j]) continue;
            sweep_row(T + (int64_t)j * ldt, rowbuf, ld, q, j, -1.0, invD, lane, dgw, j == p ? lastw : nullptr);
        }
        if (cta == 0 && tid == 0) {
            P.out_i[3 + n_acc] = hit;
            P.out_i[3 + vm + n_acc] = tested;
            P.out_ev[n_acc] = evmin;
        }
        n_acc += 1;
        pa -= 1;
        cur = hit + 1;
        bar.sync();
    }
    if (cta == 0 && tid == 0) { P.out_i[0] = n_acc; P.out_i[1] = tested; P.out_i[2] = bad; }
}

}  // namespace killb
