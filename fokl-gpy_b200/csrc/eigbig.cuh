// eigbig.cuh -- K3a for wide candidate models: Cholesky + blocked one-sided Jacobi eigensolver, many CTAs per model.
//
// Replaces `Lamb, Q = eigh(XtX)` (src/FoKL/FoKLRoutines.py:1499) where the cluster eigensolver of candidates.cu
// (cand_eigj_kernel: every column resident in the shared memory of one thread-block cluster) no longer fits -- models
// of more than ~700 columns, i.e. the 3-way substages of a 16-input problem (BASELINE.json configs[4]: up to 1680 new
// columns on top of the accepted model).
//
// Method (the same mathematics as the cluster solver, Veselic-Hari): A = L L' (Cholesky), then right rotations
// W <- W V on W = L until the columns of W are mutually orthogonal; then W = U diag(sigma) with A = U diag(sigma^2) U',
// i.e. lambda_j = |w_j|^2 and w_j / |w_j| is the eigenvector.  Working on the factor keeps the condition number of the
// iteration at sqrt(cond(A)) and gives the small eigenvalues (which 1 / Lamb of FR:1502 amplifies) to high relative
// accuracy.  The rotations are applied block-wise (one-sided block Jacobi):
//   * the columns are cut into nb blocks of 8; a sweep is nb - 1 rounds of a round-robin tournament in which every
//     block meets every other block once; the nb / 2 block pairs of a round are independent;
//   * a pair visit by ONE CTA: (1) H = [W_I W_J]' [W_I W_J] (16 x 16) on the FP64 tensor pipe (mma.sync.m8n8k4.f64,
//     operands streamed from L2 as 16-byte loads, both operands of the symmetric product from the same registers),
//     (2) one cyclic two-sided Jacobi sweep on H by one warp in shared memory, accumulating the 16 x 16 rotation R,
//     (3) [W_I W_J] <- [W_I W_J] R, again DMMA, in place (16-row slabs); pairs that are already orthogonal to the
//     tolerance skip (2) and (3);
//   * a sweep without any rotation ends the iteration (the test |h_ij| <= tol sqrt(h_ii h_jj) is made on Gram entries
//     freshly computed from W, exactly like the cluster solver's).
// W lives in global memory and stays L2-resident (a 2072-column model is 34 MB of the 126 MB L2).  A model is worked on
// by a *team* of co-resident CTAs (cooperative launch), pair slot k of every round belongs to CTA k mod team.  There is
// NO barrier between rounds: a visit waits only for the two visits of the previous round that produced its blocks
// (a per-block round counter in global memory, release / acquire), so fast visits (already orthogonal pairs) run ahead.
// The team meets at a counter barrier once per sweep (convergence flag) and in the Cholesky phase (two per 32-column
// panel).  A launch processes any number of models; teams fetch the next one from a dispenser.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace eigb {

constexpr int kThreads = 512;
constexpr int kWarps = kThreads / 32;
constexpr int kB = 8;                    // columns per block
constexpr int kB2 = 2 * kB;              // columns per block pair
constexpr int kMaxSweeps = 40;
constexpr int kFlagStride = kMaxSweeps + 8;
constexpr int kPanel = 32;               // Cholesky panel width
constexpr int kTile = 64;                // trailing-update tile

struct Job {
    int32_t cand;            // candidate index of the batch
    int32_t p;               // model width
    int32_t ld;              // rows of W: p rounded up to 16 (rows >= p are zero)
    int32_t nb;              // column blocks (even); W has nb * 8 columns, columns >= p are zero and never rotate
    int32_t set_off;         // into col_sets
    int32_t pad;
    int64_t w_off;           // doubles into the W workspace
    int64_t vec_off, mat_off;
};

struct Params {
    const double *G;
    int64_t ldg;
    const int32_t *col_sets;
    const Job *jobs;
    int n_jobs;
    int team;                // CTAs per team
    double *W;
    double *lam_raw;         // per candidate at vec_off + 64 * cand: squared norm of column j
    double *lamb, *Q;        // outputs packed like fokl_candidates_eval's
    int32_t *status;         // per candidate: number of sweeps, or -1 (Gram not positive definite: nothing written)
    unsigned *bar;           // [n_teams] barrier counters (zeroed before the launch)
    int *slot;               // [n_teams] job the team is working on
    int *next;               // job dispenser (zeroed before the launch)
    int *flags;              // [n_teams][kFlagStride] "some pair rotated in sweep s"
    int *ver;                // [n_teams][ver_stride] rounds completed on block i of the team's current model
    int ver_stride;
    int max_inner;           // inner Jacobi sweeps per visit
    int sort_diag;           // 1: symmetric permutation by decreasing diagonal before the Cholesky factorisation
};

__device__ __forceinline__ void dmma884(double &c0, double &c1, double a, double b)
{
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
                 : "+d"(c0), "+d"(c1)
                 : "d"(a), "d"(b));
}

__device__ __forceinline__ double2 ldcg2(const double *p) { return __ldcg(reinterpret_cast<const double2 *>(p)); }

__device__ __forceinline__ int ld_acquire(const int *p)
{
    int v;
    asm volatile("ld.acquire.gpu.global.s32 %0, [%1];\n" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_release(int *p, int v)
{
    asm volatile("st.release.gpu.global.s32 [%0], %1;\n" ::"l"(p), "r"(v) : "memory");
}

// All CTAs of a team arrive; nobody leaves before everybody has arrived.  Global writes made before the barrier by any
// thread of the team are visible (through L2: readers use ld.global.cg) to every thread after it.
struct TeamBarrier {
    unsigned *ctr;
    unsigned target;
    int team;
    __device__ __forceinline__ void sync()
    {
        __syncthreads();
        if (team > 1 && threadIdx.x == 0) {
            target += (unsigned)team;
            __threadfence();
            atomicAdd(ctr, 1u);
            unsigned v;
            do {
                asm volatile("ld.acquire.gpu.global.u32 %0, [%1];\n" : "=r"(v) : "l"(ctr) : "memory");
            } while ((int)(v - target) < 0);
        }
        __syncthreads();
    }
};

// ---- shared memory: the Jacobi visit and the Cholesky phase use the same bytes at different times -------------------
constexpr int kLDH = kB2 + 1;      // H row stride (doubles): column walks are conflict-free
constexpr int kLDR = kB2 + 4;      // R^T row stride: B-fragment loads conflict-free
constexpr int kRedDoubles = kWarps * kB2 * kB2;
constexpr int kVisitDoubles = kRedDoubles + kB2 * kLDH + kB2 * kLDR + 2 * kB2 + 2 * kB2 /* ints */ + 8;
constexpr int kLDL = kPanel + 1;
constexpr int kLDP = kTile + 2;
constexpr int kCholDoubles = kPanel * kLDL + kPanel + 2 * kPanel * kLDP + 8;
constexpr int kSmemDoubles = kVisitDoubles > kCholDoubles ? kVisitDoubles : kCholDoubles;
constexpr size_t kSmemBytes = (size_t)kSmemDoubles * sizeof(double);

// One visit of the block pair (columns colI .. colI + 7 and colJ .. colJ + 7 of W).  Returns (to every thread) whether
// any plane rotation was applied.
__device__ bool visit_pair(double *W, int ld, int colI, int colJ, double tol2, int max_inner, double *sm)
{
    double *red = sm;                       // [kWarps][16 * 16]
    double *Hs = red + kRedDoubles;         // [16][kLDH]
    double *Rt = Hs + kB2 * kLDH;           // [16][kLDR]   Rt[n][k] = R[k][n]
    double *rc = Rt + kB2 * kLDR;           // [8] cos, [8] sin
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int lq = lane >> 2, lr = lane & 3;

    // ---- (1) H = X' X, X = [W_I W_J] (ld x 16): every warp takes every kWarps-th group of 8 rows -------------------------
    {
        double a00[2] = {0.0, 0.0}, a01[2] = {0.0, 0.0}, a11[2] = {0.0, 0.0};
        const double *c0 = W + (size_t)(colI + lq) * ld + 2 * lr;     // rows 2 lr, 2 lr + 1 of every 8-row group
        const double *c1 = W + (size_t)(colJ + lq) * ld + 2 * lr;
        const int ngroups = ld >> 3;
        for (int g = warp; g < ngroups; g += 4 * kWarps) {           // four groups in flight
            double2 u[4], v[4];
#pragma unroll
            for (int x = 0; x < 4; ++x) {
                const int gg = g + x * kWarps;
                if (gg < ngroups) { u[x] = ldcg2(c0 + 8 * gg); v[x] = ldcg2(c1 + 8 * gg); }
                else { u[x] = make_double2(0.0, 0.0); v[x] = make_double2(0.0, 0.0); }
            }
#pragma unroll
            for (int x = 0; x < 4; ++x) {
                dmma884(a00[0], a00[1], u[x].x, u[x].x);
                dmma884(a01[0], a01[1], u[x].x, v[x].x);
                dmma884(a11[0], a11[1], v[x].x, v[x].x);
                dmma884(a00[0], a00[1], u[x].y, u[x].y);
                dmma884(a01[0], a01[1], u[x].y, v[x].y);
                dmma884(a11[0], a11[1], v[x].y, v[x].y);
            }
        }
        double *mine = red + warp * (kB2 * kB2);
        *reinterpret_cast<double2 *>(mine + lq * kB2 + 2 * lr) = make_double2(a00[0], a00[1]);
        *reinterpret_cast<double2 *>(mine + lq * kB2 + 8 + 2 * lr) = make_double2(a01[0], a01[1]);
        *reinterpret_cast<double2 *>(mine + (8 + lq) * kB2 + 8 + 2 * lr) = make_double2(a11[0], a11[1]);
    }
    __syncthreads();
    if (tid < kB2 * kB2) {
        const int m = tid >> 4, n = tid & 15;
        if (m <= n) {
            double s = 0.0;
#pragma unroll
            for (int w = 0; w < kWarps; ++w) s += red[w * (kB2 * kB2) + tid];
            Hs[m * kLDH + n] = s;
            Hs[n * kLDH + m] = s;
        }
        Rt[m * kLDR + n] = (m == n) ? 1.0 : 0.0;
    }
    __syncthreads();
    {
        int viol = 0;
        if (tid < kB2 * kB2) {
            const int m = tid >> 4, n = tid & 15;
            if (m < n) {
                const double ga = Hs[m * kLDH + n];
                viol = ga * ga > tol2 * (Hs[m * kLDH + m] * Hs[n * kLDH + n]);
            }
        }
        if (!__syncthreads_or(viol)) return false;
    }

    // ---- (2) cyclic two-sided Jacobi on H by warp 0, R accumulated ---------------------------------------------------------
    // Per round: lanes 0..7 compute the 8 rotations; then every lane applies both sides at once to two of the 64 2 x 2
    // blocks (rows i_k, j_k x columns i_l, j_l: columns first, then rows, like the scalar form) and rotates four
    // entries of R -- all loads, then the arithmetic, then all stores, so the shared-memory latency is paid once per
    // round instead of once per element.
    if (warp == 0) {
        constexpr int nm1 = kB2 - 1, half = kB2 / 2;
        double *rs = rc + half;
        for (int isw = 0; isw < max_inner; ++isw) {
            int any = 0;
            for (int r = 0; r < nm1; ++r) {
                auto pair_of = [&](int k, int &i, int &j) {
                    if (k == 0) { i = nm1; j = r; }
                    else {
                        i = r + k; if (i >= nm1) i -= nm1;
                        j = r - k; if (j < 0) j += nm1;
                    }
                };
                if (lane < half) {
                    int i, j;
                    pair_of(lane, i, j);
                    const double al = Hs[i * kLDH + i], be = Hs[j * kLDH + j], ga = Hs[i * kLDH + j];
                    double c = 1.0, s = 0.0;
                    if (ga * ga > tol2 * (al * be)) {
                        // rotation that annihilates ga (same formulas as the cluster solver, candidates.cu pair_rotate_reg)
                        const double d = be - al;
                        const double rinv = rsqrt(fma(d, d, 4.0 * ga * ga));
                        const double u = fma(0.5 * fabs(d), rinv, 0.5);
                        const double ic = rsqrt(u);
                        c = u * ic;
                        s = copysign(ga * rinv, ga * d) * ic;
                        any = 1;
                    }
                    rc[lane] = c; rs[lane] = s;
                }
                __syncwarp();
                int ik[2], jk[2], il[2], jl[2];
                double ck[2], sk[2], cl[2], sl[2], b00[2], b01[2], b10[2], b11[2];
#pragma unroll
                for (int h = 0; h < 2; ++h) {
                    const int blk = lane + 32 * h, k = blk >> 3, l = blk & 7;
                    pair_of(k, ik[h], jk[h]);
                    pair_of(l, il[h], jl[h]);
                    ck[h] = rc[k]; sk[h] = rs[k]; cl[h] = rc[l]; sl[h] = rs[l];
                    b00[h] = Hs[ik[h] * kLDH + il[h]]; b01[h] = Hs[ik[h] * kLDH + jl[h]];
                    b10[h] = Hs[jk[h] * kLDH + il[h]]; b11[h] = Hs[jk[h] * kLDH + jl[h]];
                }
                int ir[4], jr[4], xr[4];
                double cr_[4], sr_[4], ra[4], rb[4];
#pragma unroll
                for (int q = 0; q < 4; ++q) {
                    const int it = lane + 32 * q, k = it >> 4;
                    xr[q] = it & 15;
                    pair_of(k, ir[q], jr[q]);
                    cr_[q] = rc[k]; sr_[q] = rs[k];
                    ra[q] = Rt[ir[q] * kLDR + xr[q]]; rb[q] = Rt[jr[q] * kLDR + xr[q]];
                }
#pragma unroll
                for (int h = 0; h < 2; ++h) {
                    // columns (rotation l), then rows (rotation k)
                    const double t00 = cl[h] * b00[h] - sl[h] * b01[h], t01 = sl[h] * b00[h] + cl[h] * b01[h];
                    const double t10 = cl[h] * b10[h] - sl[h] * b11[h], t11 = sl[h] * b10[h] + cl[h] * b11[h];
                    Hs[ik[h] * kLDH + il[h]] = ck[h] * t00 - sk[h] * t10;
                    Hs[ik[h] * kLDH + jl[h]] = ck[h] * t01 - sk[h] * t11;
                    Hs[jk[h] * kLDH + il[h]] = sk[h] * t00 + ck[h] * t10;
                    Hs[jk[h] * kLDH + jl[h]] = sk[h] * t01 + ck[h] * t11;
                }
#pragma unroll
                for (int q = 0; q < 4; ++q) {
                    Rt[ir[q] * kLDR + xr[q]] = cr_[q] * ra[q] - sr_[q] * rb[q];
                    Rt[jr[q] * kLDR + xr[q]] = sr_[q] * ra[q] + cr_[q] * rb[q];
                }
                __syncwarp();
            }
            if (!__any_sync(0xffffffffu, any)) break;
        }
    }
    __syncthreads();

    // ---- (3) X <- X R in place: every warp takes every kWarps-th slab of 16 rows -------------------------------------------
    {
        double bf[2][4];
#pragma unroll
        for (int nq = 0; nq < 2; ++nq)
#pragma unroll
            for (int kq = 0; kq < 4; ++kq) bf[nq][kq] = Rt[(8 * nq + lq) * kLDR + 4 * kq + lr];   // R[k = 4 kq + lr][n = 8 nq + lq]
        const double *ap[4];
#pragma unroll
        for (int kq = 0; kq < 4; ++kq) {
            const int k = 4 * kq + lr;
            const int col = k < kB ? colI + k : colJ + (k - kB);
            ap[kq] = W + (size_t)col * ld + 2 * lq;                       // rows 2 lq, 2 lq + 1 of the slab
        }
        double *op[2][2];
#pragma unroll
        for (int nq = 0; nq < 2; ++nq)
#pragma unroll
            for (int h = 0; h < 2; ++h) {
                const int k = 8 * nq + 2 * lr + h;
                const int col = k < kB ? colI + k : colJ + (k - kB);
                op[nq][h] = W + (size_t)col * ld + 2 * lq;
            }
        const int nslabs = ld >> 4;
        // one slab per iteration, the next slab's loads issued before this slab's arithmetic and stores
        double2 a[4], nx[4];
        int sl = warp;
        if (sl < nslabs) {
#pragma unroll
            for (int kq = 0; kq < 4; ++kq) a[kq] = ldcg2(ap[kq] + 16 * sl);
        }
        for (; sl < nslabs; sl += kWarps) {
            const int r0 = 16 * sl;
            const bool more = sl + kWarps < nslabs;
#pragma unroll
            for (int kq = 0; kq < 4; ++kq) nx[kq] = more ? ldcg2(ap[kq] + r0 + 16 * kWarps) : make_double2(0.0, 0.0);
#pragma unroll
            for (int nq = 0; nq < 2; ++nq) {
                double e0 = 0.0, e1 = 0.0, o0 = 0.0, o1 = 0.0;
#pragma unroll
                for (int kq = 0; kq < 4; ++kq) {
                    dmma884(e0, e1, a[kq].x, bf[nq][kq]);
                    dmma884(o0, o1, a[kq].y, bf[nq][kq]);
                }
                // lane holds rows (2 lq, 2 lq + 1) of columns 8 nq + 2 lr + {0, 1}
                *reinterpret_cast<double2 *>(op[nq][0] + r0) = make_double2(e0, o0);
                *reinterpret_cast<double2 *>(op[nq][1] + r0) = make_double2(e1, o1);
            }
#pragma unroll
            for (int kq = 0; kq < 4; ++kq) a[kq] = nx[kq];
        }
    }
    return true;
}

// Blocks of pair k in round r of the round-robin tournament over nb players (nb even): player nb - 1 stays, the others
// rotate.  Every two blocks meet exactly once in rounds 0 .. nb - 2.
__device__ __forceinline__ void tournament_pair(int nb, int r, int k, int &bi, int &bj)
{
    const int nm1 = nb - 1;
    if (k == 0) { bi = nm1; bj = r; }
    else {
        bi = r + k; if (bi >= nm1) bi -= nm1;
        bj = r - k; if (bj < 0) bj += nm1;
    }
}

// ---- Cholesky of the lower triangle held in W (ld x ncols, column-major), right-looking, 32-column panels -------------------
// Returns false (in every thread of every CTA of the team) if a pivot is not positive.
__device__ bool team_cholesky(TeamBarrier &bar, double *W, int p, int ld, int cr, int T, double *sm)
{
    double *Ls = sm;                           // [32][kLDL] factor of the diagonal block
    double *inv = Ls + kPanel * kLDL;          // [32] 1 / L_jj
    double *Pr = inv + kPanel;                 // [32][kLDP] panel rows of the tile's row range, k-major
    double *Pc = Pr + kPanel * kLDP;           // [32][kLDP] ... of the tile's column range
    int *okf = reinterpret_cast<int *>(Pc + kPanel * kLDP);
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    for (int k0 = 0; k0 < p; k0 += kPanel) {
        const int w = (p - k0 < kPanel) ? p - k0 : kPanel;
        // (a) every CTA factors the diagonal block itself (same bits everywhere): lane r of warp 0 holds row r
        if (warp == 0) {
            double a[kPanel];
#pragma unroll
            for (int j = 0; j < kPanel; ++j)
                a[j] = (lane < w && j <= lane) ? __ldcg(W + (size_t)(k0 + j) * ld + k0 + lane) : ((j == lane) ? 1.0 : 0.0);
            bool ok = true;
#pragma unroll
            for (int k = 0; k < kPanel; ++k) {
                const double d = __shfl_sync(0xffffffffu, a[k], k);
                if (!(d > 0.0) || !(d < 1.7e308)) ok = false;
                const double r = sqrt(d);
                if (lane == k) a[k] = r;
                else if (lane > k) a[k] = a[k] / r;
#pragma unroll
                for (int j = k + 1; j < kPanel; ++j) {
                    const double ljk = __shfl_sync(0xffffffffu, a[k], j);
                    if (lane >= j) a[j] -= a[k] * ljk;
                }
            }
#pragma unroll
            for (int j = 0; j < kPanel; ++j) Ls[lane * kLDL + j] = (j <= lane) ? a[j] : 0.0;
            inv[lane] = 1.0 / a[lane];
            if (lane == 0) okf[0] = ok ? 1 : 0;
        }
        __syncthreads();
        if (!okf[0]) return false;
        // (b) panel rows below the block: x L_D' = b, one row per thread
        for (int r = k0 + w + cr * kThreads + tid; r < p; r += T * kThreads) {
            double x[kPanel];
#pragma unroll
            for (int j = 0; j < kPanel; ++j) x[j] = (j < w) ? __ldcg(W + (size_t)(k0 + j) * ld + r) : 0.0;
#pragma unroll
            for (int j = 0; j < kPanel; ++j) {
                double s = x[j];
#pragma unroll
                for (int k = 0; k < j; ++k) s -= x[k] * Ls[j * kLDL + k];
                x[j] = s * inv[j];
            }
#pragma unroll
            for (int j = 0; j < kPanel; ++j)
                if (j < w) W[(size_t)(k0 + j) * ld + r] = x[j];
        }
        bar.sync();
        // the diagonal block itself (nobody reads it any more)
        if (cr == 0)
            for (int e = tid; e < w * w; e += kThreads) {
                const int j = e / w, r = e - j * w;
                if (r >= j) W[(size_t)(k0 + j) * ld + k0 + r] = Ls[r * kLDL + j];
            }
        // (c) trailing update of the lower triangle, 64 x 64 tiles dealt round-robin
        const int base = k0 + w;
        const int nrem = p - base;
        if (nrem > 0) {
            const int nt = (nrem + kTile - 1) / kTile;
            const int ntiles = nt * (nt + 1) / 2;
            const int tx = tid & 15, ty = tid >> 4;
            for (int t = cr; t < ntiles; t += T) {
                int ti = (int)((sqrt(8.0 * (double)t + 1.0) - 1.0) * 0.5);
                while (ti * (ti + 1) / 2 > t) --ti;
                while ((ti + 1) * (ti + 2) / 2 <= t) ++ti;
                const int tj = t - ti * (ti + 1) / 2;
                const int rb = base + kTile * ti, cb = base + kTile * tj;
                __syncthreads();
                for (int e = tid; e < kPanel * kTile; e += kThreads) {
                    const int k = e >> 6, i = e & 63;
                    const double *col = W + (size_t)(k0 + k) * ld;
                    Pr[k * kLDP + i] = (k < w && rb + i < p) ? __ldcg(col + rb + i) : 0.0;
                    Pc[k * kLDP + i] = (k < w && cb + i < p) ? __ldcg(col + cb + i) : 0.0;
                }
                __syncthreads();
                // 64 x 64 tile: thread (tx, ty) owns rows 4 tx .. 4 tx + 3 of the kTile / (kThreads / 16) columns from cpt * ty
                constexpr int cpt = kTile / (kThreads / 16);
                double acc[4][cpt];
#pragma unroll
                for (int i = 0; i < 4; ++i)
#pragma unroll
                    for (int j = 0; j < cpt; ++j) acc[i][j] = 0.0;
#pragma unroll 8
                for (int k = 0; k < kPanel; ++k) {
                    const double2 r01 = *reinterpret_cast<const double2 *>(Pr + k * kLDP + 4 * tx);
                    const double2 r23 = *reinterpret_cast<const double2 *>(Pr + k * kLDP + 4 * tx + 2);
                    const double rr[4] = {r01.x, r01.y, r23.x, r23.y};
#pragma unroll
                    for (int j = 0; j < cpt; ++j) {
                        const double cj = Pc[k * kLDP + cpt * ty + j];
#pragma unroll
                        for (int i = 0; i < 4; ++i) acc[i][j] = fma(rr[i], cj, acc[i][j]);
                    }
                }
                const int r = rb + 4 * tx;
                if (r < ld) {
#pragma unroll
                    for (int j = 0; j < cpt; ++j) {
                        const int c = cb + cpt * ty + j;
                        if (c >= p || r + 3 < c) continue;
                        double *dst = W + (size_t)c * ld + r;
                        double2 v0 = ldcg2(dst), v1 = ldcg2(dst + 2);
                        if (r >= c) v0.x -= acc[0][j];
                        if (r + 1 >= c) v0.y -= acc[1][j];
                        if (r + 2 >= c) v1.x -= acc[2][j];
                        v1.y -= acc[3][j];
                        *reinterpret_cast<double2 *>(dst) = v0;
                        *reinterpret_cast<double2 *>(dst + 2) = v1;
                    }
                }
            }
        }
        bar.sync();
    }
    return true;
}

__global__ void __launch_bounds__(kThreads, 1) eigb_kernel(const Params P)
{
    extern __shared__ __align__(16) double sm[];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int T = P.team;
    const int team_id = blockIdx.x / T, cr = blockIdx.x - team_id * T;
    TeamBarrier bar;
    bar.ctr = P.bar + team_id; bar.target = 0; bar.team = T;
    int *flags = P.flags + (size_t)team_id * kFlagStride;
    int *ver = P.ver + (size_t)team_id * P.ver_stride;
    __shared__ int s_job;

    for (;;) {
        if (cr == 0 && tid == 0) {
            const int jn = atomicAdd(P.next, 1);
            *reinterpret_cast<volatile int *>(P.slot + team_id) = jn;
            for (int s = 0; s < kFlagStride; ++s) *reinterpret_cast<volatile int *>(flags + s) = 0;
        }
        bar.sync();
        if (tid == 0) s_job = __ldcg(P.slot + team_id);
        __syncthreads();
        const int jn = s_job;
        if (jn >= P.n_jobs) break;
        const Job J = P.jobs[jn];
        const int p = J.p, ld = J.ld, nb = J.nb, ncols = nb * kB;
        double *W = P.W + J.w_off;
        const int32_t *idx = P.col_sets + J.set_off;
        const double tol = 2.220446049250313e-16 * (2.0 * sqrt((double)p) + 6.0);
        const double tol2 = tol * tol;

        // ---- symmetric permutation by decreasing diagonal (the cheap form of diagonal pivoting: the Cholesky factor of a
        // matrix whose diagonal decreases is graded, and one-sided Jacobi on a graded factor needs fewer sweeps --
        // Veselic / Hari); order[k] = position in idx of the k-th column, kept behind W ----------------------------------
        int32_t *order = reinterpret_cast<int32_t *>(W + (size_t)ld * ncols);
        if (P.sort_diag) {
            for (int c = cr * kWarps + warp; c < p; c += T * kWarps) {
                const double dc = P.G[(int64_t)idx[c] * P.ldg + idx[c]];
                int rank = 0;
                for (int q = lane; q < p; q += 32) {
                    const double dq = P.G[(int64_t)idx[q] * P.ldg + idx[q]];
                    if (dq > dc || (dq == dc && q < c)) ++rank;
                }
#pragma unroll
                for (int o = 16; o > 0; o >>= 1) rank += __shfl_xor_sync(0xffffffffu, rank, o);
                if (lane == 0) order[rank] = c;
            }
        } else {
            for (int c = cr * kThreads + tid; c < p; c += T * kThreads) order[c] = c;
        }
        for (int i = cr * kThreads + tid; i < nb; i += T * kThreads) ver[i] = 0;
        bar.sync();

        // ---- W <- lower triangle of A = G[idx][idx] (permuted), zero elsewhere (ld x ncols; G is symmetric) -------------------
        for (int c = cr * kWarps + warp; c < ncols; c += T * kWarps) {
            double *wc = W + (size_t)c * ld;
            if (c < p) {
                const double *grow = P.G + (int64_t)idx[__ldcg(order + c)] * P.ldg;
                for (int e = lane; e < ld; e += 32) wc[e] = (e >= c && e < p) ? grow[idx[__ldcg(order + e)]] : 0.0;
            } else {
                for (int e = lane; e < ld; e += 32) wc[e] = 0.0;
            }
        }
        bar.sync();

        const bool pd = team_cholesky(bar, W, p, ld, cr, T, sm);
        if (!pd) {
            if (cr == 0 && tid == 0) P.status[J.cand] = -1;
            continue;                                                    // the next job's first barrier re-aligns the team
        }

        // ---- sweeps: visits wait for their two producer visits only -----------------------------------------------------------
        int sweeps = 0, g = 0;
        for (; sweeps < kMaxSweeps; ++sweeps) {
            bool mine = false;
            for (int r = 0; r < nb - 1; ++r, ++g) {
                for (int k = cr; k < nb / 2; k += T) {
                    int bi, bj;
                    tournament_pair(nb, r, k, bi, bj);
                    if (bi > bj) { const int t = bi; bi = bj; bj = t; }
                    if (T > 1) {
                        if (tid == 0) {
                            while (ld_acquire(ver + bi) < g) { }
                            while (ld_acquire(ver + bj) < g) { }
                        }
                        __syncthreads();
                    }
                    mine |= visit_pair(W, ld, bi * kB, bj * kB, tol2, P.max_inner, sm);
                    __syncthreads();                                      // all stores of the visit issued; scratch reusable
                    if (T > 1 && tid == 0) {
                        __threadfence();
                        st_release(ver + bi, g + 1);
                        st_release(ver + bj, g + 1);
                    }
                }
            }
            if (mine && tid == 0) atomicOr(flags + sweeps, 1);
            bar.sync();
            if (tid == 0) s_job = __ldcg(flags + sweeps);
            __syncthreads();
            const int any = s_job;
            __syncthreads();
            if (!any) { ++sweeps; break; }
        }

        // ---- eigenvalues: lambda_j = |w_j|^2 --------------------------------------------------------------------------------
        double *lam_raw = P.lam_raw + J.vec_off + 64 * (int64_t)J.cand;
        for (int c = cr * kWarps + warp; c < p; c += T * kWarps) {
            const double *wc = W + (size_t)c * ld;
            double s = 0.0;
            for (int e = lane; e < p; e += 32) {
                const double v = __ldcg(wc + e);
                s = fma(v, v, s);
            }
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
            if (lane == 0) lam_raw[c] = s;
        }
        bar.sync();

        // ---- ascending order (like eigh), normalised eigenvectors ------------------------------------------------------------
        double *lamb = P.lamb + J.vec_off;
        double *Q = P.Q + J.mat_off;
        for (int c = cr * kWarps + warp; c < p; c += T * kWarps) {
            const double lj = __ldcg(lam_raw + c);
            int rank = 0;
            for (int q = lane; q < p; q += 32) {
                const double li = __ldcg(lam_raw + q);
                if (li < lj || (li == lj && q < c)) ++rank;
            }
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) rank += __shfl_xor_sync(0xffffffffu, rank, o);
            const double *wc = W + (size_t)c * ld;
            const double inv = 1.0 / sqrt(lj);
            double *qr = Q + (size_t)rank * p;
            for (int e = lane; e < p; e += 32) qr[__ldcg(order + e)] = __ldcg(wc + e) * inv;      // back to the model's column order
            if (lane == 0) lamb[rank] = lj;
        }
        if (cr == 0 && tid == 0) P.status[J.cand] = sweeps;
        // the next job's first barrier orders these reads of W before the team overwrites it
    }
}

}  // namespace eigb
