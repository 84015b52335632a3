// basis.cu -- K1: fused BSS-ANOVA design-matrix builder (sm_100a).
//
// Replaces the Python triple loop of the reference, src/FoKL/FoKLRoutines.py:1446-1485, together with
// _inputs_to_phind (FR:570-589) and evaluate_basis d = 0 (FR:834-836, 841-843).
//
// Per tile of ROWS consecutive data rows a CTA
//   phase 1: evaluates every distinct (input k, order d) factor the term list needs, once per row, into
//            shared memory F[factor][row] -- coefficient tables of the orders in use are staged in shared
//            memory (cubic: [slot][piece][4], one 32-byte gather per factor; Bernoulli: whole table);
//   phase 2: forms each term's product (increasing input index, like FR:1466-1483) from F and streams it to
//            the column-major design matrix with coalesced 16-byte stores (2 adjacent rows per thread).  Every term
//            is padded to NF factors with a row of ones (x * 1.0 is exact).  Consecutive terms of the reference's
//            lexicographic term order share their leading factors, so the running prefix product f0 (* f1) is kept
//            in registers and only the factors that changed are re-read from shared memory (the metadata word says
//            how many leading factors are unchanged): ~1.45 instead of 3 shared loads per 3-way term, same
//            left-to-right multiplication order, hence the same bits.  Term metadata sits in the kernel's
//            parameter space (constant bank), not in shared memory.
// Phase 1's coefficient gather is the other shared-memory hot spot: the staged table is [slot][coef][piece]
// (structure of arrays), so each of the 4 coefficient loads of a row is an 8-byte gather spread over all banks.
// HBM traffic = read N*M inputs once + write N*C outputs: 8*N*(M + C) bytes -- the kernel's roofline.
#include "fokl_ctx.cuh"
#include "fokl_math.cuh"
#include <algorithm>
#include <stdlib.h>
#include <string.h>
#include <type_traits>
#include <vector>

namespace {

constexpr int kThreads = 256;
constexpr int kMaxTermFactors = 7;     // interacting inputs per term (fit uses <= 3)
constexpr int kMaxFactors = 96;        // distinct (input, order) pairs per launch
constexpr int kMaxTermsPerLaunch = 2040;  // term metadata travels in the kernel parameter space (8 B each)
constexpr int kMaxPrefetchInputs = 8;  // distinct inputs of a launch whose rows phase 1 requests up front

struct FactorMeta {          // one distinct (input, order) pair
    int16_t k;               // input column
    int16_t d;               // order (1-based, phis[d-1])
    int16_t slot;            // staged-table slot of this order (cubic), -1 = read global table
    int16_t pad;             // direct path: the term index; derivative launches: derivative order e of this factor
};

struct alignas(8) TermMeta {  // one output column, read by the kernel as one packed 64-bit word:
    uint8_t f[8];             // f[0..6]: factor slots in increasing input order, padded with the "ones" slot (index
                              // n_factors); f[7]: how many leading slots equal the previous term's
};

struct BasisParams {
    const double *x;
    int64_t n, ldx;
    double *out;
    int64_t ld;
    const double *tab;       // global coefficient table
    int n_piece;             // cubic: pieces per order; bernoulli: row length
    int n_factors, n_terms, n_slots;
    const FactorMeta *factors;
    const int16_t *slot_order;   // order (1-based) staged in each slot
    int *flag;
    int64_t n_tiles;
    int tab_stride;              // cubic: doubles between the coefficient planes of a staged slot
    int direct;                  // every term is one factor (main effects): phase 1 stores straight to its column
                                 // (FactorMeta.pad = the term index), no factor buffer, no barrier, no phase 2
    const double *fac_div;       // derivative launches: per factor, the divisor span_L**e of FR:762-763 (1 for e = 0)
    int n_in;                    // distinct inputs in use if <= kMaxPrefetchInputs (phase 1 prefetch form), else 0
    int16_t in_k[kMaxPrefetchInputs];    // their input indices, ascending
    int16_t in_end[kMaxPrefetchInputs];  // one past the last factor (sorted by input) of each
    unsigned long long terms[kMaxTermsPerLaunch];   // TermMeta words
};
static_assert(sizeof(BasisParams) <= 32000, "kernel parameter space");

// Bernoulli with on-the-fly correctly rounded powers (no local array)
__device__ __forceinline__ double eval_factor_bernoulli(const double *c, int n_coef, double x)
{
    double h = x, l = 0.0, s = 0.0;
    for (int q = 1; q < n_coef; ++q) {
        if (q > 1) {
            double p = __dmul_rn(h, x);
            double e = __fma_rn(h, x, -p);
            double t = __fma_rn(l, x, e);
            double nh = __dadd_rn(p, t);
            l = __dsub_rn(t, __dsub_rn(nh, p));
            h = nh;
        }
        s = __dadd_rn(s, __dmul_rn(c[q], h));
    }
    return __dadd_rn(c[0], s);
}

// DERIV = true is the bss_derivatives form (FR:594-805): factors are evaluated at the twice-normalised input of
// FR:584-586 instead of xsm, and a factor with FactorMeta.pad = e > 0 is the e-th derivative of its basis function
// divided by fac_div (FR:780-781).  The fit path only ever instantiates DERIV = false.
// PF = true is the phase-1 form that requests the rows of every input in use up front (more registers: used for the
// main-effect launches and for launches whose shared-memory footprint allows at most 3 CTAs per SM anyway, where the
// exposed load latency of the input-by-input form is not covered by other CTAs; profiles/r01_k1_prefetch.txt).
template <int KERNEL, int RPT, int NF, bool DERIV = false, bool PF = false>
__global__ void __launch_bounds__(kThreads) basis_kernel(const __grid_constant__ BasisParams P)
{
    constexpr int ROWS = kThreads * RPT;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    // layout: [staged table][F: (n_factors + 1) * ROWS doubles, last row = ones][factors]
    double *s_tab = reinterpret_cast<double *>(smem_raw);
    size_t tab_doubles = (KERNEL == FOKL_KERNEL_CUBIC) ? (size_t)P.n_slots * P.tab_stride * 4
                                                       : (size_t)P.n_slots * P.n_piece;
    tab_doubles = (tab_doubles + 1) & ~(size_t)1;   // keep F 16-byte aligned
    double *s_F = s_tab + tab_doubles;
    FactorMeta *s_fac = reinterpret_cast<FactorMeta *>(s_F + ((NF == 1 && P.direct) ? (size_t)0 : (size_t)(P.n_factors + 1) * ROWS));

    const int tid = threadIdx.x;
    // ---- stage coefficient tables and metadata ------------------------------------------------
    if (KERNEL == FOKL_KERNEL_CUBIC) {
        // global [order][piece][4] -> shared [slot][4][tab_stride]
        const int per = P.n_piece * 4;
        for (int s = 0; s < P.n_slots; ++s) {
            const double *src = P.tab + (size_t)(P.slot_order[s] - 1) * per;
            double *dst = s_tab + (size_t)s * P.tab_stride * 4;
            for (int e = tid; e < per; e += kThreads) dst[(e & 3) * P.tab_stride + (e >> 2)] = __ldg(src + e);
        }
    } else {
        const int per = P.n_piece;
        for (int s = 0; s < P.n_slots; ++s) {
            const double *src = P.tab + (size_t)(P.slot_order[s] - 1) * per;
            for (int e = tid; e < per; e += kThreads) s_tab[(size_t)s * per + e] = __ldg(src + e);
        }
    }
    if (!(NF == 1 && P.direct))
        for (int e = tid; e < ROWS; e += kThreads) s_F[(size_t)P.n_factors * ROWS + e] = 1.0;
    for (int f = tid; f < P.n_factors; f += kThreads) s_fac[f] = P.factors[f];
    __syncthreads();

    // each CTA owns a contiguous range of row tiles: its C output streams stay inside the same few 2 MB pages
    // (one per column) for the whole launch instead of hopping pages every tile
    const int64_t per_cta = (P.n_tiles + gridDim.x - 1) / gridDim.x;
    const int64_t tile_lo = (int64_t)blockIdx.x * per_cta;
    const int64_t tile_hi = tile_lo + per_cta < P.n_tiles ? tile_lo + per_cta : P.n_tiles;
    for (int64_t tile = tile_lo; tile < tile_hi; ++tile) {
        const int64_t row0 = tile * ROWS + (int64_t)tid * RPT;
        // ---- phase 1: factor values -------------------------------------------------------------
        {
            // one factor's values for this thread's rows, to the factor buffer (or straight to its column)
            auto eval_store = [&](int f, const FactorMeta fm, const double (&xv)[RPT], const int (&ph)[RPT],
                                  const double (&xs)[RPT], const double (&x2)[RPT], const double (&x3)[RPT]) {
                double v[RPT];
                if (DERIV && fm.pad > 0) {
                    const double dv = __ldg(P.fac_div + f);
#pragma unroll
                    for (int r = 0; r < RPT; ++r) {
                        double t;
                        if (KERNEL == FOKL_KERNEL_CUBIC) {
                            const double *cp = P.tab + ((size_t)(fm.d - 1) * P.n_piece + ph[r]) * 4;
                            t = fm.pad == 1 ? fokl::cubic_basis_d1(__ldg(cp + 1), __ldg(cp + 2), __ldg(cp + 3), xs[r], x2[r])
                                            : fokl::cubic_basis_d2(__ldg(cp + 2), __ldg(cp + 3), xs[r]);
                        } else {
                            t = fokl::bernoulli_basis_deriv(P.tab + (size_t)(fm.d - 1) * P.n_piece, fm.d + 1, xv[r], fm.pad);
                        }
                        v[r] = __ddiv_rn(t, dv);
                    }
                } else if (KERNEL == FOKL_KERNEL_CUBIC) {
#pragma unroll
                    for (int r = 0; r < RPT; ++r) {
                        if (fm.slot >= 0) {
                            const double *cp = s_tab + (size_t)fm.slot * P.tab_stride * 4 + ph[r];
                            v[r] = fokl::cubic_basis(cp[0], cp[P.tab_stride], cp[2 * P.tab_stride], cp[3 * P.tab_stride],
                                                     xs[r], x2[r], x3[r]);
                        } else {
                            const double *cp = P.tab + ((size_t)(fm.d - 1) * P.n_piece + ph[r]) * 4;
                            v[r] = fokl::cubic_basis(__ldg(cp), __ldg(cp + 1), __ldg(cp + 2), __ldg(cp + 3), xs[r],
                                                     x2[r], x3[r]);
                        }
                    }
                } else {
                    const double *cp = (fm.slot >= 0) ? (s_tab + (size_t)fm.slot * P.n_piece)
                                                      : (P.tab + (size_t)(fm.d - 1) * P.n_piece);
#pragma unroll
                    for (int r = 0; r < RPT; ++r) v[r] = eval_factor_bernoulli(cp, fm.d + 1, xv[r]);
                }
                if (NF == 1 && P.direct) {
                    double *d = P.out + (int64_t)fm.pad * P.ld + row0;
                    if (RPT == 2 && row0 + 1 < P.n) {
                        __stcs(reinterpret_cast<double2 *>(d), make_double2(v[0], v[RPT - 1]));
                    } else {
#pragma unroll
                        for (int r = 0; r < RPT; ++r)
                            if (row0 + r < P.n) __stcs(d + r, v[r]);
                    }
                } else if (RPT == 2) {
                    *reinterpret_cast<double2 *>(s_F + (size_t)f * ROWS + tid * 2) = make_double2(v[0], v[RPT - 1]);
                } else {
                    s_F[(size_t)f * ROWS + tid] = v[0];
                }
            };
            auto load_rows = [&](int k, double (&xv)[RPT]) {
                const double *xc = P.x + (int64_t)k * P.ldx;
                if (RPT == 2 && row0 + 1 < P.n) {
                    double2 t = *reinterpret_cast<const double2 *>(xc + row0);
                    xv[0] = t.x;
                    xv[RPT - 1] = t.y;
                } else {
#pragma unroll
                    for (int r = 0; r < RPT; ++r) xv[r] = (row0 + r < P.n) ? xc[row0 + r] : 0.5;
                }
            };
            auto locate = [&](const double (&xv)[RPT], int (&ph)[RPT], double (&xs)[RPT], double (&x2)[RPT],
                              double (&x3)[RPT]) {
                if (KERNEL == FOKL_KERNEL_CUBIC) {
#pragma unroll
                    for (int r = 0; r < RPT; ++r) {
                        bool ok = fokl::phind_xsm(xv[r], P.n_piece, ph[r], xs[r]);
                        if (ph[r] > P.n_piece - 1) { ok = false; ph[r] = P.n_piece - 1; }
                        if (!ok && row0 + r < P.n) atomicOr(P.flag, 1);
                        if (DERIV) xs[r] = fokl::twice_normalised(xv[r], P.n_piece, ph[r]);
                        fokl::square_cube(xs[r], x2[r], x3[r]);
                    }
                }
            };
            if (PF) {
                // the rows of every input in use are requested back to back (one exposed HBM latency per tile instead
                // of one per input), then the factors are evaluated input by input (they are sorted by input)
                double xin[kMaxPrefetchInputs][RPT];
#pragma unroll
                for (int i = 0; i < kMaxPrefetchInputs; ++i)
                    if (i < P.n_in) load_rows(P.in_k[i], xin[i]);
                int f = 0;
#pragma unroll
                for (int i = 0; i < kMaxPrefetchInputs; ++i) {
                    if (i < P.n_in) {
                        double xs[RPT], x2[RPT], x3[RPT];
                        int ph[RPT];
                        locate(xin[i], ph, xs, x2, x3);
                        const int f_end = P.in_end[i];
                        for (; f < f_end; ++f) eval_store(f, s_fac[f], xin[i], ph, xs, x2, x3);
                    }
                }
            } else {
                double xv[RPT], xs[RPT], x2[RPT], x3[RPT];
                int ph[RPT];
                int cur_k = -1;
                for (int f = 0; f < P.n_factors; ++f) {
                    const FactorMeta fm = s_fac[f];
                    if (fm.k != cur_k) {
                        cur_k = fm.k;
                        load_rows(cur_k, xv);
                        locate(xv, ph, xs, x2, x3);
                    }
                    eval_store(f, fm, xv, ph, xs, x2, x3);
                }
            }
        }
        if (NF == 1 && P.direct) continue;
        __syncthreads();
        // ---- phase 2: products, streamed to HBM -----------------------------------------------------
        {
            constexpr int kShift = (RPT == 2) ? 12 : 11;       // log2(ROWS * sizeof(double))
            const bool full = (row0 + RPT - 1 < P.n);
            const unsigned char *Fb = reinterpret_cast<const unsigned char *>(s_F + tid * RPT);
            double *dst = P.out + row0;
            const int nt = P.n_terms;
            auto fload = [&](unsigned slot, double (&v)[RPT]) {
                const double *a = reinterpret_cast<const double *>(Fb + ((size_t)slot << kShift));
                if (RPT == 2) {
                    const double2 t2 = *reinterpret_cast<const double2 *>(a);
                    v[0] = t2.x; v[RPT - 1] = t2.y;
                } else {
                    v[0] = a[0];
                }
            };
            auto phase2 = [&](auto full_tag) {
            constexpr bool kFull = decltype(full_tag)::value;
            auto store = [&](double *d, const double (&v)[RPT]) {
                if (kFull) {
                    if (RPT == 2) __stcs(reinterpret_cast<double2 *>(d), make_double2(v[0], v[RPT - 1]));
                    else __stcs(d, v[0]);
                } else {
#pragma unroll
                    for (int r = 0; r < RPT; ++r)
                        if (row0 + r < P.n) __stcs(d + r, v[r]);
                }
            };
            if (NF <= 3) {
                // running prefix: pre[0] = f0, pre[1] = f0 * f1; `keep` leading factors are those of the previous term
                double f0[RPT], p01[RPT];
#pragma unroll
                for (int r = 0; r < RPT; ++r) f0[r] = p01[r] = 1.0;
#pragma unroll 4
                for (int j = 0; j < nt; ++j, dst += P.ld) {
                    const unsigned long long w = P.terms[j];
                    const unsigned keep = (unsigned)(w >> 56);
                    double v[RPT];
                    if (NF == 1) {
                        fload((unsigned)(w & 0xffu), v);
                    } else if (NF == 2) {
                        if (keep < 1u) fload((unsigned)(w & 0xffu), f0);
                        double t[RPT];
                        fload((unsigned)((w >> 8) & 0xffu), t);
#pragma unroll
                        for (int r = 0; r < RPT; ++r) v[r] = __dmul_rn(f0[r], t[r]);
                    } else {
                        if (keep < 1u) fload((unsigned)(w & 0xffu), f0);
                        if (keep < 2u) {
                            double t[RPT];
                            fload((unsigned)((w >> 8) & 0xffu), t);
#pragma unroll
                            for (int r = 0; r < RPT; ++r) p01[r] = __dmul_rn(f0[r], t[r]);
                        }
                        double t[RPT];
                        fload((unsigned)((w >> 16) & 0xffu), t);
#pragma unroll
                        for (int r = 0; r < RPT; ++r) v[r] = __dmul_rn(p01[r], t[r]);
                    }
                    store(dst, v);
                }
            } else {
                for (int j = 0; j < nt; ++j, dst += P.ld) {
                    const unsigned long long w = P.terms[j];
                    double v[RPT];
                    fload((unsigned)(w & 0xffu), v);
#pragma unroll
                    for (int q = 1; q < NF; ++q) {
                        double t[RPT];
                        fload((unsigned)((w >> (8 * q)) & 0xffu), t);
#pragma unroll
                        for (int r = 0; r < RPT; ++r) v[r] = __dmul_rn(v[r], t[r]);
                    }
                    store(dst, v);
                }
            }
            };
            if (full) phase2(std::true_type());
            else phase2(std::false_type());
        }
        __syncthreads();
    }
}

struct LaunchPlan {
    std::vector<FactorMeta> factors;
    std::vector<TermMeta> terms;
    std::vector<int16_t> slot_order;
    int first_term = 0;
    int max_cnt = 0;
};

}  // namespace

// deriv (C x M, values 0 / 1 / 2) and divisors (M x 3, host) are null for the fit path
static int basis_build_impl(fokl_ctx *ctx, int kernel, const double *x, int64_t n, int64_t ldx, int m,
                            const int16_t *terms, const uint8_t *deriv, const double *divisors, int c, double *Xnew,
                            int64_t ld)
{
    FOKL_CHECK_CTX(ctx);
    if (!x || !terms || !Xnew || n < 0 || m < 1 || c < 0 || ldx < n || ld < n)
        FOKL_FAIL(ctx, FOKL_EINVAL, "basis_build: bad argument");
    if (kernel != FOKL_KERNEL_CUBIC && kernel != FOKL_KERNEL_BERNOULLI)
        FOKL_FAIL(ctx, FOKL_EINVAL, "basis_build: unknown kernel");
    int rc = fokl_bind_device(ctx);
    if (rc) return rc;
    const double *tab = kernel == FOKL_KERNEL_CUBIC ? ctx->cubic_tab : ctx->bern_tab;
    const int n_orders = kernel == FOKL_KERNEL_CUBIC ? ctx->cubic_orders : ctx->bern_orders;
    const int row_len = kernel == FOKL_KERNEL_CUBIC ? ctx->cubic_pieces : ctx->bern_row;
    if (!tab) FOKL_FAIL(ctx, FOKL_ESTATE, "basis_build: phis table for this kernel was not set");
    if (n == 0 || c == 0) return FOKL_OK;

    // ---- plan launches: distinct (input, order) factors per chunk of terms ----------------------------
    std::vector<LaunchPlan> plans;
    {
        LaunchPlan cur;
        std::vector<int> fac_index((size_t)m * (n_orders + 1) * 3, -1);
        auto flush = [&]() {
            if (!cur.terms.empty()) plans.push_back(cur);
            int next = cur.first_term + (int)cur.terms.size();
            cur = LaunchPlan();
            cur.first_term = next;
            std::fill(fac_index.begin(), fac_index.end(), -1);
        };
        for (int j = 0; j < c; ++j) {
            const int16_t *row = terms + (size_t)j * m;
            int need_new = 0, cnt = 0;
            for (int k = 0; k < m; ++k) {
                int d = row[k];
                if (d == 0) continue;
                if (d < 0 || d > n_orders) FOKL_FAIL(ctx, FOKL_EINVAL, "basis_build: term order outside len(phis)");
                if (kernel == FOKL_KERNEL_BERNOULLI && d + 1 > row_len)
                    FOKL_FAIL(ctx, FOKL_EINVAL, "basis_build: bernoulli row too short for order");
                ++cnt;
                const int e = deriv ? deriv[(size_t)j * m + k] : 0;
                if (e < 0 || e > 2) FOKL_FAIL(ctx, FOKL_EINVAL, "basis_build: derivative order must be 0, 1 or 2");
                if (fac_index[((size_t)k * (n_orders + 1) + d) * 3 + e] < 0) ++need_new;
            }
            if (cnt > kMaxTermFactors) FOKL_FAIL(ctx, FOKL_EINVAL, "basis_build: more than 7 interacting inputs in a term");
            if ((int)cur.factors.size() + need_new > kMaxFactors || (int)cur.terms.size() >= kMaxTermsPerLaunch) flush();
            TermMeta tm;
            int filled = 0;
            for (int q = 0; q < 8; ++q) tm.f[q] = 0xff;           // 0xff = "ones" slot, patched per launch below
            for (int k = 0; k < m; ++k) {
                int d = row[k];
                if (d == 0) continue;
                const int e = deriv ? deriv[(size_t)j * m + k] : 0;
                int &fi = fac_index[((size_t)k * (n_orders + 1) + d) * 3 + e];
                if (fi < 0) {
                    fi = (int)cur.factors.size();
                    FactorMeta fm;
                    fm.k = (int16_t)k; fm.d = (int16_t)d; fm.slot = -1; fm.pad = (int16_t)e;
                    cur.factors.push_back(fm);
                }
                tm.f[filled++] = (uint8_t)fi;
            }
            cur.max_cnt = std::max(cur.max_cnt, filled);
            cur.terms.push_back(tm);
        }
        flush();
    }

    const bool aligned2 = ((ld % 2) == 0) && ((ldx % 2) == 0) && (((uintptr_t)Xnew % 16) == 0) && (((uintptr_t)x % 16) == 0);
    const size_t smem_cap = ctx->smem_optin ? ctx->smem_optin : 48 * 1024;

    for (LaunchPlan &pl : plans) {
        // phase 1 walks factors grouped by input: sort by (k, d) and remap term indices
        const int nf = (int)pl.factors.size();
        std::vector<int> order(nf), inv(nf);
        for (int i = 0; i < nf; ++i) order[i] = i;
        std::sort(order.begin(), order.end(), [&](int a, int b) {
            if (pl.factors[a].k != pl.factors[b].k) return pl.factors[a].k < pl.factors[b].k;
            if (pl.factors[a].d != pl.factors[b].d) return pl.factors[a].d < pl.factors[b].d;
            return pl.factors[a].pad < pl.factors[b].pad;
        });
        std::vector<FactorMeta> sorted(nf);
        for (int i = 0; i < nf; ++i) { sorted[i] = pl.factors[order[i]]; inv[order[i]] = i; }
        for (TermMeta &tm : pl.terms)
            for (int q = 0; q < 7; ++q) tm.f[q] = (tm.f[q] == 0xff) ? (uint8_t)nf : (uint8_t)inv[tm.f[q]];
        // f[7]: length of the factor prefix shared with the previous term (phase 2 keeps that prefix product)
        for (size_t j = 0; j < pl.terms.size(); ++j) {
            int keep = 0;
            if (j > 0)
                while (keep < 7 && pl.terms[j].f[keep] == pl.terms[j - 1].f[keep]) ++keep;
            pl.terms[j].f[7] = (uint8_t)keep;
        }
        // staged-table slots: distinct orders, as many as fit next to F
        std::vector<int16_t> orders;
        for (const FactorMeta &fm : sorted)
            if (std::find(orders.begin(), orders.end(), fm.d) == orders.end()) orders.push_back(fm.d);
        std::sort(orders.begin(), orders.end());
        const int tab_stride = row_len + 1;     // cubic coefficient planes: any stride works for the 8-byte gathers
        const size_t per_slot = (kernel == FOKL_KERNEL_CUBIC ? (size_t)tab_stride * 4 : (size_t)row_len) * sizeof(double);
        const int nt = (int)pl.terms.size();
        // main-effect substages: every term is exactly one factor and no factor is shared -> direct stores
        bool direct = pl.max_cnt == 1 && nt == nf && !deriv;
        if (direct) {
            std::vector<int> seen(nf, 0);
            for (const TermMeta &tm : pl.terms) {
                if (tm.f[0] >= nf || seen[tm.f[0]]++) { direct = false; break; }
            }
        }
        if (direct)
            for (int j = 0; j < nt; ++j) sorted[pl.terms[j].f[0]].pad = (int16_t)j;
        auto smem_need = [&](int rpt, int nslots) {
            size_t tab_bytes = ((nslots * per_slot / sizeof(double) + 1) & ~(size_t)1) * sizeof(double);
            return tab_bytes + (direct ? (size_t)0 : (size_t)(nf + 1) * kThreads * rpt * sizeof(double)) +
                   (size_t)nf * sizeof(FactorMeta) + 16;
        };
        int rpt = (aligned2 && !deriv) ? 2 : 1;
        int nslots = (int)orders.size();
        // prefer two resident CTAs per SM; otherwise shrink rows per thread, then staged slots
        if (rpt == 2 && smem_need(2, nslots) > smem_cap / 2) rpt = 1;
        while (nslots > 0 && smem_need(rpt, nslots) > smem_cap) --nslots;
        if (smem_need(rpt, nslots) > smem_cap) FOKL_FAIL(ctx, FOKL_EINVAL, "basis_build: term list needs too much shared memory");
        orders.resize(nslots);
        for (FactorMeta &fm : sorted) {
            auto it = std::find(orders.begin(), orders.end(), fm.d);
            fm.slot = (it == orders.end()) ? (int16_t)-1 : (int16_t)(it - orders.begin());
        }
        // upload metadata
        size_t off_fac = 0;
        size_t off_slot = (off_fac + (size_t)nf * sizeof(FactorMeta) + 15) & ~(size_t)15;
        size_t off_div = (off_slot + (size_t)(nslots + 1) * sizeof(int16_t) + 15) & ~(size_t)15;
        size_t meta_bytes = off_div + (deriv ? (size_t)nf * sizeof(double) : 0);
        std::vector<unsigned char> host(meta_bytes, 0);
        memcpy(host.data() + off_fac, sorted.data(), (size_t)nf * sizeof(FactorMeta));
        if (deriv) {
            double *dv = reinterpret_cast<double *>(host.data() + off_div);
            for (int i = 0; i < nf; ++i) dv[i] = sorted[i].pad > 0 ? divisors[(size_t)sorted[i].k * 3 + sorted[i].pad] : 1.0;
        }
        if (nslots) memcpy(host.data() + off_slot, orders.data(), (size_t)nslots * sizeof(int16_t));
        // the copy is stream-ordered behind any kernel still reading the previous metadata
        unsigned char *dmeta = (unsigned char *)fokl_scratch(ctx, fokl_ctx::B_BASIS, meta_bytes);
        if (!dmeta) return FOKL_ENOMEM;
        FOKL_CUDA(ctx, cudaMemcpyAsync(dmeta, host.data(), meta_bytes, cudaMemcpyHostToDevice, ctx->stream));

        BasisParams P;
        P.x = x; P.n = n; P.ldx = ldx;
        P.out = Xnew + (int64_t)pl.first_term * ld; P.ld = ld;
        P.tab = tab; P.n_piece = row_len;
        P.n_factors = nf; P.n_terms = nt; P.n_slots = nslots;
        P.factors = reinterpret_cast<const FactorMeta *>(dmeta + off_fac);
        P.tab_stride = tab_stride;
        P.direct = direct ? 1 : 0;
        P.fac_div = deriv ? reinterpret_cast<const double *>(dmeta + off_div) : nullptr;
        {
            int n_in = 0;
            bool fits = true;
            for (int i = 0; i < nf && fits; ++i) {
                if (n_in == 0 || sorted[i].k != P.in_k[n_in - 1]) {
                    if (n_in == kMaxPrefetchInputs) { fits = false; break; }
                    P.in_k[n_in++] = sorted[i].k;
                }
                P.in_end[n_in - 1] = (int16_t)(i + 1);
            }
            for (int i = fits ? n_in : 0; i < kMaxPrefetchInputs; ++i) { P.in_k[i] = 0; P.in_end[i] = 0; }
            P.n_in = fits ? n_in : 0;
        }
        memcpy(P.terms, pl.terms.data(), (size_t)nt * sizeof(TermMeta));
        P.slot_order = reinterpret_cast<const int16_t *>(dmeta + off_slot);
        P.flag = ctx->d_flag;
        const int rows = kThreads * rpt;
        P.n_tiles = (n + rows - 1) / rows;
        const size_t smem = smem_need(rpt, nslots);
        const int ctas_per_sm = (int)std::max<size_t>(1, std::min<size_t>(4, smem_cap / std::max<size_t>(smem, 1)));
        int grid = (int)std::min<int64_t>(P.n_tiles, (int64_t)ctx->num_sms * ctas_per_sm);
        void (*kern)(const BasisParams) = nullptr;
        const bool pf = P.n_in > 0 && !deriv && !getenv("FOKL_BASIS_NOPF") &&
                        (direct || smem_cap / std::max<size_t>(smem, 1) <= 3);
        static_assert(sizeof(TermMeta) == sizeof(unsigned long long), "TermMeta is one 64-bit word");
        const int nfc = pl.max_cnt <= 1 ? 1 : (pl.max_cnt == 2 ? 2 : (pl.max_cnt == 3 ? 3 : 7));
#define FOKL_PICK(K, R, F)                                                                                            \
    (nfc == 1 ? basis_kernel<K, R, 1, false, F> : nfc == 2 ? basis_kernel<K, R, 2, false, F>                          \
                                                : nfc == 3 ? basis_kernel<K, R, 3, false, F> : basis_kernel<K, R, 7, false, F>)
        if (kernel == FOKL_KERNEL_CUBIC) {
            if (pf) kern = rpt == 2 ? FOKL_PICK(FOKL_KERNEL_CUBIC, 2, true) : FOKL_PICK(FOKL_KERNEL_CUBIC, 1, true);
            else kern = rpt == 2 ? FOKL_PICK(FOKL_KERNEL_CUBIC, 2, false) : FOKL_PICK(FOKL_KERNEL_CUBIC, 1, false);
        } else {
            if (pf) kern = rpt == 2 ? FOKL_PICK(FOKL_KERNEL_BERNOULLI, 2, true) : FOKL_PICK(FOKL_KERNEL_BERNOULLI, 1, true);
            else kern = rpt == 2 ? FOKL_PICK(FOKL_KERNEL_BERNOULLI, 2, false) : FOKL_PICK(FOKL_KERNEL_BERNOULLI, 1, false);
        }
#undef FOKL_PICK
        if (deriv)
            kern = kernel == FOKL_KERNEL_CUBIC ? basis_kernel<FOKL_KERNEL_CUBIC, 1, 7, true>
                                               : basis_kernel<FOKL_KERNEL_BERNOULLI, 1, 7, true>;
        FOKL_CUDA(ctx, cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_cap));
        kern<<<grid, kThreads, smem, ctx->stream>>>(P);
        FOKL_LAUNCH_CHECK(ctx);
    }
    return FOKL_OK;
}

extern "C" int fokl_basis_build(fokl_ctx *ctx, int kernel, const double *x, int64_t n, int64_t ldx, int m,
                                const int16_t *terms, int c, double *Xnew, int64_t ld)
{
    return basis_build_impl(ctx, kernel, x, n, ldx, m, terms, nullptr, nullptr, c, Xnew, ld);
}

// bss_derivatives form of K1 (FR:594-805): column j = prod_k f_jk with f_jk the basis function of order terms[j][k]
// (deriv[j][k] = 0) or its first / second derivative divided by divisors[k][deriv[j][k]] (FR:780-781), every factor
// evaluated at the twice-normalised input of FR:584-586.  deriv and divisors are host arrays.
extern "C" int fokl_basis_build_deriv(fokl_ctx *ctx, int kernel, const double *x, int64_t n, int64_t ldx, int m,
                                      const int16_t *terms, const uint8_t *deriv, const double *divisors, int c,
                                      double *Xnew, int64_t ld)
{
    FOKL_CHECK_CTX(ctx);
    if (!deriv || !divisors) FOKL_FAIL(ctx, FOKL_EINVAL, "basis_build_deriv: bad argument");
    return basis_build_impl(ctx, kernel, x, n, ldx, m, terms, deriv, divisors, c, Xnew, ld);
}
