// cand_math.cuh -- per-candidate spectral factorisation, betahat/BIC and the eigenbasis Gibbs chain.
//
// Written against a tiny "Team" abstraction (one CTA on the device; a single sequential thread in the
// host emulation build, tests/host_emu) so the numerical logic can be checked against the CPU oracle
// without a GPU.  Reference stages replaced (src/FoKL/FoKLRoutines.py):
//   eigh(XtX)              FR:1499     -> jacobi_eigh   (one-sided cyclic Jacobi, round-robin pairs)
//   betahat                FR:1502-04  -> ols_and_bic
//   BIC                    FR:1551-54  -> ols_and_bic   (from Gram quantities, centred on mean(y))
//   draw loop              FR:1519-48  -> gibbs_chain   (eigenbasis form, O(p) per draw)
#pragma once
#include "fokl_math.cuh"

namespace fokl {

struct Team {
    int tid, nthr;    // thread in CTA
    int lane, nlane;  // lane in warp (nlane = 1 in host emulation)
    int warp, nwarp;
    FOKL_HD void sync() const
    {
#if defined(__CUDA_ARCH__)
        __syncthreads();
#endif
    }
};

FOKL_HD void warp_sum3(const Team &, double &a, double &b, double &c)
{
#if defined(__CUDA_ARCH__)
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        a += __shfl_xor_sync(0xffffffffu, a, o);
        b += __shfl_xor_sync(0xffffffffu, b, o);
        c += __shfl_xor_sync(0xffffffffu, c, o);
    }
#else
    (void)a; (void)b; (void)c;
#endif
}

FOKL_HD double warp_sum1(const Team &, double a)
{
#if defined(__CUDA_ARCH__)
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) a += __shfl_xor_sync(0xffffffffu, a, o);
#endif
    return a;
}

// CTA-wide sum of three values; every thread returns the same bits (fixed summation order).
// red: shared scratch of 2 * 3 * nwarp doubles; parity alternates between calls.
FOKL_HD void team_sum3(const Team &t, double &a, double &b, double &c, double *red, int parity)
{
    warp_sum3(t, a, b, c);
    double *r = red + parity * 3 * t.nwarp;
    if (t.lane == 0) {
        r[3 * t.warp + 0] = a;
        r[3 * t.warp + 1] = b;
        r[3 * t.warp + 2] = c;
    }
    t.sync();
    double sa = 0.0, sb = 0.0, sc = 0.0;
    for (int w = 0; w < t.nwarp; ++w) {
        sa += r[3 * w + 0];
        sb += r[3 * w + 1];
        sc += r[3 * w + 2];
    }
    a = sa; b = sb; c = sc;
}

// ---- one-sided Jacobi ---------------------------------------------------------------------------
// Orthogonalise columns i, j of W (and apply the same rotation to V).  One warp per pair.
FOKL_HD bool jacobi_pair(const Team &t, double *wi, double *wj, double *vi, double *vj, int p, double tol)
{
    double al = 0.0, be = 0.0, ga = 0.0;
    for (int e = t.lane; e < p; e += t.nlane) {
        double a = wi[e], b = wj[e];
        al += a * a;
        be += b * b;
        ga += a * b;
    }
    warp_sum3(t, al, be, ga);
    if (!(fabs(ga) > tol * sqrt(al * be))) return false;
    double zeta = (be - al) / (2.0 * ga);
    double tt = copysign(1.0, zeta) / (fabs(zeta) + sqrt(1.0 + zeta * zeta));
    double c = 1.0 / sqrt(1.0 + tt * tt);
    double s = c * tt;
    for (int e = t.lane; e < p; e += t.nlane) {
        double a = wi[e], b = wj[e];
        wi[e] = c * a - s * b;
        wj[e] = s * a + c * b;
        a = vi[e]; b = vj[e];
        vi[e] = c * a - s * b;
        vj[e] = s * a + c * b;
    }
    return true;
}

// W (in: symmetric matrix, column-major ld; out: W = G V with mutually orthogonal columns),
// V (out: orthogonal).  flag: one shared int.  Returns the number of sweeps used.
FOKL_HD int jacobi_eigh(const Team &t, double *W, double *V, int p, int ld, int max_sweeps, double tol,
                        volatile int *flag)
{
    for (int j = t.warp; j < p; j += t.nwarp)
        for (int e = t.lane; e < p; e += t.nlane) V[(int64_t)j * ld + e] = (e == j) ? 1.0 : 0.0;
    t.sync();
    if (p < 2) return 0;
    const int n = p + (p & 1);
    const int nm1 = n - 1;
    int sweep = 0;
    for (; sweep < max_sweeps; ++sweep) {
        if (t.tid == 0) *flag = 0;
        t.sync();
        for (int r = 0; r < nm1; ++r) {
            for (int k = t.warp; k < n / 2; k += t.nwarp) {
                int i, j;
                if (k == 0) { i = nm1; j = r; }
                else { i = (r + k) % nm1; j = (r - k + nm1) % nm1; }
                if (i > j) { int q = i; i = j; j = q; }
                if (j >= p) continue;
                bool rot = jacobi_pair(t, W + (int64_t)i * ld, W + (int64_t)j * ld, V + (int64_t)i * ld,
                                       V + (int64_t)j * ld, p, tol);
                if (rot && t.lane == 0) *flag = 1;
            }
            t.sync();
        }
        int any = *flag;
        t.sync();
        if (!any) { ++sweep; break; }
    }
    return sweep;
}

// After jacobi_eigh: Rayleigh quotients, ascending order (like scipy.linalg.eigh), eigenvectors.
//   lam_raw (p scratch), perm (p ints scratch); outputs lamb[p] ascending, Q column-major p x p (ld = p).
FOKL_HD void eig_finish(const Team &t, const double *W, const double *V, int p, int ld, double *lam_raw,
                        int *perm, double *lamb, double *Q)
{
    for (int j = t.warp; j < p; j += t.nwarp) {
        double s = 0.0;
        for (int e = t.lane; e < p; e += t.nlane) s += V[(int64_t)j * ld + e] * W[(int64_t)j * ld + e];
        s = warp_sum1(t, s);
        if (t.lane == 0) lam_raw[j] = s;
    }
    t.sync();
    for (int j = t.tid; j < p; j += t.nthr) {
        double lj = lam_raw[j];
        int rank = 0;
        for (int i = 0; i < p; ++i) {
            double li = lam_raw[i];
            rank += (li < lj || (li == lj && i < j)) ? 1 : 0;
        }
        perm[rank] = j;
    }
    t.sync();
    for (int r = t.warp; r < p; r += t.nwarp) {
        int j = perm[r];
        for (int e = t.lane; e < p; e += t.nlane) Q[(int64_t)r * p + e] = V[(int64_t)j * ld + e];
        if (t.lane == 0) lamb[r] = lam_raw[j];
    }
    t.sync();
}

// ---- Cholesky + one-sided Jacobi on the factor (Veselic-Hari) ----------------------------------------
// G = L L'.  Rotating the columns of W = L until they are mutually orthogonal gives W = U diag(sqrt(lambda)):
// eigenvectors u_j = w_j / |w_j| and eigenvalues |w_j|^2 of G, with no accumulated rotation matrix (half the
// memory and flops of jacobi_eigh) and high relative accuracy for the small eigenvalues.

// Right-looking Cholesky of the lower triangle stored column-major in L (ld = p): on return column j holds
// L[i][j], i >= j, and zeros above the diagonal.  Returns false if a pivot is not positive.
// Packed form: column j of the lower triangle holds its p - j entries (j .. p-1) at chol_col(j, p); half the footprint of
// the square form, so the factor of a 226-column model is computed in shared memory.  One warp per trailing column,
// lanes along the rows (contiguous), no index arithmetic per element.
FOKL_HD int64_t chol_col(int j, int p) { return (int64_t)j * p - (int64_t)j * (j - 1) / 2; }

FOKL_HD bool cholesky_lower_packed(const Team &t, double *L, int p)
{
    bool ok = true;
    for (int j = 0; j < p; ++j) {
        double *Lj = L + chol_col(j, p) - j;                     // Lj[i] = L(i, j), i >= j
        const double d = Lj[j];
        if (!(d > 0.0) || !(d < 1.7e308)) { ok = false; break; }
        const double r = sqrt(d);
        t.sync();
        for (int i = j + t.tid; i < p; i += t.nthr) Lj[i] = (i == j) ? r : Lj[i] / r;
        t.sync();
        for (int m = j + 1 + t.warp; m < p; m += t.nwarp) {
            double *Lm = L + chol_col(m, p) - m;
            const double lmj = Lj[m];
            for (int i = m + t.lane; i < p; i += t.nlane) Lm[i] -= Lj[i] * lmj;
        }
        t.sync();
    }
    t.sync();
    return ok;
}

FOKL_HD bool cholesky_lower(const Team &t, double *L, int p)
{
    bool ok = true;
    for (int j = 0; j < p; ++j) {
        const double d = L[(int64_t)j * p + j];
        if (!(d > 0.0) || !(d < 1.7e308)) { ok = false; break; }
        const double r = sqrt(d);
        t.sync();
        for (int i = j + t.tid; i < p; i += t.nthr) L[(int64_t)j * p + i] = (i == j) ? r : L[(int64_t)j * p + i] / r;
        t.sync();
        const int rem = p - j - 1;
        for (int64_t e = t.tid; e < (int64_t)rem * rem; e += t.nthr) {
            int mc = (int)(e / rem), mr = (int)(e - (int64_t)mc * rem);
            if (mr < mc) continue;
            int m = j + 1 + mc, i = j + 1 + mr;
            L[(int64_t)m * p + i] -= L[(int64_t)j * p + i] * L[(int64_t)j * p + m];
        }
        t.sync();
    }
    t.sync();
    return ok;
}

// Orthogonalise columns wi, wj (length p); the rotated columns go to wi_out / wj_out (either may alias its input;
// a column whose output differs from its input is copied there even when no rotation is needed, together with the
// `extra` trailing elements that travel with it).  One warp per pair.
FOKL_HD bool jacobi_pair_w(const Team &t, const double *wi, const double *wj, double *wi_out, double *wj_out, int p,
                           int extra, double tol)
{
    double al = 0.0, be = 0.0, ga = 0.0;
    for (int e = t.lane; e < p; e += t.nlane) {
        double a = wi[e], b = wj[e];
        al += a * a;
        be += b * b;
        ga += a * b;
    }
    warp_sum3(t, al, be, ga);
    const bool rot = fabs(ga) > tol * sqrt(al * be);
    if (rot) {
        // tan(theta) = 2 ga / (d + sign(d) sqrt(d^2 + 4 ga^2)), d = |wj|^2 - |wi|^2 (the smaller root), one sqrt, one
        // division, one reciprocal square root on the critical path
        const double d = be - al;
        const double r = sqrt(d * d + 4.0 * ga * ga);
        const double tt = (2.0 * ga) / (d + copysign(r, d));
#if defined(__CUDA_ARCH__)
        const double c = rsqrt(1.0 + tt * tt);
#else
        const double c = 1.0 / sqrt(1.0 + tt * tt);
#endif
        const double s = c * tt;
        for (int e = t.lane; e < p; e += t.nlane) {
            double a = wi[e], b = wj[e];
            wi_out[e] = c * a - s * b;
            wj_out[e] = s * a + c * b;
        }
        if (wi_out != wi)
            for (int e = p + t.lane; e < p + extra; e += t.nlane) wi_out[e] = wi[e];
        if (wj_out != wj)
            for (int e = p + t.lane; e < p + extra; e += t.nlane) wj_out[e] = wj[e];
    } else {
        if (wi_out != wi)
            for (int e = t.lane; e < p + extra; e += t.nlane) wi_out[e] = wi[e];
        if (wj_out != wj)
            for (int e = t.lane; e < p + extra; e += t.nlane) wj_out[e] = wj[e];
    }
    return rot;
}

// Single-team driver (host emulation and reference for the cluster kernel): W (in: L, column-major ld) is rotated
// in place.  Returns the number of sweeps.
FOKL_HD int jacobi_w_sweeps(const Team &t, double *W, int p, int ld, int max_sweeps, double tol, volatile int *flag)
{
    if (p < 2) return 0;
    const int n = p + (p & 1);
    const int nm1 = n - 1;
    int sweep = 0;
    for (; sweep < max_sweeps; ++sweep) {
        if (t.tid == 0) *flag = 0;
        t.sync();
        for (int r = 0; r < nm1; ++r) {
            for (int k = t.warp; k < n / 2; k += t.nwarp) {
                int i, j;
                if (k == 0) { i = nm1; j = r; }
                else { i = (r + k) % nm1; j = (r - k + nm1) % nm1; }
                if (i > j) { int q = i; i = j; j = q; }
                if (j >= p) continue;
                double *wi = W + (int64_t)i * ld, *wj = W + (int64_t)j * ld;
                bool rot = jacobi_pair_w(t, wi, wj, wi, wj, p, 0, tol);
                if (rot && t.lane == 0) *flag = 1;
            }
            t.sync();
        }
        int any = *flag;
        t.sync();
        if (!any) { ++sweep; break; }
    }
    return sweep;
}

// After jacobi_w_sweeps: eigenvalues = squared column norms, ascending; Q row r = normalised column (ld_q = p).
FOKL_HD void eig_finish_w(const Team &t, const double *W, int p, int ld, double *lam_raw, int *perm, double *lamb,
                          double *Q)
{
    for (int j = t.warp; j < p; j += t.nwarp) {
        double s = 0.0;
        for (int e = t.lane; e < p; e += t.nlane) s += W[(int64_t)j * ld + e] * W[(int64_t)j * ld + e];
        s = warp_sum1(t, s);
        if (t.lane == 0) lam_raw[j] = s;
    }
    t.sync();
    for (int j = t.tid; j < p; j += t.nthr) {
        double lj = lam_raw[j];
        int rank = 0;
        for (int i = 0; i < p; ++i) {
            double li = lam_raw[i];
            rank += (li < lj || (li == lj && i < j)) ? 1 : 0;
        }
        perm[rank] = j;
    }
    t.sync();
    for (int r = t.warp; r < p; r += t.nwarp) {
        int j = perm[r];
        const double inv = 1.0 / sqrt(lam_raw[j]);
        for (int e = t.lane; e < p; e += t.nlane) Q[(int64_t)r * p + e] = W[(int64_t)j * ld + e] * inv;
        if (t.lane == 0) lamb[r] = lam_raw[j];
    }
    t.sync();
}

struct CandConst {
    double a, b, atau, btau, sigsqd0, tausqd0, yty, sum_y;
    double n;     // number of rows as double
    int draws, from0, from1;
};

// betahat = Q diag(1/lamb) Q' Xty (FR:1502-1504, no clamp on tiny/negative eigenvalues) and the BIC of
// FR:1551-1554 computed from Gram quantities.  G (row-major, ldg) is the master Gram, idx the candidate's
// column list (idx[0] must be the intercept column).  ct[p] (out) = Q' Xty.  scratch: p doubles.
// Returns ev in every thread.
FOKL_HD double ols_and_bic(const Team &t, const double *G, int64_t ldg, const double *Xty, const int *idx,
                           int p, const double *lamb, const double *Q, const CandConst &k, double *ct,
                           double *betahat, double *scratch, double *red)
{
    for (int r = t.warp; r < p; r += t.nwarp) {
        double s = 0.0;
        for (int e = t.lane; e < p; e += t.nlane) s += Q[(int64_t)r * p + e] * Xty[idx[e]];
        s = warp_sum1(t, s);
        if (t.lane == 0) ct[r] = s;
    }
    t.sync();
    for (int i = t.tid; i < p; i += t.nthr) {
        double s = 0.0;
        for (int r = 0; r < p; ++r) s += Q[(int64_t)r * p + i] * (ct[r] / lamb[r]);
        betahat[i] = s;
    }
    t.sync();
    // centred quantities: y_c = y - ybar, beta_c = betahat - ybar e_0, r = y_c - X beta_c
    const double ybar = k.sum_y / k.n;
    const int64_t row0 = (int64_t)idx[0] * ldg;    // G[0][*] = column sums
    double d1 = 0.0, d2 = 0.0, d3 = 0.0;           // bc'xc, bc'G bc, g0'bc
    for (int i = t.tid; i < p; i += t.nthr) {
        double bi = betahat[i] - (i == 0 ? ybar : 0.0);
        scratch[i] = bi;
    }
    t.sync();
    for (int i = t.tid; i < p; i += t.nthr) {
        const double *gi = G + (int64_t)idx[i] * ldg;
        double s = 0.0;
        for (int j = 0; j < p; ++j) s += gi[idx[j]] * scratch[j];
        double g0 = G[row0 + idx[i]];
        double xc = Xty[idx[i]] - ybar * g0;
        d1 += scratch[i] * xc;
        d2 += scratch[i] * s;
        d3 += g0 * scratch[i];
    }
    team_sum3(t, d1, d2, d3, red, 0);
    t.sync();
    double yty_c = k.yty - k.n * ybar * ybar;
    double srr = yty_c - 2.0 * d1 + d2;
    double sr = -d3;
    double siglik = srr / k.n - (sr / k.n) * (sr / k.n);
    double lik = -(k.n / 2.0) * log(siglik) - (k.n - 1.0) / 2.0;
    return (double)p * log(k.n) - 2.0 * lik;
}

// One variate of the free-running stream of a `gibbs` call: row d of the chain consumes normal(p), then
// standard_gamma(astar), standard_gamma(atau_star) -- the order of FR:1527, 1541, 1547.  Counter-based, so the
// D x (p + 2) table can be filled by any number of threads in any order (cand_variates_kernel) and the chain
// itself only streams it.
FOKL_HD double philox_variate(const Philox &g, uint32_t stream_lo, uint32_t stream_hi, int d, int e, int p,
                              double astar, double atau_star)
{
    if (e < p) return philox_normal(g, stream_lo, stream_hi, (uint32_t)d, (uint32_t)e);
    if (e == p) return philox_gamma(g, stream_lo, stream_hi, (uint32_t)d, 0u, astar);
    return philox_gamma(g, stream_lo, stream_hi, (uint32_t)d, 1024u, atau_star);
}

FOKL_HD double chain_astar(const CandConst &k, int p) { return k.a + 1.0 + k.n / 2.0 + (double)p / 2.0; }
FOKL_HD double chain_atau_star(const CandConst &k, int p) { return k.atau + (double)(p - 1) / 2.0; }

#if defined(__CUDA_ARCH__)
#define FOKL_RCP(x) __drcp_rn(x)
#define FOKL_RSQRT(x) rsqrt(x)
#else
#define FOKL_RCP(x) (1.0 / (x))
#define FOKL_RSQRT(x) (1.0 / sqrt(x))
#endif

// The draw loop of FR:1519-1548 in the eigenbasis (gamma = Q' beta):
//   d_j = 1/(lamb_j + 1/tau^2);  gamma_j = d_j ct_j + sqrt(sig^2) sqrt(d_j) z_j
//   bstar = b + 0.5 (sum lamb gamma^2 - 2 sum gamma ct + yty + sum gamma^2 / tau^2)
//   sig^2 = 1 / ((1/bstar) G1);  btau* = (1/(2 sig^2)) sum gamma^2 + btau;  tau^2 = 1 / ((1/btau*) G2)
// The D draws are strictly sequential, so the loop is arranged around its dependent chain: the state carried from
// draw to draw is (sqrt(sig^2), 1/tau^2); sig^2 = bstar / G1 and tau^2 = btau* / G2 are formed with the reciprocals
// of the (prefetched) gamma variates, 1/(2 sig^2) = G1 / (2 bstar), and sqrt(d_j) is one rsqrt -- two reciprocals and
// one rsqrt on the critical path per draw instead of seven divisions and two square roots (same values to a few ulp).
// variates: D rows of [z_0 .. z_{p-1}, G1, G2] (injected numpy stream or the Philox table); sign_fix optional (p).
// canon (used when sign_fix is null, i.e. free-running mode): orient every eigenvector so that its projection on
// X'y is non-negative, z_j *= sign(ct_j).  The sign an eigensolver returns is arbitrary and flips under last-place
// changes of the Gram (another summation order: row shards, row permutations), which would pair the normals with
// other directions; with this convention the chain is a function of the Gram alone.
// gam (out) D x p row-major, sigs/taus (out) D.  Returns 1 if bstar < 0 was seen.
FOKL_HD int gibbs_chain(const Team &t, int p, const double *lamb, const double *ct, const CandConst &k,
                        const double *variates, const double *sign_fix, double *gam, double *sigs, double *taus,
                        double *red, bool canon = false)
{
    const int D = k.draws;
    const int w = p + 2;
    double ssig = sqrt(k.sigsqd0), itau = 1.0 / k.tausqd0;
    int bad = 0;
    // each thread owns elements e = tid, tid + nthr, ...; the first two are kept in registers, next row prefetched
    const int e0 = t.tid, e1 = t.tid + t.nthr;
    const bool h0 = e0 < p, h1 = e1 < p;
    double l0 = 0.0, c0 = 0.0, f0 = 1.0, l1 = 0.0, c1 = 0.0, f1 = 1.0;
    if (h0) { l0 = lamb[e0]; c0 = ct[e0]; if (sign_fix) f0 = sign_fix[e0]; else if (canon && c0 < 0.0) f0 = -1.0; }
    if (h1) { l1 = lamb[e1]; c1 = ct[e1]; if (sign_fix) f1 = sign_fix[e1]; else if (canon && c1 < 0.0) f1 = -1.0; }
    // software pipeline over the variate table, two rows deep: warps issue in order, so a value loaded in iteration
    // d - 1 is first touched in iteration d (row d + 1 is only *loaded* while row d is being used)
    double z0 = h0 ? f0 * variates[e0] : 0.0, z1 = h1 ? f1 * variates[e1] : 0.0;
    double g1 = variates[p], g2 = variates[p + 1];
    double ig1 = 1.0 / g1, ig2 = 1.0 / g2;
    double zr0 = 0.0, zr1 = 0.0, gr1 = 1.0, gr2 = 1.0;              // raw row d + 1
    if (D > 1) {
        const double *r1 = variates + w;
        if (h0) zr0 = r1[e0];
        if (h1) zr1 = r1[e1];
        gr1 = r1[p];
        gr2 = r1[p + 1];
    }
    for (int d = 0; d < D; ++d) {
        const double *row = variates + (int64_t)d * w;
        double zq0 = 0.0, zq1 = 0.0, gq1 = 1.0, gq2 = 1.0;          // raw row d + 2 (in flight during this iteration)
        if (d + 2 < D) {
            const double *r2 = row + 2 * w;
            if (h0) zq0 = r2[e0];
            if (h1) zq1 = r2[e1];
            gq1 = r2[p];
            gq2 = r2[p + 1];
        }
        double s1 = 0.0, s2 = 0.0, s3 = 0.0;
        if (h0) {
            double rs = FOKL_RSQRT(l0 + itau);
            double g = (rs * rs) * c0 + (ssig * rs) * z0;
            gam[(int64_t)d * p + e0] = g;
            s1 += l0 * g * g; s2 += g * c0; s3 += g * g;
        }
        if (h1) {
            double rs = FOKL_RSQRT(l1 + itau);
            double g = (rs * rs) * c1 + (ssig * rs) * z1;
            gam[(int64_t)d * p + e1] = g;
            s1 += l1 * g * g; s2 += g * c1; s3 += g * g;
        }
        for (int e = t.tid + 2 * t.nthr; e < p; e += t.nthr) {
            double z = row[e];
            double l = lamb[e], c = ct[e];
            if (sign_fix) z *= sign_fix[e];
            else if (canon && c < 0.0) z = -z;
            double rs = FOKL_RSQRT(l + itau);
            double g = (rs * rs) * c + (ssig * rs) * z;
            gam[(int64_t)d * p + e] = g;
            s1 += l * g * g; s2 += g * c; s3 += g * g;
        }
        // next row's values (loaded one iteration ago): off the critical path, overlaps the reduction
        const double zn0 = f0 * zr0, zn1 = f1 * zr1;
        const double ign1 = 1.0 / gr1, ign2 = 1.0 / gr2;
        team_sum3(t, s1, s2, s3, red, d & 1);
        const double bstar = k.b + 0.5 * (s1 - 2.0 * s2 + k.yty + s3 * itau);
        double sig, rb;
        if (bstar < 0.0) { sig = nan(""); rb = sig; bad = 1; }
        else { sig = bstar * ig1; rb = FOKL_RCP(bstar); }
        const double btau_star = (0.5 * g1 * rb) * s3 + k.btau;
        itau = g2 * FOKL_RCP(btau_star);
        ssig = sqrt(sig);
        if (t.tid == 0) { sigs[d] = sig; taus[d] = btau_star * ig2; }
        z0 = zn0; z1 = zn1; g1 = gr1; g2 = gr2; ig1 = ign1; ig2 = ign2;
        zr0 = zq0; zr1 = zq1; gr1 = gq1; gr2 = gq2;
    }
    return bad;
}

#if defined(__CUDACC__)
// The same draw loop for ONE warp holding E elements per lane in registers (p <= 32 E): the three sums are five
// shuffle steps with no shared memory and no CTA barrier on the per-draw critical path, and the variate rows are
// prefetched two draws ahead into registers.  The D draws are sequential, so a chain's time is D times the latency
// of one draw -- this halves it against the multi-warp version for the model sizes of a fit (p <= 256).
// Branch-free reciprocal square root / reciprocal of a positive double: the double-precision MUFU seed
// (rsqrt.approx.ftz.f64 / rcp.approx.ftz.f64, ~2^-22) and two Newton steps (~1e-16 relative error).  Unlike rsqrt() /
// 1.0 / x there is no slow-path branch or call, so the E independent evaluations of a draw interleave in the one
// warp that runs the chain and the dependent chain per evaluation is MUFU + 6 FMA-class instructions.
__device__ __forceinline__ double rsqrt_pos(double x)
{
    double y;
    asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(x));
    const double h = 0.5 * x;
    double t = fma(-h * y, y, 0.5);
    y = fma(y, t, y);
    t = fma(-h * y, y, 0.5);
    return fma(y, t, y);
}
__device__ __forceinline__ double rcp_pos(double x)
{
    double y;
    asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(x));
    double r = fma(-x, y, 1.0);
    y = fma(y, r, y);
    r = fma(-x, y, 1.0);
    return fma(y, r, y);
}

template <int E>
__device__ int gibbs_chain_warp(int lane, int p, const double *lamb, const double *ct, const CandConst &k,
                                const double *variates, const double *sign_fix, double *gam, double *sigs, double *taus,
                                bool canon = false)
{
    const int D = k.draws;
    const int w = p + 2;
    double ssig = sqrt(k.sigsqd0), itau = 1.0 / k.tausqd0;
    int bad = 0;
    double l[E], c[E], f[E], z[E], zr[E];
    bool h[E];
#pragma unroll
    for (int q = 0; q < E; ++q) {
        const int e = lane + 32 * q;
        h[q] = e < p;
        l[q] = h[q] ? lamb[e] : 1.0;
        c[q] = h[q] ? ct[e] : 0.0;
        f[q] = (h[q] && sign_fix) ? sign_fix[e] : ((canon && c[q] < 0.0) ? -1.0 : 1.0);
        z[q] = h[q] ? f[q] * variates[e] : 0.0;
        zr[q] = (h[q] && D > 1) ? variates[w + e] : 0.0;
    }
    double g1 = variates[p], g2 = variates[p + 1];
    double ig1 = rcp_pos(g1), ig2 = rcp_pos(g2);
    double gr1 = D > 1 ? variates[w + p] : 1.0, gr2 = D > 1 ? variates[w + p + 1] : 1.0;
    for (int d = 0; d < D; ++d) {
        const double *r2 = variates + (int64_t)(d + 2) * w;
        const bool more = d + 2 < D;
        double zq[E];
#pragma unroll
        for (int q = 0; q < E; ++q) zq[q] = (more && h[q]) ? r2[lane + 32 * q] : 0.0;
        const double gq1 = more ? r2[p] : 1.0, gq2 = more ? r2[p + 1] : 1.0;
        double *grow = gam + (int64_t)d * p;
        double gq[E];
#pragma unroll
        for (int q = 0; q < E; ++q) {
            const double rs = rsqrt_pos(l[q] + itau);
            gq[q] = h[q] ? (rs * rs) * c[q] + (ssig * rs) * z[q] : 0.0;
            if (h[q]) grow[lane + 32 * q] = gq[q];
        }
        // two independent partial sums per quantity: the adds of a draw do not form one E-long dependent chain
        double s1 = 0.0, s2 = 0.0, s3 = 0.0, s1b = 0.0, s2b = 0.0, s3b = 0.0;
#pragma unroll
        for (int q = 0; q < E; q += 2) {
            const double g = gq[q], gg = g * g;
            s1 = fma(l[q], gg, s1); s2 = fma(g, c[q], s2); s3 += gg;
            if (q + 1 < E) {
                const double g_ = gq[q + 1], gg_ = g_ * g_;
                s1b = fma(l[q + 1], gg_, s1b); s2b = fma(g_, c[q + 1], s2b); s3b += gg_;
            }
        }
        s1 += s1b; s2 += s2b; s3 += s3b;
        // next row's values (loaded one iteration ago): off the critical path, overlaps the reduction
        double zn[E];
#pragma unroll
        for (int q = 0; q < E; ++q) zn[q] = f[q] * zr[q];
        const double ign1 = rcp_pos(gr1), ign2 = rcp_pos(gr2);
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            s1 += __shfl_xor_sync(0xffffffffu, s1, o);
            s2 += __shfl_xor_sync(0xffffffffu, s2, o);
            s3 += __shfl_xor_sync(0xffffffffu, s3, o);
        }
        const double bstar = k.b + 0.5 * (s1 - 2.0 * s2 + k.yty + s3 * itau);
        double sig, rb;
        if (bstar < 0.0) { sig = nan(""); rb = sig; bad = 1; }
        else { sig = bstar * ig1; rb = rcp_pos(bstar); }
        const double btau_star = (0.5 * g1 * rb) * s3 + k.btau;
        itau = g2 * rcp_pos(btau_star);
        ssig = sig * rsqrt_pos(sig);
        if (lane == 0) { sigs[d] = sig; taus[d] = btau_star * ig2; }
#pragma unroll
        for (int q = 0; q < E; ++q) { z[q] = zn[q]; zr[q] = zq[q]; }
        g1 = gr1; g2 = gr2; ig1 = ign1; ig2 = ign2;
        gr1 = gq1; gr2 = gq2;
    }
    return bad;
}
#endif

}  // namespace fokl

// ---- kill-proposal scores without refactorising each candidate ------------------------------------------
// For the model with Gram A = G[idx][idx] (p x p, intercept first) the OLS fit of the model that drops column q
// has  SSE_{-q} = SSE + betahat_q^2 / (A^-1)_qq.  One Cholesky factorisation of A therefore yields the BIC
// (FR:1551-1554) of *every* single-column deletion the kill loop proposes (FR:1673-1685), instead of one
// eigendecomposition per proposal.  y is centred on its mean for accuracy (column 0 is the ones column, so the
// residual mean is zero and siglik = SSE / n).
//   L: p x p workspace, column-major (ld = p); z, beta: p scratch; wbuf: nwarp * p scratch
//   props: positions (within idx, never 0) of the proposals, k of them
//   ev_out[0..k-1] = BIC of the model without props[j];  ev_out[k] = BIC of the model itself
// Returns 0, or 1 if A is not numerically positive definite (caller falls back to the spectral path).
namespace fokl {

FOKL_HD void warp_sync(const Team &)
{
#if defined(__CUDA_ARCH__)
    __syncwarp();
#endif
}

FOKL_HD int kill_scores(const Team &t, const double *G, int64_t ldg, const double *Xty, const int *idx, int p,
                        const int *props, int k, const CandConst &c, double *L, double *z, double *beta,
                        double *wbuf, double *ev_out, volatile int *flag, double *red)
{
    for (int e = t.tid; e < p * p; e += t.nthr) {
        int col = e / p, row = e - col * p;
        L[e] = (row >= col) ? G[(int64_t)idx[row] * ldg + idx[col]] : 0.0;
    }
    if (t.tid == 0) *flag = 0;
    t.sync();
    // right-looking Cholesky; column j holds L[i][j], i >= j, at L[j*p + i]
    for (int j = 0; j < p; ++j) {
        const double d = L[(int64_t)j * p + j];
        // pivot relative to the column's own squared norm: below ~1e-11 the Schur complement is rounding noise
        // (duplicated / linearly dependent columns) and the deletion formula would divide by it
        if (!(d > 1e-11 * G[(int64_t)idx[j] * ldg + idx[j]])) {
            if (t.tid == 0) *flag = 1;
            break;
        }
        const double r = sqrt(d);
        t.sync();
        for (int i = j + t.tid; i < p; i += t.nthr) L[(int64_t)j * p + i] = (i == j) ? r : L[(int64_t)j * p + i] / r;
        t.sync();
        const int rem = p - j - 1;
        for (int64_t e = t.tid; e < (int64_t)rem * rem; e += t.nthr) {
            int mc = (int)(e / rem), mr = (int)(e - (int64_t)mc * rem);
            if (mr < mc) continue;
            int m = j + 1 + mc, i = j + 1 + mr;
            L[(int64_t)m * p + i] -= L[(int64_t)j * p + i] * L[(int64_t)j * p + m];
        }
        t.sync();
    }
    t.sync();
    if (*flag) return 1;
    const double ybar = c.sum_y / c.n;
    const int64_t row0 = (int64_t)idx[0] * ldg;
    // forward solve L z = X'y_c
    for (int i = t.tid; i < p; i += t.nthr) z[i] = Xty[idx[i]] - ybar * G[row0 + idx[i]];
    t.sync();
    for (int j = 0; j < p; ++j) {
        const double zj = z[j] / L[(int64_t)j * p + j];
        t.sync();
        if (t.tid == 0) z[j] = zj;
        for (int i = j + 1 + t.tid; i < p; i += t.nthr) z[i] -= L[(int64_t)j * p + i] * zj;
        t.sync();
    }
    double s1 = 0.0, s2 = 0.0, s3 = 0.0;
    for (int i = t.tid; i < p; i += t.nthr) s1 += z[i] * z[i];
    team_sum3(t, s1, s2, s3, red, 0);
    t.sync();
    const double sse = (c.yty - c.n * ybar * ybar) - s1;
    // back solve L' beta = z
    for (int i = t.tid; i < p; i += t.nthr) beta[i] = z[i];
    t.sync();
    for (int j = p - 1; j >= 0; --j) {
        const double bj = beta[j] / L[(int64_t)j * p + j];
        t.sync();
        if (t.tid == 0) beta[j] = bj;
        for (int i = t.tid; i < j; i += t.nthr) beta[i] -= L[(int64_t)i * p + j] * bj;
        t.sync();
    }
    // (A^-1)_qq = |L^-1 e_q|^2, one warp per proposal
    const double ln_n = log(c.n);
    for (int a = t.warp; a < k; a += t.nwarp) {
        const int q = props[a];
        double *w = wbuf + (int64_t)t.warp * p;
        for (int i = q + t.lane; i < p; i += t.nlane) w[i] = (i == q) ? 1.0 : 0.0;
        warp_sync(t);
        double acc = 0.0;
        for (int j = q; j < p; ++j) {
            const double wj = w[j] / L[(int64_t)j * p + j];
            acc += wj * wj;
            warp_sync(t);
            for (int i = j + 1 + t.lane; i < p; i += t.nlane) w[i] -= L[(int64_t)j * p + i] * wj;
            warp_sync(t);
        }
        if (t.lane == 0) {
            const double sig = (sse + beta[q] * beta[q] / acc) / c.n;
            ev_out[a] = (double)(p - 1) * ln_n - 2.0 * (-(c.n / 2.0) * log(sig) - (c.n - 1.0) / 2.0);
        }
    }
    if (t.tid == 0) {
        const double sig = sse / c.n;
        ev_out[k] = (double)p * ln_n - 2.0 * (-(c.n / 2.0) * log(sig) - (c.n - 1.0) / 2.0);
    }
    t.sync();
    return 0;
}

}  // namespace fokl

// ---- the whole kill loop of one substage on the device ---------------------------------------------------------
// FR:1666-1690 asks `gibbs` for one BIC per proposal, sequentially, each against the kill set accepted so far.
// All of those BICs follow from the *sweep operator* on the augmented, y-centred normal equations
//     T = [ A  z ; z' s ],  A = G[idx][idx],  z = X'(y - ybar),  s = (y - ybar)'(y - ybar):
// after sweeping every column, T = [ -A^-1  betahat ; betahat'  SSE ]; removing column q from the model is the
// reverse sweep on q (one rank-1 update, O(p^2), fully parallel) and costs  SSE_{-q} = SSE + T_qy^2 / (-T_qq).
// So one O(p^3) inversion + one O(p^2) update per *accepted* kill replaces one eigendecomposition per proposal,
// and the loop runs without a host round trip.  The acceptance threshold reads |mean intercept draw| of the last
// accepted model (FR:1671); the kernel uses the value it is given for the whole loop and the host verifies every
// round against the true chains afterwards (FoKL/_selection.py), re-running from the first round that differs.
//   T        (p + 1)^2 workspace (column-major, ld = p + 1)
//   cand_pos position in idx (>= 1) of candidate i, candidates in the reference's order (ascending |mean|)
//   out_i    [0] accepted kills, [1] proposals tested (= `gibbs` calls of the reference), [2] error flag,
//            [3 + k] candidate index of the k-th accepted kill, [3 + vm + k] proposals tested up to and incl. it
//   out_ev   [k] BIC (incl. the aic adjustment) of the model after the k-th accepted kill
namespace fokl {

struct KillLoopIn {
    double threshav, threshstda, threshstdb;
    double icpt;       // |mean(beters[h0:, 0])| used by the threshold test
    double evmin;      // BIC (incl. aic adjustment) of the model the loop starts from
    double aic_adj;    // (2 - ln n) if aic else 0, per model column
    int start;         // first candidate index to consider
};

FOKL_HD void team_atomic_min(int *addr, int v)
{
#if defined(__CUDA_ARCH__)
    atomicMin(addr, v);
#else
    if (v < *addr) *addr = v;
#endif
}

FOKL_HD void team_atomic_add(int *addr, int v)
{
#if defined(__CUDA_ARCH__)
    atomicAdd(addr, v);
#else
    *addr += v;
#endif
}

// one (reverse) sweep of pivot k on the symmetric (ld x ld) matrix T; sign = +1 sweep, -1 reverse sweep
// One pivot of the sweep operator on the symmetric (ld x ld) tableau T.  rowbuf (ld doubles of shared memory) takes a
// copy of pivot row k, after which every entry is final in ONE pass: T[i][j] -= (ck[i] ck[j]) / d off the pivot
// cross, sign * ck[i] / d on it, -1 / d at the pivot.  A warp walks whole rows (consecutive lanes = consecutive
// addresses) four strides at a time, so each lane has four independent loads in flight: the tableau of a 220-column
// model lives in L2, and the earlier one-element-at-a-time walk paid the full L2 latency per element.
FOKL_HD void sweep_pivot(const Team &t, double *T, int ld, int k, double sign, double *rowbuf)
{
    for (int i = t.tid; i < ld; i += t.nthr) rowbuf[i] = T[(int64_t)k * ld + i];
    t.sync();
    const double invD = 1.0 / rowbuf[k];
    const int step = t.nlane;
    for (int j = t.warp; j < ld; j += t.nwarp) {
        double *Tj = T + (int64_t)j * ld;
        const double ckj = rowbuf[j];
        if (j == k) {
            for (int i = t.lane; i < ld; i += step) Tj[i] = (i == k) ? -invD : sign * rowbuf[i] * invD;
            continue;
        }
        const double vk = sign * ckj * invD;
        int i = t.lane;
        for (; i + 3 * step < ld; i += 4 * step) {
            const double a0 = Tj[i], a1 = Tj[i + step], a2 = Tj[i + 2 * step], a3 = Tj[i + 3 * step];
            const double r0 = a0 - (rowbuf[i] * ckj) * invD, r1 = a1 - (rowbuf[i + step] * ckj) * invD;
            const double r2 = a2 - (rowbuf[i + 2 * step] * ckj) * invD, r3 = a3 - (rowbuf[i + 3 * step] * ckj) * invD;
            Tj[i] = (i == k) ? vk : r0;
            Tj[i + step] = (i + step == k) ? vk : r1;
            Tj[i + 2 * step] = (i + 2 * step == k) ? vk : r2;
            Tj[i + 3 * step] = (i + 3 * step == k) ? vk : r3;
        }
        for (; i < ld; i += step) {
            const double r0 = Tj[i] - (rowbuf[i] * ckj) * invD;
            Tj[i] = (i == k) ? vk : r0;
        }
    }
    t.sync();
}

// The tableau is symmetric, and the sweep keeps it so: the packed form stores the lower triangle row by row
// (T(i, j), i >= j, at i (i + 1) / 2 + j), half the footprint -- the tableau of a 226-column model (206 KB) then fits the
// shared memory of one SM instead of living in L2 (3 x 6 ms -> 3 x ~1 ms of kill loop per cfg4 fit).
FOKL_HD int64_t tri_index(int i, int j) { return i >= j ? (int64_t)i * (i + 1) / 2 + j : (int64_t)j * (j + 1) / 2 + i; }

FOKL_HD void sweep_pivot_packed(const Team &t, double *T, int ld, int k, double sign, double *rowbuf)
{
    for (int i = t.tid; i < ld; i += t.nthr) rowbuf[i] = T[tri_index(k, i)];
    t.sync();
    const double invD = 1.0 / rowbuf[k];
    for (int j = t.warp; j < ld; j += t.nwarp) {                  // row j: entries (j, 0 .. j), contiguous
        double *Tj = T + (int64_t)j * (j + 1) / 2;
        const double ckj = rowbuf[j];
        if (j == k) {
            for (int i = t.lane; i <= j; i += t.nlane) Tj[i] = (i == k) ? -invD : sign * rowbuf[i] * invD;
            continue;
        }
        const double vk = sign * ckj * invD;
        for (int i = t.lane; i <= j; i += t.nlane) {
            const double r0 = Tj[i] - (rowbuf[i] * ckj) * invD;
            Tj[i] = (i == k) ? vk : r0;
        }
    }
    t.sync();
}

template <bool PACKED>
FOKL_HD int kill_loop_t(const Team &t, const double *G, int64_t ldg, const double *Xty, const int *idx, int p,
                      const int *cand_pos, const double *bv0, const double *bv1, int vm, const CandConst &c,
                      const KillLoopIn &in, double *T, int *out_i, double *out_ev, int *sh /* 4 shared ints */,
                      double *rowbuf /* p + 1 shared doubles */, const double *T0 = nullptr, int ldt0 = 0)
{
    const int ld = p + 1;
    const double ybar = c.sum_y / c.n;
    const int64_t row0 = (int64_t)idx[0] * ldg;
    // T0 (optional): the tableau as it is AFTER the p forward sweeps, formed from the model's eigendecomposition by
    // kill_tableau_kernel ((p + 1) x ldt0, symmetric; the entry behind it is 1.0 if it may be used, see there) -- the p
    // sequential pivots, two thirds of this kernel's time on a 226-column model, are then not run.
    const bool pre = T0 != nullptr && T0[(int64_t)ld * ldt0] > 0.5;
    for (int e = t.tid; e < ld * ld; e += t.nthr) {
        int j = e / ld, i = e - j * ld;
        if (PACKED && i < j) continue;
        double v;
        if (pre) v = T0[(int64_t)i * ldt0 + j];
        else if (i < p && j < p) v = G[(int64_t)idx[i] * ldg + idx[j]];
        else if (i == p && j == p) v = c.yty - c.n * ybar * ybar;
        else {
            int q = i < p ? i : j;
            v = Xty[idx[q]] - ybar * G[row0 + idx[q]];
        }
        T[PACKED ? tri_index(i, j) : (int64_t)e] = v;
    }
    if (t.tid == 0) { sh[0] = 0; sh[1] = 0; sh[2] = 0; sh[3] = 0; }
    t.sync();
    int bad = 0;
    for (int k = 0; k < p && !pre; ++k) {
        const double d = T[PACKED ? tri_index(k, k) : (int64_t)k * ld + k];
        if (!(d > 1e-11 * G[(int64_t)idx[k] * ldg + idx[k]])) { bad = 1; break; }   // not numerically positive definite
        t.sync();
        if (PACKED) sweep_pivot_packed(t, T, ld, k, 1.0, rowbuf);
        else sweep_pivot(t, T, ld, k, 1.0, rowbuf);
    }
    int n_acc = 0, tested = 0, pa = p, cur = in.start;
    double evmin = in.evmin;
    const double ln_n = log(c.n);
    const double thr = in.threshav * in.icpt;
    while (!bad && cur < vm) {
        if (t.tid == 0) { sh[0] = 0x7fffffff; sh[1] = 0; sh[2] = 0; }
        t.sync();
        const double sse = T[PACKED ? tri_index(p, p) : (int64_t)p * ld + p];
        for (int i = cur + t.tid; i < vm; i += t.nthr) {
            const bool prop = (bv1[i] > in.threshstdb) || (bv1[i] > in.threshstda && bv0[i] < thr);
            if (!prop) continue;
            const int q = cand_pos[i];
            const double tqq = T[PACKED ? tri_index(q, q) : (int64_t)q * ld + q], tqy = T[PACKED ? tri_index(p, q) : (int64_t)p * ld + q];
            if (!(tqq < 0.0)) { sh[2] = 1; continue; }
            const double sig = (sse + tqy * tqy / (-tqq)) / c.n;
            const double evt = (double)(pa - 1) * ln_n - 2.0 * (-(c.n / 2.0) * log(sig) - (c.n - 1.0) / 2.0) +
                               in.aic_adj * (double)(pa - 1);
            if (evt < evmin) team_atomic_min(&sh[0], i);
        }
        t.sync();
        const int hit = sh[0];
        if (sh[2]) { bad = 2; break; }
        for (int i = cur + t.tid; i < vm && i <= hit; i += t.nthr) {
            const bool prop = (bv1[i] > in.threshstdb) || (bv1[i] > in.threshstda && bv0[i] < thr);
            if (prop) team_atomic_add(&sh[1], 1);
        }
        t.sync();
        tested += sh[1];
        if (hit == 0x7fffffff) break;
        const int q = cand_pos[hit];
        const double tqq = T[PACKED ? tri_index(q, q) : (int64_t)q * ld + q], tqy = T[PACKED ? tri_index(p, q) : (int64_t)p * ld + q];
        const double sig = (sse + tqy * tqy / (-tqq)) / c.n;
        evmin = (double)(pa - 1) * ln_n - 2.0 * (-(c.n / 2.0) * log(sig) - (c.n - 1.0) / 2.0) + in.aic_adj * (double)(pa - 1);
        t.sync();
        if (PACKED) sweep_pivot_packed(t, T, ld, q, -1.0, rowbuf);
        else sweep_pivot(t, T, ld, q, -1.0, rowbuf);
        if (t.tid == 0) {
            out_i[3 + n_acc] = hit;
            out_i[3 + vm + n_acc] = tested;
            out_ev[n_acc] = evmin;
        }
        n_acc += 1;
        pa -= 1;
        cur = hit + 1;
    }
    if (t.tid == 0) { out_i[0] = n_acc; out_i[1] = tested; out_i[2] = bad; }
    t.sync();
    return bad;
}

// T: (p + 1)^2 doubles (full form) or (p + 1)(p + 2) / 2 doubles (packed form)
FOKL_HD int kill_loop(const Team &t, const double *G, int64_t ldg, const double *Xty, const int *idx, int p,
                      const int *cand_pos, const double *bv0, const double *bv1, int vm, const CandConst &c,
                      const KillLoopIn &in, double *T, int *out_i, double *out_ev, int *sh /* 4 shared ints */,
                      double *rowbuf /* p + 1 shared doubles */, bool packed = false, const double *T0 = nullptr, int ldt0 = 0)
{
    return packed ? kill_loop_t<true>(t, G, ldg, Xty, idx, p, cand_pos, bv0, bv1, vm, c, in, T, out_i, out_ev, sh, rowbuf, T0, ldt0)
                  : kill_loop_t<false>(t, G, ldg, Xty, idx, p, cand_pos, bv0, bv1, vm, c, in, T, out_i, out_ev, sh, rowbuf, T0, ldt0);
}

}  // namespace fokl
