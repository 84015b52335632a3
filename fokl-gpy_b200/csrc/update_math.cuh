// update_math.cuh -- the three-case sampler of `fitupdate` (update=True) in spectral coordinates.
//
// Replaces the draw loops of `gibbs_Xin_update` (src/FoKL/FoKLRoutines.py, "FR"):
//   case 1  FR:2097-2140   first fit of an update model: the eigenbasis chain of `fit` with the per-draw
//                          log-likelihood FR:2111-2115, whose maximum is the model's evidence (FR:2143)
//   case 2  FR:2194-2255   same terms, Gaussian prior N(mu_old, tausqd * Sigma_old) on all coefficients.  The reference
//                          calls eigh and inv on  XotXo + Sigma_old^-1 / tausqd  inside the loop (2000 factorisations per
//                          call); here ONE generalised eigendecomposition T'(XotXo)T = D, T'(Sigma_old^-1)T = I
//                          (host side: FoKL/_update.py) diagonalises all of them, and a draw costs O(p)
//   case 3  FR:2352-2417   new terms next to the old ones: block Gibbs between beta_old (prior N(mu_old, Sigma_old)) and
//                          beta_new (prior N(0, tausqd)); in the coordinates gam_o = Q_o' beta_o, gam_n = Q_n' beta_n of
//                          the two FIXED eigendecompositions (FR:2296, 2312) every per-draw inverse (FR:2353, 2365) is a
//                          diagonal scaling and the coupling one po x pn matrix M = Q_o' Xo'Xn Q_n
// Written against the Team abstraction of cand_math.cuh (one CTA on the device, one sequential thread in the host
// emulation build, tests/host_emu), so the arithmetic is checked against the CPU oracle without a GPU.
#pragma once
#include "cand_math.cuh"

namespace fokl {

struct UpdModel {
    int mode;                 // 1, 2, 3: the reference's case
    int po, pn;               // widths of the old / new coefficient blocks (mode 1: po = 0; mode 2: pn = 0)
    int draws;
    double astar, atau_star;  // gamma shapes (FR:2086-2087, 2176-2177, 2327-2328)
    double b, btau;           // FR:1926-1929
    double sigsqd0;           // chain start; tausqd starts at 1 / sigsqd0 (FR:2067, 2167, 2282)
    double yty;               // y'y
    double squerr;            // mode 1: |y - X betahat|^2 (FR:2083)
    double n;                 // number of data rows
};

struct UpdArrays {
    // mode 1: lam_n, c_n = Q'X'y                                              (pn = p)
    // mode 2: lam_o = D, c_o = T'X'y, m_o = T^-1 mu_old                       (po = p)
    // mode 3: lam_o, c_o = Q_o'(Xo'y + Sigma^-1 mu), t_o = Q_o'Xo'y, m_o = Q_o'mu, lam_n, c_n = Q_n'Xn'y,
    //         M (po x pn row-major), Mt (pn x po row-major), K = Q_o'Xo'XoQ_o, W = Q_o'Sigma^-1 Q_o (po x po row-major)
    const double *lam_o, *c_o, *t_o, *m_o, *lam_n, *c_n, *M, *Mt, *K, *W;
};

// shared scratch of update_chain, in doubles
FOKL_HD int update_scratch_doubles(int po, int pn, int nwarp) { return 4 * po + pn + 2 * 8 * nwarp + 8; }

// CTA-wide sums of NQ values, every thread returns the same bits (fixed order).  red: 2 * 8 * nwarp doubles.
template <int NQ>
FOKL_HD void team_sum_n(const Team &t, double *q, double *red, int parity)
{
#if defined(__CUDA_ARCH__)
#pragma unroll
    for (int k = 0; k < NQ; ++k)
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) q[k] += __shfl_xor_sync(0xffffffffu, q[k], o);
#endif
    double *r = red + parity * 8 * t.nwarp;
    if (t.lane == 0)
        for (int k = 0; k < NQ; ++k) r[8 * t.warp + k] = q[k];
    t.sync();
    for (int k = 0; k < NQ; ++k) {
        double s = 0.0;
        for (int w = 0; w < t.nwarp; ++w) s += r[8 * w + k];
        q[k] = s;
    }
}

// One variate of the call: row d = [z_0 .. z_{po + pn - 1}, G1, G2] in the order the reference consumes them
// (normal(po) FR:2213 / 2355, normal(pn) FR:2106 / 2367, gamma(astar) FR:2132 / 2233 / 2395, gamma(atau_star)).
struct UpdVariates {
    const double *table;      // injected (parity mode) or null
    Philox g;
    uint32_t s_lo, s_hi;
    int w;                    // po + pn
    double astar, atau_star;
    FOKL_HD double z(int d, int e) const
    {
        if (table) return table[(int64_t)d * (w + 2) + e];
        return philox_normal(g, s_lo, s_hi, (uint32_t)d, (uint32_t)e);
    }
    FOKL_HD double gamma1(int d) const
    {
        if (table) return table[(int64_t)d * (w + 2) + w];
        return philox_gamma(g, s_lo, s_hi, (uint32_t)d, 0u, astar);
    }
    FOKL_HD double gamma2(int d) const
    {
        if (table) return table[(int64_t)d * (w + 2) + w + 1];
        return philox_gamma(g, s_lo, s_hi, (uint32_t)d, 1024u, atau_star);
    }
};

// row . vec by one warp (every lane returns the sum on the device; the lone host thread sums everything)
FOKL_HD double warp_row_dot(const Team &t, const double *row, const double *vec, int len)
{
    double s = 0.0;
    for (int j = t.lane; j < len; j += t.nlane) s += row[j] * vec[j];
    return warp_sum1(t, s);
}

// The chain.  gam_o (draws x po), gam_n (draws x pn): the draws in spectral coordinates (beta = Q gam, formed by the
// caller); sigs, taus, lik: draws each.  sh: update_scratch_doubles() doubles of shared memory.  Returns 1 if bstar < 0
// was seen (FR:2128, 2229, 2391: sigsqd = nan).
FOKL_HD int update_chain(const Team &t, const UpdModel &m, const UpdArrays &A, const UpdVariates &V, double *gam_o,
                         double *gam_n, double *sigs, double *taus, double *lik, double *sh)
{
    const int po = m.po, pn = m.pn, D = m.draws;
    double *go = sh, *dv = go + po, *u = dv + po, *io = u + po, *gn = io + po, *red = gn + pn;
    // io: 1 / lam_o (mode 3); u = M gam_n of the previous draw (zero before the first: FR:2352 reads betas_new[-1])
    for (int e = t.tid; e < po; e += t.nthr) {
        u[e] = 0.0;
        io[e] = (m.mode == 3) ? 1.0 / A.lam_o[e] : 0.0;
    }
    t.sync();
    double sig = m.sigsqd0, itau = m.sigsqd0;          // 1 / tausqd0 = sigsqd0
    const double half_n = m.n / 2.0;
    int bad = 0;
    for (int d = 0; d < D; ++d) {
        const double ssig = sqrt(sig);
        const double g1 = V.gamma1(d), g2 = V.gamma2(d);
        double q[7] = {0.0, 0.0, 0.0, 0.0, 0.0, 0.0, 0.0};
        double sse, bstar;
        if (m.mode == 1) {
            // gam = d c + sqrt(sig) sqrt(d) z, d = 1 / (lam + 1 / tausqd)
            for (int e = t.tid; e < pn; e += t.nthr) {
                const double l = A.lam_n[e], c = A.c_n[e];
                const double dd = 1.0 / (l + itau);
                const double g = dd * c + ssig * sqrt(dd) * V.z(d, e);
                gam_n[(int64_t)d * pn + e] = g;
                const double r = c / l - g;
                q[0] += l * g * g; q[1] += g * c; q[2] += g * g; q[3] += l * r * r;
            }
            team_sum_n<4>(t, q, red, d & 1);
            // likelihood at the draw, with the sigma squared it was drawn under (FR:2111-2115)
            if (t.tid == 0) lik[d] = -half_n * log(sig) - (m.squerr + q[3]) / (2.0 * sig);
            bstar = m.b + 0.5 * (q[0] - 2.0 * q[1] + m.yty + q[2] * itau);
            sse = 0.0;
        } else if (m.mode == 2) {
            for (int e = t.tid; e < po; e += t.nthr) {
                const double l = A.lam_o[e], c = A.c_o[e], mc = A.m_o[e];
                const double dd = 1.0 / (l + itau);
                const double g = dd * (c + itau * mc) + ssig * sqrt(dd) * V.z(d, e);
                gam_o[(int64_t)d * po + e] = g;
                const double r = g - mc;
                q[0] += l * g * g; q[1] += g * c; q[2] += r * r;
            }
            team_sum_n<3>(t, q, red, d & 1);
            sse = m.yty - 2.0 * q[1] + q[0];
            bstar = 0.5 * sse + 0.5 * itau * q[2] + m.b;
        } else {
            // old block: mean (Xo'Xo + Sigma^-1)^-1 (Xo'y - Xo'Xn beta_n + Sigma^-1 mu), FR:2352-2356
            for (int e = t.tid; e < po; e += t.nthr) {
                const double g = (A.c_o[e] - u[e]) * io[e] + ssig * sqrt(io[e]) * V.z(d, e);
                go[e] = g;
                dv[e] = g - A.m_o[e];
                gam_o[(int64_t)d * po + e] = g;
                q[0] += g * A.t_o[e];
            }
            t.sync();
            // new block: mean (Xn'Xn + I / tausqd)^-1 (Xn'y - Xn'Xo beta_o), FR:2358-2368
            for (int j = t.warp; j < pn; j += t.nwarp) {
                const double v = warp_row_dot(t, A.Mt + (int64_t)j * po, go, po);
                if (t.lane == 0) {
                    const double l = A.lam_n[j], c = A.c_n[j];
                    const double dd = 1.0 / (l + itau);
                    const double g = dd * (c - v) + ssig * sqrt(dd) * V.z(d, po + j);
                    gn[j] = g;
                    gam_n[(int64_t)d * pn + j] = g;
                    q[1] += g * c; q[4] += l * g * g; q[6] += g * g;
                }
            }
            t.sync();
            // quadratic forms of FR:2371-2384 (and u = M gam_n for the next draw)
            for (int i = t.warp; i < po; i += t.nwarp) {
                const double r1 = warp_row_dot(t, A.M + (int64_t)i * pn, gn, pn);
                const double r2 = warp_row_dot(t, A.K + (int64_t)i * po, go, po);
                const double r3 = warp_row_dot(t, A.W + (int64_t)i * po, dv, po);
                if (t.lane == 0) {
                    u[i] = r1;
                    q[3] += go[i] * r1; q[2] += go[i] * r2; q[5] += dv[i] * r3;
                }
            }
            team_sum_n<7>(t, q, red, d & 1);
            sse = m.yty - 2.0 * (q[0] + q[1]) + q[2] + 2.0 * q[3] + q[4];
            bstar = 0.5 * sse + 0.5 * itau * q[6] + 0.5 * q[5] + m.b;
        }
        if (bstar < 0.0) { sig = nan(""); bad = 1; }
        else sig = 1.0 / ((1.0 / bstar) * g1);
        double btau_star;
        if (m.mode == 1) btau_star = (1.0 / (2.0 * sig)) * q[2] + m.btau;             // FR:2137
        else if (m.mode == 2) btau_star = 0.5 * (1.0 / sig) * q[2] + m.btau;           // FR:2238-2243
        else btau_star = (1.0 / (2.0 * sig)) * q[6] + m.btau;                          // FR:2400
        const double tau = 1.0 / ((1.0 / btau_star) * g2);
        itau = 1.0 / tau;
        if (t.tid == 0) {
            sigs[d] = sig;
            taus[d] = tau;
            if (m.mode != 1) lik[d] = -half_n * log(sig) - 0.5 / sig * sse;            // FR:2249-2254, 2408-2416
        }
    }
    return bad;
}

}  // namespace fokl
