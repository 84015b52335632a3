// fokl_math.cuh -- scalar math shared by the sm_100a kernels and the host emulation build.
//
// Every expression here that must reproduce numpy's separately-rounded float64 arithmetic uses the
// explicit round-to-nearest intrinsics (device) or plain operators compiled with -ffp-contract=off
// (host emulation, tests/host_emu), so nvcc's default FMA contraction cannot change results.
#pragma once
#include <stdint.h>
#include <math.h>

#if defined(__CUDACC__)
#define FOKL_HD __host__ __device__ __forceinline__
#else
#define FOKL_HD inline
#endif

#if defined(__CUDA_ARCH__)
#define FOKL_MUL(a, b) __dmul_rn((a), (b))
#define FOKL_ADD(a, b) __dadd_rn((a), (b))
#define FOKL_SUB(a, b) __dsub_rn((a), (b))
#define FOKL_FMA(a, b, c) __fma_rn((a), (b), (c))
#else
#define FOKL_MUL(a, b) ((a) * (b))
#define FOKL_ADD(a, b) ((a) + (b))
#define FOKL_SUB(a, b) ((a) - (b))
#define FOKL_FMA(a, b, c) fma((a), (b), (c))
#endif

namespace fokl {

// ---- _inputs_to_phind (FoKLRoutines.py:570-589) ------------------------------------------------
// phind = ceil(fl(n_piece * x)) as uint16, 0 -> 1, then - 1;  xsm = fl(fl(n_piece * x) - phind).
// Returns false when x is outside the range the reference accepts (it raises / indexes out of range).
FOKL_HD bool phind_xsm(double x, int n_piece, int &ph, double &xsm)
{
    double t = FOKL_MUL(x, (double)n_piece);
    double c = ceil(t);
    bool ok = (c >= 0.0) && (c <= (double)n_piece);   // -0.0 >= 0.0 holds; NaN fails
    int p = ok ? (int)c : 1;
    if (p == 0) p = 1;
    p -= 1;
    ph = p;
    xsm = FOKL_SUB(t, (double)p);
    return ok;
}

// Correctly-rounded x^2 and x^3 (Python evaluates x ** 2, x ** 3 through libm pow()).
FOKL_HD void square_cube(double x, double &x2, double &x3)
{
    x2 = FOKL_MUL(x, x);
    double e = FOKL_FMA(x, x, -x2);              // x*x = x2 + e exactly
    x3 = FOKL_FMA(x2, x, FOKL_MUL(e, x));        // (x2 + e) * x rounded once
}

// evaluate_basis, cubic, d = 0 (FoKLRoutines.py:836): c0 + c1*x + c2*(x**2) + c3*(x**3), left to right.
FOKL_HD double cubic_basis(double c0, double c1, double c2, double c3, double x, double x2, double x3)
{
    double s = FOKL_ADD(c0, FOKL_MUL(c1, x));
    s = FOKL_ADD(s, FOKL_MUL(c2, x2));
    s = FOKL_ADD(s, FOKL_MUL(c3, x3));
    return s;
}

// Powers x^1 .. x^deg, each correctly rounded (double-double running product), into pw[1..deg].
FOKL_HD void powers_dd(double x, int deg, double *pw)
{
    double h = x, l = 0.0;
    pw[1] = x;
    for (int q = 2; q <= deg; ++q) {
        double p = FOKL_MUL(h, x);
        double e = FOKL_FMA(h, x, -p);
        double t = FOKL_FMA(l, x, e);
        double nh = FOKL_ADD(p, t);
        l = FOKL_SUB(t, FOKL_SUB(nh, p));
        h = nh;
        pw[q] = h;
    }
}

// evaluate_basis, Bernoulli, d = 0 (FoKLRoutines.py:843): c[0] + sum(c[k] * x**k, k = 1 .. n_coef-1)
// with Python's sum() accumulating left to right from 0.
FOKL_HD double bernoulli_basis(const double *c, int n_coef, const double *pw)
{
    double s = 0.0;
    for (int q = 1; q < n_coef; ++q) s = FOKL_ADD(s, FOKL_MUL(c[q], pw[q]));
    return FOKL_ADD(c[0], s);
}

// ---- bss_derivatives (FoKLRoutines.py:594-805) --------------------------------------------------
// The derivative routine does not use xsm: it evaluates every factor at the "twice normalised" input of
// FR:584-586,  X = fl(fl(x - fl(phind * r)) / r)  with  r = fl(1 / n_piece)  (phind 0-based).
FOKL_HD double twice_normalised(double x, int n_piece, int ph)
{
    const double r = 1.0 / (double)n_piece;
    const double xmin = FOKL_MUL((double)ph, r);
#if defined(__CUDA_ARCH__)
    return __ddiv_rn(__dsub_rn(x, xmin), r);
#else
    return (x - xmin) / r;
#endif
}

// evaluate_basis, cubic, d = 1 (FR:838): c1 + 2*c2*x + 3*c3*(x**2), left to right (2*c2 is exact).
FOKL_HD double cubic_basis_d1(double c1, double c2, double c3, double x, double x2)
{
    double s = FOKL_ADD(c1, FOKL_MUL(FOKL_MUL(2.0, c2), x));
    return FOKL_ADD(s, FOKL_MUL(FOKL_MUL(3.0, c3), x2));
}

// evaluate_basis, cubic, d = 2 (FR:840): 2*c2 + 6*c3*x.
FOKL_HD double cubic_basis_d2(double c2, double c3, double x)
{
    return FOKL_ADD(FOKL_MUL(2.0, c2), FOKL_MUL(FOKL_MUL(6.0, c3), x));
}

// evaluate_basis, Bernoulli, d = 1 (FR:845): c[1] + sum(k*c[k]*x**(k-1), k = 2 ..) and d = 2 (FR:847):
// sum((k-1)*k*c[k]*x**(k-2), k = 2 ..), Python's sum() accumulating left to right from 0; the powers are correctly
// rounded (double-double running product, like powers_dd) and x**0 == 1 exactly.
FOKL_HD double bernoulli_basis_deriv(const double *c, int n_coef, double x, int d)
{
    double h = x, l = 0.0, s = 0.0;      // h + l ~ x^e
    int e = 1;
    for (int q = 2; q < n_coef; ++q) {
        const int need = q - d;
        const double coef = FOKL_MUL(d == 1 ? (double)q : (double)((q - 1) * q), c[q]);
        double term = coef;
        if (need > 0) {
            while (e < need) {
                double p = FOKL_MUL(h, x);
                double err = FOKL_FMA(h, x, -p);
                double t = FOKL_FMA(l, x, err);
                double nh = FOKL_ADD(p, t);
                l = FOKL_SUB(t, FOKL_SUB(nh, p));
                h = nh;
                ++e;
            }
            term = FOKL_MUL(coef, h);
        }
        s = FOKL_ADD(s, term);
    }
    return d == 1 ? FOKL_ADD(c[1], s) : s;
}

// ---- Philox4x32-10 counter RNG (free-running mode) ---------------------------------------------
struct Philox {
    uint32_t k0, k1;
    FOKL_HD static void mulhilo(uint32_t a, uint32_t b, uint32_t &hi, uint32_t &lo)
    {
        uint64_t p = (uint64_t)a * (uint64_t)b;
        hi = (uint32_t)(p >> 32);
        lo = (uint32_t)p;
    }
    FOKL_HD void operator()(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3, uint32_t out[4]) const
    {
        uint32_t ka = k0, kb = k1;
        for (int r = 0; r < 10; ++r) {
            uint32_t hi0, lo0, hi1, lo1;
            mulhilo(0xD2511F53u, c0, hi0, lo0);
            mulhilo(0xCD9E8D57u, c2, hi1, lo1);
            uint32_t n0 = hi1 ^ c1 ^ ka, n1 = lo1, n2 = hi0 ^ c3 ^ kb, n3 = lo0;
            c0 = n0; c1 = n1; c2 = n2; c3 = n3;
            ka += 0x9E3779B9u; kb += 0xBB67AE85u;
        }
        out[0] = c0; out[1] = c1; out[2] = c2; out[3] = c3;
    }
};

// uniform in (0, 1) from 53 random bits
FOKL_HD double u01(uint32_t a, uint32_t b)
{
    uint64_t v = (((uint64_t)a << 32) | b) >> 11;
    return ((double)v + 0.5) * (1.0 / 9007199254740992.0);
}

// one standard normal from a (stream, draw, index) counter: Box-Muller
FOKL_HD double philox_normal(const Philox &g, uint32_t stream_lo, uint32_t stream_hi, uint32_t draw,
                             uint32_t idx)
{
    uint32_t r[4];
    g(draw, idx, stream_lo, stream_hi, r);
    double u1 = u01(r[0], r[1]);
    double u2 = u01(r[2], r[3]);
    double rad = sqrt(-2.0 * log(u1));
    return rad * cos(6.283185307179586476925286766559 * u2);
}

// standard_gamma(shape) by Marsaglia-Tsang (shape < 1 boosted); idx_base selects a private sub-stream.
FOKL_HD double philox_gamma(const Philox &g, uint32_t stream_lo, uint32_t stream_hi, uint32_t draw,
                            uint32_t idx_base, double shape)
{
    double boost = 1.0;
    double a = shape;
    uint32_t r[4];
    if (a < 1.0) {
        g(draw, idx_base, stream_lo, stream_hi ^ 0x80000000u, r);
        boost = pow(u01(r[0], r[1]), 1.0 / a);
        a += 1.0;
    }
    double d = a - 1.0 / 3.0;
    double c = 1.0 / sqrt(9.0 * d);
    for (uint32_t it = 1; it < 64; ++it) {
        g(draw, idx_base + it, stream_lo, stream_hi ^ 0x80000000u, r);
        double u1 = u01(r[0], r[1]);
        double u2 = u01(r[2], r[3]);
        double rad = sqrt(-2.0 * log(u1));
        double x = rad * cos(6.283185307179586476925286766559 * u2);
        double v = 1.0 + c * x;
        if (v <= 0.0) continue;
        v = v * v * v;
        g(draw, idx_base + it, stream_lo ^ 0x5bd1e995u, stream_hi ^ 0x80000000u, r);
        double u = u01(r[0], r[1]);
        double x2 = x * x;
        if (u < 1.0 - 0.0331 * x2 * x2) return boost * d * v;
        if (log(u) < 0.5 * x2 + d * (1.0 - v + log(v))) return boost * d * v;
    }
    return boost * d;   // unreachable in practice (acceptance > 95 % per iteration)
}

}  // namespace fokl
