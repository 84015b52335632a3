/* CPU oracle helper -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.
 *
 * Plain-C restatement of the design-matrix loop of the reference,
 *   /root/reference/src/FoKL/FoKLRoutines.py:1446-1485  (X[i][j] = prod_k phi(x_ik))
 * with   _inputs_to_phind   FoKLRoutines.py:570-589      (piece index, local coordinate)
 * and    evaluate_basis d=0 FoKLRoutines.py:834-836, 841-843.
 *
 * Python evaluates `x ** k` on numpy float64 scalars through libm pow(); every product and sum is
 * rounded separately.  Build with -ffp-contract=off so gcc does not fuse them.
 */
#include <math.h>
#include <stdint.h>
#ifdef _OPENMP
#include <omp.h>
#endif

/* kernel: 0 = cubic  (table [n_orders][row_len = n_piece][4])
 *         1 = bernoulli (table [n_orders][row_len]; order n uses n + 2 coefficients) */
int oracle_basis_columns(const double *x, int64_t n, int m, const int32_t *terms, int c, int kernel,
                         const double *table, int n_orders, int row_len, double *out, int n_threads)
{
    int bad = 0;
    (void)n_orders;
#ifdef _OPENMP
    if (n_threads < 1) n_threads = 1;
#pragma omp parallel for num_threads(n_threads) schedule(static) reduction(| : bad)
#endif
    for (int64_t i = 0; i < n; ++i) {
        for (int j = 0; j < c; ++j) {
            double phi = 1.0;
            for (int k = 0; k < m; ++k) {
                int num = terms[(int64_t)j * m + k];
                if (num == 0) continue;
                int nid = num - 1;
                double xi = x[i * m + k];
                double basis;
                if (kernel == 0) {
                    /* FR:571-589 */
                    double t = ceil(xi * (double)row_len);
                    if (!(t >= 0.0 && t <= 65535.0)) { bad = 1; t = 1.0; }
                    int ph = (int)(uint16_t)t;
                    if (ph == 0) ph = 1;
                    ph -= 1;
                    if (ph > row_len - 1) { bad = 1; ph = row_len - 1; }
                    double xs = (double)row_len * xi - (double)ph;
                    const double *cf = table + ((int64_t)nid * row_len + ph) * 4;
                    /* FR:836 */
                    basis = cf[0] + cf[1] * xs + cf[2] * pow(xs, 2.0) + cf[3] * pow(xs, 3.0);
                } else {
                    /* FR:843: c[0] + sum(c[k] * x**k for k in 1..n+1); Python's sum() starts at int 0 */
                    const double *cf = table + (int64_t)nid * row_len;
                    double s = 0.0;
                    for (int q = 1; q < nid + 2; ++q) s = s + cf[q] * pow(xi, (double)q);
                    basis = cf[0] + s;
                }
                phi = phi * basis;
            }
            out[i * c + j] = phi;
        }
    }
    return bad;
}
