"""Regenerate a 'Cubic Splines' coefficient table (test infrastructure + product data generator).

The upstream data file `src/FoKL/kernels/splineCoefficient500_highPrecision_smoothed.txt` is a missing
large blob in the reference checkout (`.MISSING_LARGE_BLOBS:4`), so `getKernels.sp500()`
(`src/FoKL/getKernels.py:221-267`) cannot run.  This script rebuilds a table of the same *layout*
(`phis[s][k][piece]`, s < 500, k < 4, piece < 499; `FoKLRoutines.py:117-119`) following the upstream
description of how the basis functions were derived
(`docs/_dev/basis_functions/bernoulli_polynomials/main.ipynb` cell 0; kernel formula
`src/FoKL/getKernels.py:280-290`):

  1. kappa_1(x, x') = B1(x)B1(x') + B2(x)B2(x') - B4(|x - x'|)/24 on linspace(0, 1, 500)
  2. eigendecompose, order by decreasing eigenvalue, scale eigenvectors by sqrt(eigenvalue)
  3. fit a cubic spline through each scaled eigenvector
  4. express each of the 499 pieces as a cubic in the local coordinate t = 499*x - piece in (0, 1]
     (the coordinate `_inputs_to_phind` produces, `FoKLRoutines.py:570-589`)

Sign convention: the sign of phi_s(0) follows the shipped Bernoulli table for s < 20
(`main.ipynb` cell 9 flipped the Bernoulli set to match the splines) and (-1)**(s+1) beyond.

It is NOT bit-identical to upstream's smoothed file; the same saved table is always injected into
both the reference/oracle and the CUDA path via the documented `phis=` hyperparameter.
"""
import numpy as np


def _b1(x):
    return x - 0.5


def _b2(x):
    return x ** 2 - x + 1.0 / 6.0


def _b4(x):
    return x ** 4 - 2 * x ** 3 + x ** 2 - 1.0 / 30.0


# sign of the constant coefficient of each row of the shipped Bernoulli table
# (`src/FoKL/kernels/orthogonal_Bn_scaled.txt`, column 0)
_BERNOULLI_SIGN_AT_0 = (-1, 1, -1, 1, 1, -1, 1, -1, 1, -1, 1, -1, 1, 1, 1, -1, -1, 1, -1, 1)


def generate(n_orders=500, n_grid=500):
    """Return float64 array [n_orders][499][4] (order, piece, power-of-t)."""
    from scipy.interpolate import CubicSpline

    x = np.linspace(0.0, 1.0, n_grid)
    xi, xj = np.meshgrid(x, x)
    k = _b1(xi) * _b1(xj) + _b2(xi) * _b2(xj) - _b4(np.abs(xi - xj)) / 24.0
    lam, vec = np.linalg.eigh(k)
    order = np.argsort(lam)[::-1]
    lam = lam[order]
    vec = vec[:, order]
    n_orders = min(n_orders, n_grid)
    h = 1.0 / (n_grid - 1)
    out = np.zeros((n_orders, n_grid - 1, 4))
    for s in range(n_orders):
        f = vec[:, s] * np.sqrt(max(lam[s], 0.0))
        if s < len(_BERNOULLI_SIGN_AT_0):
            want_negative = _BERNOULLI_SIGN_AT_0[s] < 0
        else:
            want_negative = (s % 2 == 0)
        f0 = f[0] if f[0] != 0.0 else f[1]
        if (f0 < 0) != want_negative:
            f = -f
        cs = CubicSpline(x, f)                # cs.c[m, piece]: coefficient of (x - x_piece)**(3 - m)
        for p in range(4):
            out[s, :, p] = cs.c[3 - p, :] * h ** p
    return out


def to_phis(table):
    """[n][499][4] array -> reference `phis` layout: tuple of n lists of 4 arrays(499)."""
    return tuple([np.ascontiguousarray(table[s, :, p]) for p in range(4)] for s in range(table.shape[0]))


def from_phis(phis):
    """reference `phis` layout -> [n][499][4] float64 array."""
    n = len(phis)
    npiece = len(phis[0][0])
    out = np.zeros((n, npiece, 4))
    for s in range(n):
        for p in range(4):
            out[s, :, p] = np.asarray(phis[s][p], dtype=np.float64)
    return out


if __name__ == "__main__":
    import sys
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 48
    dst = sys.argv[2] if len(sys.argv) > 2 else "tests/golden/phis_cubic_%d.npy" % n
    tab = generate(n)
    np.save(dst, tab)
    print("saved", dst, tab.shape)
