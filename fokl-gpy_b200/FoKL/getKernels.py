"""Basis-function coefficient tables (`phis`) in the reference's layout.

Mirrors the interface of the reference's src/FoKL/getKernels.py (`sp500()` :221-267, `bernoulli()` :308-326):
    sp500()     -> tuple of 500 lists [c0, c1, c2, c3], each an array over 499 spline pieces
    bernoulli() -> tuple of 20 lists; entry n holds the n + 2 monomial coefficients of polynomial n + 1
"""
import os
import warnings

import numpy as np

_HERE = os.path.dirname(os.path.realpath(__file__))
_KDIR = os.path.join(_HERE, 'kernels')
UPSTREAM_SPLINE_FILE = 'splineCoefficient500_highPrecision_smoothed.txt'


def _to_phis(table):
    return tuple([np.ascontiguousarray(table[s, :, k]) for k in range(4)] for s in range(table.shape[0]))


def _b1(x):
    return x - 0.5


def _b2(x):
    return x ** 2 - x + 1.0 / 6.0


def _b4(x):
    return x ** 4 - 2 * x ** 3 + x ** 2 - 1.0 / 30.0


_SIGN_AT_0 = (-1, 1, -1, 1, 1, -1, 1, -1, 1, -1, 1, -1, 1, 1, 1, -1, -1, 1, -1, 1)


def regenerate_spline_table(n_orders=500, n_grid=500):
    """Rebuild the cubic-spline table from its published derivation (the upstream data file is a large
    blob that is not always present): eigendecompose the BSS-ANOVA kernel core
    kappa_1(x, x') = B1 B1' + B2 B2' - B4(|x - x'|) / 24 on linspace(0, 1, 500), scale the eigenvectors by
    sqrt(eigenvalue), fit cubic splines, and express each piece in the local coordinate t = 499 x - piece.
    Returns [n_orders][n_grid - 1][4]."""
    from scipy.interpolate import CubicSpline
    x = np.linspace(0.0, 1.0, n_grid)
    xi, xj = np.meshgrid(x, x)
    k = _b1(xi) * _b1(xj) + _b2(xi) * _b2(xj) - _b4(np.abs(xi - xj)) / 24.0
    lam, vec = np.linalg.eigh(k)
    order = np.argsort(lam)[::-1]
    lam, vec = lam[order], vec[:, order]
    n_orders = min(n_orders, n_grid)
    h = 1.0 / (n_grid - 1)
    out = np.zeros((n_orders, n_grid - 1, 4))
    for s in range(n_orders):
        f = vec[:, s] * np.sqrt(max(lam[s], 0.0))
        neg = (_SIGN_AT_0[s] < 0) if s < len(_SIGN_AT_0) else (s % 2 == 0)
        f0 = f[0] if f[0] != 0.0 else f[1]
        if (f0 < 0) != neg:
            f = -f
        cs = CubicSpline(x, f)
        for q in range(4):
            out[s, :, q] = cs.c[3 - q, :] * h ** q
    return out


def sp500(**kwargs):
    """Return 'phis', a [500 x 4 x 499] tuple of lists of double-precision cubic-spline coefficients.

    If the upstream file kernels/splineCoefficient500_highPrecision_smoothed.txt is present it is loaded
    exactly as the reference does; otherwise a table regenerated from the kernel's definition is used (cached
    as kernels/spline500_regenerated.npy) and a UserWarning says so."""
    for kw in kwargs:
        if kw not in ('Smooth', 'Save'):
            raise ValueError(f"Unexpected keyword argument: {kw}")
    smooth, save = kwargs.get('Smooth', 0), kwargs.get('Save', 0)
    if save == 1 and smooth == 0:
        warnings.warn("The spline coefficients were not save because they were not requested to be smoothed.",
                      category=UserWarning)
    if smooth:
        # GK:10-218 is a development tool that re-smooths the end pieces of the (already smoothed) shipped table; it is
        # not on the fit path (SURVEY 8a row a1 = the loader, GK:245-255) and is not rebuilt here
        raise NotImplementedError("sp500(Smooth=1) (getKernels.smooth_coefficients, a development option upstream) is "
                                  "not part of the B200 build; the table is returned as stored with Smooth=0.")
    upstream = os.path.join(_KDIR, UPSTREAM_SPLINE_FILE)
    if os.path.exists(upstream):
        raw = np.loadtxt(upstream, delimiter=',', dtype=np.double)
        table = np.transpose(raw.reshape(500, 499, 4), (0, 1, 2)).copy()
        return _to_phis(table)
    cache = os.path.join(_KDIR, 'spline500_regenerated.npy')
    if os.path.exists(cache):
        table = np.load(cache)
    else:
        table = regenerate_spline_table(500)
        try:
            os.makedirs(_KDIR, exist_ok=True)
            np.save(cache, table)
        except OSError:
            pass
    warnings.warn("Upstream spline coefficient file not found; using a table regenerated from the BSS-ANOVA "
                  "kernel definition (not bit-identical to upstream's smoothed file).", category=UserWarning)
    return _to_phis(table)


def smooth_coefficients(phis):
    """Upstream's end-piece smoothing of a spline table (GK:10-218): a development tool, not rebuilt (see sp500)."""
    raise NotImplementedError("getKernels.smooth_coefficients is a development option upstream and is not part of the "
                              "B200 build.")


def bss_anova(n=500):
    """Development helper kept for interface parity (GK:270-305): writes the square roots of the eigenvalues of the
    n x n BSS-ANOVA kernel matrix, largest first, to 'BSS-ANOVA__sqrt-eigvals__K-500x500.txt' in the working directory
    (upstream used them to scale the orthonormal Bernoulli polynomials) and returns None."""
    x = np.linspace(0.0, 1.0, n)
    xi, xj = np.meshgrid(x, x)
    k = _b1(xi) * _b1(xj) + _b2(xi) * _b2(xj) - _b4(np.abs(xi - xj)) / 24
    lam = np.linalg.eigvalsh(k)
    np.savetxt("BSS-ANOVA__sqrt-eigvals__K-500x500.txt", np.flip(np.sqrt(lam)), delimiter=",")


def bernoulli(file='orthogonal_Bn_scaled.txt'):
    """Return coefficients of the scaled orthonormal Bernoulli polynomials (20 polynomials; GK:308-326).  The table ships
    as kernels/orthogonal_Bn_scaled.npy -- the float64 values `np.loadtxt` reads from upstream's text file -- and is used
    when the named text file is not present next to it."""
    path = os.path.join(_KDIR, file)
    if not os.path.exists(path) and os.path.exists(os.path.splitext(path)[0] + '.npy'):
        path = os.path.splitext(path)[0] + '.npy'
    if path.endswith('.npy'):
        coeffs = np.load(path)
    else:
        coeffs = np.loadtxt(path, delimiter=' ', dtype=np.double)
    return tuple(list(coeffs[n, :(n + 2)]) for n in range(coeffs.shape[0]))
