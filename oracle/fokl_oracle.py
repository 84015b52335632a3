"""CPU oracle for the FoKL.fit hot path  --  TEST INFRASTRUCTURE, NOT PRODUCT CODE.

Only `tests/`, `__graft_entry__.smoke()` and `bench.py`'s cpu_baseline / `--impl reference` legs may
import this module.  The product (`fokl-gpy_b200/`) never does; it fails loudly without its CUDA library.

It restates, in numpy float64 (plus a small C helper for the basis loop), the algorithm of
`/root/reference/src/FoKL/FoKLRoutines.py` (abbreviated FR below):

    inputs_to_phind     FR:570-589      spline piece index / local coordinate
    eval_basis          FR:834-843      one basis function at one point (d = 0)
    basis_columns       FR:1446-1485    X[i, j] = prod_k phi_{d_jk}(x_ik)
    bss_derivatives     FR:594-805      partial derivatives of the fitted function (a section-8f "next" row)
    evaluate            FR:851-980      mean prediction and 95 % bounds from the draws (a section-8f "next" row)
    default_b_btau      FR:1322-1348    data-dependent defaults of b, btau
    gibbs               FR:1396-1558    Gram, eigh, betahat, Gibbs chain, BIC
    distinct_perms      FR:1350-1354 + FR:1616   == np.unique(perms(v), axis=0)
    fit                 FR:1561-1760    forward selection

Third-party arithmetic the reference leans on and that is restated by *calling the same library*:
numpy (`dot`, legacy `np.random.normal/gamma` on the global MT19937 state) and
`scipy.linalg.eigh` (LAPACK dsyevr).  Upstream pins no versions (pyproject.toml:17-23).

Parity pinning: `tests/test_oracle_golden.py` checks this module against (1) the BIC traces printed
in `examples/isotherm/isotherm_benchmark.ipynb:244-279, 453-459` and (2) outputs of the unmodified
reference run in the build container (`oracle/gen_golden.py` -> `tests/golden/*.npz`).
"""
import ctypes
import math
import os
import subprocess

import numpy as np
from scipy.linalg import eigh

_HERE = os.path.dirname(os.path.abspath(__file__))

CUBIC = 'Cubic Splines'
BERNOULLI = 'Bernoulli Polynomials'


# --------------------------------------------------------------------------------------------------
# basis functions
# --------------------------------------------------------------------------------------------------

def inputs_to_phind(x, n_piece=499):
    """FR:570-589.  x in [0, 1] -> (phind uint16 in [0, n_piece-1], xsm = n_piece*x - phind)."""
    x = np.asarray(x, dtype=np.float64)
    phind = np.array(np.ceil(x * n_piece), dtype=np.uint16)
    phind = phind + (phind == 0)
    phind = phind - 1
    xsm = np.array(n_piece * x - phind, dtype=x.dtype)
    return phind, xsm


def eval_basis(c, x, kernel):
    """FR:834-836 (cubic) and FR:841-843 (Bernoulli), d = 0, Python-scalar semantics (libm pow)."""
    if kernel == CUBIC:
        return c[0] + c[1] * x + c[2] * (x ** 2) + c[3] * (x ** 3)
    return c[0] + sum(c[k] * (x ** k) for k in range(1, len(c)))


def basis_columns_py(x, terms, phis, kernel):
    """Literal triple loop of FR:1446-1485 for the columns described by `terms` (C x M).

    Slow (Python); used for tiny cases and to pin the C helper."""
    x = np.asarray(x, dtype=np.float64)
    n, m = x.shape
    terms = np.asarray(terms)
    out = np.zeros((n, terms.shape[0]))
    if kernel == CUBIC:
        phind, xsm = inputs_to_phind(x, len(phis[0][0]))
    else:
        phind, xsm = None, x
    for i in range(n):
        for j in range(terms.shape[0]):
            phi = 1
            for k in range(m):
                num = terms[j][k]
                if num != 0:
                    nid = int(num - 1)
                    if kernel == CUBIC:
                        coeffs = [phis[nid][order][phind[i, k]] for order in range(4)]
                    else:
                        coeffs = phis[nid]
                    phi = phi * eval_basis(coeffs, xsm[i, k], kernel)
            out[i][j] = phi
    return out


_clib = None


def _load_clib():
    """Build (if needed) and load oracle/basis_oracle.c: the same loop in C with libm pow()."""
    global _clib
    if _clib is not None:
        return _clib
    so = os.path.join(_HERE, '_build', 'libbasis_oracle.so')
    src = os.path.join(_HERE, 'basis_oracle.c')
    if (not os.path.exists(so)) or os.path.getmtime(so) < os.path.getmtime(src):
        os.makedirs(os.path.dirname(so), exist_ok=True)
        subprocess.check_call(['gcc', '-O2', '-fPIC', '-shared', '-ffp-contract=off', '-fopenmp',
                               '-o', so, src, '-lm'])
    lib = ctypes.CDLL(so)
    lib.oracle_basis_columns.restype = ctypes.c_int
    lib.oracle_basis_columns.argtypes = [
        ctypes.c_void_p, ctypes.c_int64, ctypes.c_int,      # x (row-major N x M), N, M
        ctypes.c_void_p, ctypes.c_int,                      # terms int32 (C x M), C
        ctypes.c_int, ctypes.c_void_p, ctypes.c_int, ctypes.c_int,   # kernel, table, n_orders, row_len
        ctypes.c_void_p, ctypes.c_int]                      # out (row-major N x C), n_threads
    _clib = lib
    return lib


def phis_to_table(phis, kernel):
    """Pack the reference `phis` structure (GK:245-255 / GK:319-326) into a dense float64 table.

    cubic     -> [n_orders][499][4]          (order, piece, power)
    bernoulli -> [n_orders][n_orders + 1]    (row n uses its first n + 2 entries)"""
    n = len(phis)
    if kernel == CUBIC:
        npiece = len(phis[0][0])
        tab = np.zeros((n, npiece, 4))
        for s in range(n):
            for p in range(4):
                tab[s, :, p] = np.asarray(phis[s][p], dtype=np.float64)
        return tab
    width = max(len(r) for r in phis)
    tab = np.zeros((n, width))
    for s in range(n):
        tab[s, :len(phis[s])] = np.asarray(phis[s], dtype=np.float64)
    return tab


def basis_columns(x, terms, phis, kernel, table=None, threads=1):
    """Same result as `basis_columns_py` (bit-exact: same libm pow, no FMA contraction), in C."""
    lib = _load_clib()
    x = np.ascontiguousarray(x, dtype=np.float64)
    n, m = x.shape
    terms = np.ascontiguousarray(np.asarray(terms), dtype=np.int32).reshape(-1, m)
    if table is None:
        table = phis_to_table(phis, kernel)
    table = np.ascontiguousarray(table, dtype=np.float64)
    if kernel == CUBIC:
        kid, n_ord, row_len = 0, table.shape[0], table.shape[1]
    else:
        kid, n_ord, row_len = 1, table.shape[0], table.shape[1]
    if terms.size and terms.max() > n_ord:
        raise IndexError('term order exceeds len(phis)')
    out = np.zeros((n, terms.shape[0]))
    rc = lib.oracle_basis_columns(x.ctypes.data, n, m, terms.ctypes.data, terms.shape[0], kid,
                                  table.ctypes.data, n_ord, row_len, out.ctypes.data, threads)
    if rc != 0:
        raise ValueError('inputs are not normalised to [0, 1] (FR:590-591)')
    return out


# --------------------------------------------------------------------------------------------------
# bss_derivatives (FR:594-805)
# --------------------------------------------------------------------------------------------------

def twice_normalised(x, n_piece=499):
    """FR:570-586: the `X` output of _inputs_to_phind, (x - (phind - 1) * r) / r with r = 1 / n_piece and the
    1-based phind -- what bss_derivatives evaluates the cubic factors at (not xsm)."""
    x = np.asarray(x, dtype=np.float64)
    phind1 = np.array(np.ceil(x * n_piece), dtype=np.uint16)
    phind1 = phind1 + (phind1 == 0)
    r = 1 / n_piece
    xmin = np.array((phind1 - 1) * r, dtype=x.dtype)
    return (x - xmin) / r, phind1 - 1


def eval_basis_d(c, x, kernel, d):
    """FR:834-847 for d = 0, 1, 2 (Python / numpy scalar semantics)."""
    if d == 0:
        return eval_basis(c, x, kernel)
    if kernel == CUBIC:
        if d == 1:
            return c[1] + 2 * c[2] * x + 3 * c[3] * (x ** 2)
        return 2 * c[2] + 6 * c[3] * x
    if d == 1:
        return c[1] + sum(k * c[k] * (x ** (k - 1)) for k in range(2, len(c)))
    return sum((k - 1) * k * c[k] * (x ** (k - 2)) for k in range(2, len(c)))


def derivative_columns(x, terms, phis, kernel, wrt, order, span):
    """Design-matrix columns of d^order / d x_wrt^order of every term (FR:757-790): N x B, zero for the terms that do
    not contain input `wrt`.  `span` = max - min of that input's normalisation (FR:742-745)."""
    x = np.asarray(x, dtype=np.float64)
    n, m = x.shape
    terms = np.asarray(terms)
    if kernel == CUBIC:
        n_piece = len(phis[0][0])
        xe, phind = twice_normalised(x, n_piece)
        span_l = span / n_piece
    else:
        xe, phind, span_l = x, None, span / 1
    div = [1, span_l, span_l ** 2]
    out = np.zeros((n, terms.shape[0]))
    for i in range(n):
        for b in range(terms.shape[0]):
            if int(terms[b][wrt]) == 0:
                continue                              # FR:785-787: phi = 0
            phi = 1
            for k in range(m):
                num = int(terms[b][k])
                if num == 0:
                    continue
                if kernel == CUBIC:
                    c = [phis[num - 1][q][int(phind[i, k])] for q in range(4)]
                else:
                    c = phis[num - 1]
                if k == wrt:
                    phi *= eval_basis_d(c, xe[i, k], kernel, order) / div[order]
                else:
                    phi *= eval_basis_d(c, xe[i, k], kernel, 0)
            out[i, b] = phi
    return out


def bss_derivatives(inputs, betas, mtx, phis, kernel, minmax, d1=None, d2=None, draws=None,
                    individual_draws=False, full_array=False):
    """FR:594-805 with d1 / d2 given as boolean masks over the inputs (None: all first, no second derivatives).
    Returns what the reference returns (squeezed N x (#requested) [x draws] array, or N x M x 2 [x draws])."""
    inputs = np.asarray(inputs, dtype=np.float64)
    if inputs.ndim == 1:
        inputs = inputs[:, None]
    betas = np.asarray(betas, dtype=np.float64)
    mtx = np.asarray(mtx)
    n, m = inputs.shape
    nb = mtx.shape[0]
    if draws is None:
        draws = betas.shape[0]
    masks = [np.ones(m, bool) if d1 is None else np.asarray(d1, bool),
             np.zeros(m, bool) if d2 is None else np.asarray(d2, bool)]
    dy = np.zeros((draws, n, m, 2))
    for k in range(m):
        span = minmax[k][1] - minmax[k][0]
        for di in (0, 1):
            if not masks[di][k]:
                continue
            cols = derivative_columns(inputs, mtx, phis, kernel, k, di + 1, span)
            for b in range(nb):                       # FR:790: accumulated term by term
                dy[:, :, k, di] = dy[:, :, k, di] + betas[-draws:, b + 1][:, None] * cols[:, b][None, :]
    dy = np.transpose(dy, (1, 2, 3, 0))
    if not individual_draws and draws > 1:
        dy = np.mean(dy, axis=3)[:, :, :, None]
    if not full_array:
        dy = np.concatenate([dy[:, :, 0, :], dy[:, :, 1, :]], axis=1)
        dy = dy[:, ~np.all(dy == 0, axis=0)]
    return np.squeeze(dy)


# --------------------------------------------------------------------------------------------------
# evaluate (FR:851-980) -- the prediction path, a section-8f "next" row
# --------------------------------------------------------------------------------------------------

def evaluate(normputs, betas, mtx, phis, kernel, setnos, draws, return_bounds=False):
    """Mean prediction (and 95 % bounds) at already-normalised inputs from the draws `betas[setnos[:draws]]`:
    X as in FR:940-957, one matrix-vector product per draw (FR:960-962), bounds from the sorted draws (FR:966-972)."""
    normputs = np.asarray(normputs, dtype=np.float64)
    if normputs.ndim == 1:
        normputs = normputs[:, None]
    n = normputs.shape[0]
    X = np.hstack([np.ones((n, 1)), basis_columns(normputs, np.asarray(mtx).astype(int), phis, kernel)])
    if draws == 1:
        setnos = [0]
    modells = np.zeros((n, draws))
    for i in range(draws):
        modells[:, i] = np.transpose(np.matmul(X, np.transpose(np.array(betas[setnos[i], :]))))
    mean = np.mean(modells, 1)
    if not return_bounds:
        return mean
    bounds = np.zeros((n, 2))
    cut = int(np.floor(draws * 0.025) + 1)
    for i in range(n):
        drawset = np.sort(modells[i, :])
        bounds[i, 0] = drawset[cut]
        bounds[i, 1] = drawset[draws - cut]
    return mean, bounds


# --------------------------------------------------------------------------------------------------
# term generation
# --------------------------------------------------------------------------------------------------

def distinct_perms(v):
    """Distinct permutations of v in ascending lexicographic row order == np.unique(perms(v), axis=0)
    (FR:1353, FR:1616) without enumerating len(v)! rows."""
    a = sorted(float(t) for t in v)
    n = len(a)
    rows = [list(a)]
    while True:
        i = n - 2
        while i >= 0 and a[i] >= a[i + 1]:
            i -= 1
        if i < 0:
            break
        j = n - 1
        while a[j] <= a[i]:
            j -= 1
        a[i], a[j] = a[j], a[i]
        a[i + 1:] = a[i + 1:][::-1]
        rows.append(list(a))
    return np.array(rows, dtype=np.float64)


def initial_indvec(ind, m, sett):
    """FR:1605-1613: spread `ind` ones round-robin over the first `sett` slots."""
    v = np.zeros(m)
    summ = ind
    while summ:
        for j in range(sett):
            v[j] += 1
            summ -= 1
            if summ == 0:
                break
    return v


def advance_indvec(v, m, way3):
    """FR:1722-1740.  Returns False when the partition walk of this `ind` is over."""
    if m == 1:
        return False
    if way3:
        if v[1] > v[2]:
            v[0] += 1
            v[1] -= 1
        elif v[2]:
            v[1] += 1
            v[2] -= 1
            if v[1] > v[0]:
                v[0] += 1
                v[1] -= 1
        else:
            return False
        return True
    if v[1]:
        v[0] += 1
        v[1] -= 1
        return True
    return False


# --------------------------------------------------------------------------------------------------
# hyper-parameter defaults
# --------------------------------------------------------------------------------------------------

def default_b_btau(data, a, atau, b=None, btau=None):
    """FR:1322-1348 for float64 data."""
    if b is None or btau is None:
        sigmasq = np.var(data)
        data_mean = np.mean(data)
        if b is None:
            b = sigmasq * (a + 1)
        if btau is None:
            btau = (np.abs(data_mean) / sigmasq) * (atau + 1)
    return b, btau


# --------------------------------------------------------------------------------------------------
# gibbs
# --------------------------------------------------------------------------------------------------

def draw_variates(p, draws, astar, atau_star):
    """Consume the global legacy numpy RNG exactly as one `gibbs` call does (FR:1527, 1541, 1547):
    per draw normal(size=(p, 1)), then standard_gamma(astar), then standard_gamma(atau_star)
    (np.random.gamma(k, s) == s * standard_gamma(k) bitwise).  Assumes bstar >= 0 throughout."""
    z = np.empty((draws, p))
    g1 = np.empty(draws)
    g2 = np.empty(draws)
    for k in range(draws):
        z[k] = np.random.normal(loc=0, scale=1, size=(p, 1))[:, 0]
        g1[k] = np.random.standard_gamma(astar)
        g2[k] = np.random.standard_gamma(atau_star)
    return z, g1, g2


def gibbs_from_X(X, data, a, b, atau, btau, draws, sigsqd, tausqd, dtd, literal=True, variates=None, gram=None):
    """FR:1492-1558 given the finished design matrix X (N x P, column 0 = ones).

    literal=True  : dense products per draw exactly as the reference writes them (bit-exact on the
                    same BLAS); draws come from the global numpy RNG.
    literal=False : eigenbasis form, O(p) per draw; `variates=(z, g1, g2)` may be injected.
    gram=(XtX, Xty): parity-harness hook -- use these Gram bits (e.g. the device's) instead of X'X, X'y, so
                    that LAPACK's eigenvector signs are those of the matrix the device factorised.
    Returns dict(betas, sigs, taus, betahat, ev, XtX, Xty, Lamb, Q)."""
    mmtx = X.shape[1] - 1
    if gram is None:
        XtX = np.transpose(X).dot(X)
        Xty = np.transpose(X).dot(data)
    else:
        XtX = np.array(gram[0], dtype=np.float64)
        Xty = np.array(gram[1], dtype=np.float64).reshape(-1, 1)
    Lamb, Q = eigh(XtX)
    Lamb_inv = np.diag(1 / Lamb)
    betahat = Q.dot(Lamb_inv).dot(np.transpose(Q)).dot(Xty)

    n = len(data)
    astar = a + 1 + n / 2 + (mmtx + 1) / 2
    atau_star = atau + mmtx / 2

    betas = np.zeros((draws, mmtx + 1))
    sigs = np.zeros((draws, 1))
    taus = np.zeros((draws, 1))

    if literal:
        for k in range(draws):
            Lamb_tausqd = np.diag(Lamb) + (1 / tausqd) * np.identity(mmtx + 1)
            Lamb_tausqd_inv = np.diag(1 / np.diag(Lamb_tausqd))
            mun = Q.dot(Lamb_tausqd_inv).dot(np.transpose(Q)).dot(Xty)
            S = Q.dot(np.diag(np.diag(Lamb_tausqd_inv) ** (1 / 2)))
            vec = np.random.normal(loc=0, scale=1, size=(mmtx + 1, 1))
            betas[k][:] = np.transpose(mun + sigsqd ** (1 / 2) * (S).dot(vec))
            bk = betas[k][:]
            bstar = b + 0.5 * (bk.dot(XtX.dot(np.transpose([bk]))) - 2 * bk.dot(Xty) + dtd +
                               bk.dot(np.transpose([bk])) / tausqd)
            if bstar < 0:
                sigsqd = math.nan
            else:
                sigsqd = 1 / np.random.gamma(astar, 1 / bstar)
            sigs[k] = sigsqd
            btau_star = (1 / (2 * sigsqd)) * (bk.dot(np.reshape(bk, (len(bk), 1)))) + btau
            tausqd = 1 / np.random.gamma(atau_star, 1 / btau_star)
            taus[k] = tausqd
    else:
        p = mmtx + 1
        if variates is None:
            variates = draw_variates(p, draws, astar, atau_star)
        z, g1, g2 = variates
        ct = np.transpose(Q).dot(Xty)[:, 0]
        gam = np.zeros((draws, p))
        dtd_s = float(np.asarray(dtd).reshape(-1)[0])
        sig = float(sigsqd)
        tau = float(tausqd)
        for k in range(draws):
            d = 1 / (Lamb + 1 / tau)
            g = d * ct + sig ** (1 / 2) * (d ** (1 / 2)) * z[k]
            gam[k] = g
            s1 = np.sum(Lamb * g * g)
            s2 = np.sum(g * ct)
            s3 = np.sum(g * g)
            bstar = b + 0.5 * (s1 - 2 * s2 + dtd_s + s3 / tau)
            if bstar < 0:
                sig = math.nan
            else:
                sig = 1 / ((1 / bstar) * g1[k])
            sigs[k] = sig
            btau_star = (1 / (2 * sig)) * s3 + btau
            tau = 1 / ((1 / btau_star) * g2[k])
            taus[k] = tau
        betas = gam.dot(np.transpose(Q))

    siglik = np.var(data - np.matmul(X, betahat))
    lik = -(n / 2) * np.log(siglik) - (n - 1) / 2
    ev = (mmtx + 1) * np.log(n) - 2 * np.max(lik)
    return dict(betas=betas, sigs=sigs, taus=taus, betahat=betahat, ev=ev, XtX=XtX, Xty=Xty,
                Lamb=Lamb, Q=Q)


def bic_from_gram(G, Xty, yty, sum_y, n):
    """BIC of the OLS fit from Gram quantities only: column 0 of X is ones so
    sum(r) = sum(y) - G[0, :] betahat and sum(r^2) = yty - 2 b'Xty + b'G b."""
    Lamb, Q = eigh(G)
    betahat = Q.dot(np.diag(1 / Lamb)).dot(Q.T).dot(Xty)
    sr = sum_y - G[0, :].dot(betahat)
    srr = yty - 2 * betahat.dot(Xty) + betahat.dot(G.dot(betahat))
    siglik = srr / n - (sr / n) ** 2
    p = G.shape[0]
    return p * np.log(n) - 2 * (-(n / 2) * np.log(siglik) - (n - 1) / 2)


# --------------------------------------------------------------------------------------------------
# fit
# --------------------------------------------------------------------------------------------------

class FitResult(dict):
    __getattr__ = dict.__getitem__


def fit(inputs, data, phis, kernel=CUBIC, a=4, b=None, atau=4, btau=None, tolerance=3, burnin=1000,
        draws=1000, gimmie=False, way3=False, threshav=0.05, threshstda=0.5, threshstdb=2, aic=False,
        basis='c', literal=True, threads=1, on_gibbs=None, console=False, gram_hook=None, on_substage=None):
    """Forward selection of FR:1561-1760 on already-normalised `inputs` (N x M in [0, 1]) and
    `data` (N x 1).  relats_in = [] only (anything else crashes upstream at FR:1631).

    basis: 'c' (C helper) or 'py' (literal Python triple loop, the reference's real cost profile).
    on_gibbs(info): optional callback per `gibbs` invocation (for recording golden vectors / timing).
    on_substage(ind, ev): optional callback at the end of every substage (where the reference prints [ind, ev]).
    gram_hook(discmtx) -> (XtX, Xty) or None: parity-harness hook, see gibbs_from_X.
    Returns FitResult(betas, mtx, evs, n_gibbs, betas_full)."""
    inputs = np.asarray(inputs, dtype=np.float64)
    data = np.asarray(data, dtype=np.float64).reshape(-1, 1)
    n, m = inputs.shape
    b, btau = default_b_btau(data, a, atau, b, btau)
    total = burnin + draws
    sigsqd0 = b / (1 + a)
    tausqd0 = btau / (1 + atau)
    dtd = np.transpose(data).dot(data)
    table = phis_to_table(phis, kernel)
    build = basis_columns_py if basis == 'py' else \
        (lambda x, t, ph, k: basis_columns(x, t, ph, k, table=table, threads=threads))
    counter = [0]

    def gibbs(Xin, discmtx):
        # FR:1435-1485: only columns nxin .. mmtx are (re)built
        mmtx = discmtx.shape[0]
        if np.size(Xin) == 0:
            Xin = np.ones((n, 1))
        nxin = Xin.shape[1]
        if mmtx - nxin < 0:
            X = Xin
        else:
            newcols = build(inputs, discmtx[nxin - 1:mmtx], phis, kernel)
            X = np.append(Xin, newcols, axis=1)
        gram = gram_hook(discmtx) if gram_hook is not None else None
        r = gibbs_from_X(X, data, a, b, atau, btau, total, sigsqd0, tausqd0, dtd, literal=literal, gram=gram)
        counter[0] += 1
        X = X[:, 0:mmtx + 1]
        if on_gibbs is not None:
            on_gibbs(dict(call=counter[0], discmtx=discmtx.copy(), X=X, **r))
        return r['betas'], X, r['ev']

    damtx = np.array([])
    evs = np.array([])
    ind = 1
    greater = 0
    finished = 0
    X = []
    sett = 1 if m == 1 else (3 if way3 else 2)
    h0 = int(np.ceil(total / 2))
    h1 = int(np.ceil((total / 2) + 1))
    h1b = int(np.ceil(total / 2) + 1)

    while True:
        indvec = initial_indvec(ind, m, sett)
        while True:
            vecs = distinct_perms(indvec)
            vm = vecs.shape[0]
            damtx = vecs if np.size(damtx) == 0 else np.append(damtx, vecs, axis=0)
            dam = damtx.shape[0]

            beters, xers, ev = gibbs(X, damtx)
            if aic:
                ev = ev + (2 - np.log(n)) * (dam + 1)

            # FR:1656-1664
            betavs = np.abs(np.mean(beters[h1:total, (dam - vm + 1):dam + 1], axis=0))
            betavs2 = np.divide(np.std(np.array(beters[h1b:total, dam - vm + 1:dam + 1]), axis=0),
                                np.abs(np.mean(beters[h0:total, dam - vm + 1:dam + 2], axis=0)))
            betavs3 = np.array(range(dam - vm + 2, dam + 2))
            betavs = np.transpose(np.array([betavs, betavs2, betavs3]))
            if np.shape(betavs)[1] > 0:
                betavs = betavs[np.argsort(betavs[:, 0])]

            killset = []
            evmin = ev
            for i in range(0, vm):
                if betavs[i, 1] > threshstdb or betavs[i, 1] > threshstda and betavs[i, 0] < threshav * \
                        np.mean(np.abs(np.mean(beters[h0:total, 0]))):
                    killtest = np.append(killset, (betavs[i, 2] - 1))
                    if killtest.size > 1:
                        killtest[::-1].sort()
                    damtx_test = damtx
                    for k in range(0, np.size(killtest)):
                        damtx_test = np.delete(damtx_test, int(np.array(killtest[k]) - 1), 0)
                    damtest = damtx_test.shape[0]
                    betertest, Xtest, evtest = gibbs(X, damtx_test)
                    if aic:
                        evtest = evtest + (2 - np.log(n)) * (damtest + 1)
                    if evtest < evmin:
                        killset = killtest
                        evmin = evtest
                        xers = Xtest
                        beters = betertest
            for k in range(0, np.size(killset)):
                damtx = np.delete(damtx, int(np.array(killset[k]) - 1), 0)

            ev = evmin
            X = xers
            if console:
                print([ind, float(ev)])
            if on_substage is not None:
                on_substage(ind, float(ev))
            if np.size(evs) > 0:
                if ev < np.min(evs):
                    betas = beters
                    mtx = damtx
                    greater = 1
                    evs = np.append(evs, ev)
                elif greater < tolerance:
                    greater = greater + 1
                    evs = np.append(evs, ev)
                else:
                    finished = 1
                    evs = np.append(evs, ev)
                    break
            else:
                greater = greater + 1
                betas = beters
                mtx = damtx
                evs = np.append(evs, ev)
            if not advance_indvec(indvec, m, way3):
                break
        if finished != 0:
            break
        ind = ind + 1
        if ind > len(phis):
            break

    if gimmie:
        betas = beters
        mtx = damtx
    return FitResult(betas=betas[-draws::, :], mtx=mtx, evs=evs, n_gibbs=counter[0], betas_full=betas)


def normalize(inputs, minmax=None):
    """FR:373-377, 436-437: per-column (x - min) / (max - min)."""
    x = np.array(inputs, dtype=np.float64)
    if x.ndim == 1:
        x = x[:, None]
    if minmax is None:
        minmax = [[np.min(x[:, j]), np.max(x[:, j])] for j in range(x.shape[1])]
    for j in range(x.shape[1]):
        x[:, j] = (x[:, j] - minmax[j][0]) / (minmax[j][1] - minmax[j][0])
    return x, minmax
