"""Update fits (`update=True`, FR:1365-1367 -> fitupdate FR:1850-2583) driven like examples/sigmoid/updateSig.py:54-118,
runnable on the unmodified reference (this file as a script) and on this package (imported by
tests/test_reference_differential.py, CPU stand-in engine).  TEST INFRASTRUCTURE."""
import pickle
import sys
import warnings

import numpy as np

from fit_cases import cubic_phis, digest

# name -> (inputs, rows per batch, seed, constructor keywords, burn)
CASES = {
    'bern_m2': (2, 160, 31, dict(sigsqd0=0.01, a=9, b=0.01, atau=3, btau=4000, draws=80, burnin=0), 30),
    'cubic_m2_aic': (2, 200, 32, dict(cubic=True, sigsqd0=0.02, a=9, b=0.01, atau=3, btau=4000, aic=True, draws=90, burnin=10), 40),
    'bern_m3_tol2': (3, 220, 33, dict(sigsqd0=0.02, a=6, b=0.05, atau=3, btau=2000, tolerance=2, draws=70, burnin=20), 30),
    'cubic_m3_gimmie': (3, 180, 34, dict(cubic=True, sigsqd0=0.01, a=9, b=0.01, atau=3, btau=4000, gimmie=True, draws=60, burnin=0), 20),
    'bern_m2_tol1': (2, 150, 35, dict(sigsqd0=0.05, a=4, b=0.1, atau=4, btau=1000, tolerance=1, draws=60, burnin=5), 25),
}


def surface(x):
    y = 1 / ((1 + np.exp(-5 * x[:, 0] + 2.5)) * (1 + np.exp(-5 * x[:, 1] + 2.5)))
    if x.shape[1] > 2:
        y = y + 0.3 * np.sin(3 * x[:, 2]) * x[:, 0]
    return y


def run_case(FR, name, n_fits=2):
    m, nb, seed, ckw, burn = CASES[name]
    ckw = dict(ckw)
    rng = np.random.default_rng(200 + seed)
    x = rng.random((nb * n_fits, m))
    y = (surface(x) + 0.01 * rng.standard_normal(len(x)))[:, None]
    kern = dict(phis=cubic_phis()) if ckw.pop('cubic', False) else dict(kernel=1)
    out = []
    with warnings.catch_warnings():
        warnings.simplefilter('ignore')
        np.random.seed(seed)
        model = FR.FoKL(UserWarnings=False, ConsoleOutput=False, **kern, **ckw)
        model.update, model.built, model.burn = True, False, burn
        for f in range(n_fits):
            lo, hi = f * nb, (f + 1) * nb
            if f == 0:
                model.clean(x[lo:hi], y[lo:hi], minmax=[[0, 1]] * m)
            else:
                model.data = y[lo:hi]
                model.inputs = model.clean(x[lo:hi])
            try:
                betas, mtx, evs = model.fit()
            except Exception as exc:  # noqa: BLE001
                out.append(dict(raised=type(exc).__name__))
                break
            out.append(dict(mtx=np.asarray(mtx, dtype=np.float64), evs=np.asarray(evs, dtype=np.float64),
                            betas_shape=tuple(np.shape(betas)), betas_type=type(betas).__name__, built=bool(model.built),
                            digest=digest(), betas_mean=np.asarray(betas).mean(axis=0)))
    return out


if __name__ == '__main__':
    from FoKL import FoKLRoutines as FR
    with open(sys.argv[1], 'wb') as f:
        pickle.dump(dict(file=FR.__file__, fits={name: run_case(FR, name) for name in CASES}), f)
