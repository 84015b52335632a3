"""Small whole fits with unusual hyper-parameters, runnable on the unmodified reference (this file as a script, with the
reference first on sys.path) and on this package (imported by tests/test_reference_differential.py, which runs them on
the CPU stand-in engine).  TEST INFRASTRUCTURE."""
import hashlib
import os
import pickle
import sys

import numpy as np

# name -> (n, m, data seed, constructor keywords, fit keywords)
CASES = {
    'defaults_m2': (90, 2, 1, dict(), dict()),
    'one_input': (70, 1, 2, dict(), dict()),
    'way3_m3': (110, 3, 3, dict(way3=True), dict()),
    'way3_m4_tol1': (120, 4, 4, dict(way3=True, tolerance=1), dict()),
    'aic': (80, 2, 5, dict(aic=True), dict()),
    'gimmie': (80, 2, 6, dict(gimmie=True), dict()),
    'gimmie_way3_aic': (100, 3, 7, dict(gimmie=True, way3=True, aic=True), dict()),
    'tolerance5': (80, 2, 8, dict(tolerance=5), dict()),
    'loose_thresholds': (90, 3, 9, dict(threshav=0.5, threshstda=0.1, threshstdb=0.8), dict()),
    'tight_thresholds': (90, 3, 10, dict(threshav=0.0, threshstda=5.0, threshstdb=50.0), dict()),
    'strong_prior': (80, 2, 11, dict(a=50, b=2.0, atau=9, btau=3.0), dict()),
    'weak_prior': (80, 2, 12, dict(a=1.5, atau=0.5), dict()),
    'hypers_in_fit': (80, 2, 13, dict(), dict(a=6, atau=2, tolerance=2, way3='yes', aic='on')),
    'odd_draws': (80, 2, 14, dict(draws=17, burnin=4), dict()),
    'burnin_zero': (80, 2, 15, dict(draws=21, burnin=0), dict()),
    'minmax_pillow': (80, 2, 16, dict(), dict(minmax=[[-0.5, 1.5], [0.0, 2.0]], pillow=0.05)),
    'train_half': (140, 2, 17, dict(), dict(train=0.5)),
    'noisy_m4': (100, 4, 18, dict(), dict()),
    # cubic splines (the regenerated table injected through `phis=`, as everywhere in this repo)
    'cubic_m2': (90, 2, 19, dict(cubic=True), dict()),
    'cubic_way3_m3_aic': (100, 3, 20, dict(cubic=True, way3=True, aic=True), dict()),
    'cubic_one_input_tol4': (60, 1, 21, dict(cubic=True, tolerance=4), dict()),
    'cubic_gimmie_loose': (90, 2, 22, dict(cubic=True, gimmie=True, threshstda=0.2, threshav=0.3), dict()),
    'way3_m5': (130, 5, 23, dict(way3=True), dict()),
    'two_way_m5_tol2': (130, 5, 24, dict(tolerance=2), dict()),
    # containers and call sequences of the public API
    'inputs_as_list_of_columns': (90, 2, 25, dict(container='list'), dict()),
    'pandas_frame_and_series': (90, 3, 26, dict(container='pandas'), dict()),
    'refit_same_model_on_other_data': (80, 2, 27, dict(container='refit'), dict()),
    'clean_first_then_fit_without_arguments': (80, 2, 28, dict(container='preclean'), dict(pillow=0.1)),
}


if os.environ.get('FOKL_DIFF_CASES'):          # exploratory runs (tools/diff_fuzz.py): another case list, same format
    with open(os.environ['FOKL_DIFF_CASES'], 'rb') as _f:
        CASES = pickle.load(_f)


def cubic_phis():
    here = os.path.dirname(os.path.abspath(__file__))
    sys.path.insert(0, os.path.join(here, '..', '..', 'oracle'))
    import spline_table
    return spline_table.to_phis(np.load(os.path.join(here, '..', 'golden', 'phis_cubic_48.npy')))


def data(n, m, seed):
    rng = np.random.default_rng(100 + seed)
    x = rng.random((n, m)) * 3.0 - 1.0                    # raw, un-normalised inputs
    u = (x + 1.0) / 3.0
    y = np.sin(2 * np.pi * u[:, 0]) + 0.3 * rng.standard_normal(n) * (0.2 if seed != 18 else 1.0)
    if m > 1:
        y = y + u[:, 0] * u[:, 1]
    if m > 2:
        y = y + (u[:, 2] - 0.5) ** 2
    return x, y


def digest():
    st = np.random.get_state()
    return hashlib.sha256(st[1].tobytes() + bytes(str((st[2], st[3], repr(st[4]))), 'ascii')).hexdigest()


def run_case(FR, name):
    n, m, seed, ckw, fkw = CASES[name]
    x, y = data(n, m, seed)
    ckw = dict(dict(draws=30, burnin=30), **ckw)
    kern = dict(phis=cubic_phis()) if ckw.pop('cubic', False) else dict(kernel=1)
    container = ckw.pop('container', None)
    np.random.seed(seed)
    model = FR.FoKL(UserWarnings=False, ConsoleOutput=False, **kern, **ckw)
    try:
        if container == 'list':
            betas, mtx, evs = model.fit([x[:, k] for k in range(m)], list(y), clean=True, **fkw)
        elif container == 'pandas':
            import pandas as pd
            frame = pd.DataFrame({'c%d' % k: x[:, k] for k in range(m)})
            betas, mtx, evs = model.fit(frame, pd.Series(y), clean=True, **fkw)
        elif container == 'refit':
            model.fit(x[:40], y[:40], clean=True, **fkw)
            betas, mtx, evs = model.fit(x[40:], y[40:], clean=True, **fkw)
        elif container == 'preclean':
            model.clean(x, y, _setattr=True, **fkw)
            betas, mtx, evs = model.fit()
        else:
            betas, mtx, evs = model.fit(x, y, clean=True, **fkw)
    except Exception as exc:  # noqa: BLE001 -- a fit the reference cannot finish: the exception type is the outcome
        return dict(raised=type(exc).__name__)
    return dict(betas_shape=tuple(betas.shape), mtx=np.asarray(mtx, dtype=np.float64), evs=np.asarray(evs, dtype=np.float64),
                digest=digest(), inputs=np.asarray(model.inputs), data=np.asarray(model.data), b=float(model.b),
                btau=float(model.btau), minmax=np.asarray(model.minmax, dtype=np.float64),
                betas_mean=betas.mean(axis=0), betas=betas,
                trainlog=None if model.trainlog is None else np.asarray(model.trainlog))


if __name__ == '__main__':
    from FoKL import FoKLRoutines as FR
    out = {name: run_case(FR, name) for name in CASES}
    with open(sys.argv[1], 'wb') as f:
        pickle.dump(dict(file=FR.__file__, fits=out), f)
