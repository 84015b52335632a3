"""Host-only API calls run identically on either FoKL package (this repo's or the unmodified reference): the driver of
tests/test_reference_differential.py.  TEST INFRASTRUCTURE.

    python host_api_cases.py <out.pkl>      # with the package under test first on sys.path (PYTHONPATH)

Every case is a small function of (FR, rng); its outcome -- returned arrays, attributes, warnings' categories or the
exception type -- is recorded in a plain dict and pickled."""
import pickle
import sys
import warnings

import numpy as np


def _phis(FR):
    from FoKL import getKernels
    return getKernels.bernoulli()


def _model(FR, **kw):
    return FR.FoKL(kernel=1, UserWarnings=kw.pop('UserWarnings', False), ConsoleOutput=False, **kw)


def _state(m):
    out = {}
    for k in ('inputs', 'data', 'minmax', 'trainlog'):
        v = getattr(m, k, 'ABSENT')
        out[k] = None if v is None else (v if isinstance(v, str) else np.asarray(v).tolist())
        if v is not None and not isinstance(v, str):
            out[k + '_dtype'] = str(np.asarray(v).dtype)
    return out


def cases():
    c = {}

    def case(fn):
        c[fn.__name__] = fn
        return fn

    # ---- clean / _format: input containers and shapes -------------------------------------------------------------
    @case
    def clean_list_of_columns(FR, rng):
        x = rng.random((30, 3))
        m = _model(FR)
        out = m.clean([x[:, 0], x[:, 1], x[:, 2]], rng.random(30), _setattr=True)
        return [np.asarray(o).tolist() for o in out], _state(m)

    @case
    def clean_wide_array_is_transposed(FR, rng):
        x = rng.random((3, 40))
        m = _model(FR)
        out = m.clean(x, rng.random(40), _setattr=True)
        return [np.asarray(o).shape for o in out], _state(m)

    @case
    def clean_no_autotranspose(FR, rng):
        x = rng.random((3, 40))
        m = _model(FR)
        out = m.clean(x, rng.random(3), AutoTranspose=False, _setattr=True)
        return [np.asarray(o).shape for o in out], _state(m)

    @case
    def clean_single_instance(FR, rng):
        m = _model(FR)
        m.clean(rng.random((25, 4)), rng.random(25), _setattr=True)
        out = m.clean(rng.random(4), SingleInstance=True, minmax=m.minmax)
        return np.asarray(out).tolist()

    @case
    def clean_pandas(FR, rng):
        import pandas as pd
        df = pd.DataFrame({'a': rng.random(20) * 7 - 3, 'b': rng.random(20) * 100, 'y': rng.standard_normal(20)})
        m = _model(FR)
        out = m.clean(df[['a', 'b']], df['y'], _setattr=True)
        return [np.asarray(o).tolist() for o in out], _state(m)

    @case
    def clean_integers_are_converted(FR, rng):
        x = rng.integers(0, 50, size=(20, 2))
        m = _model(FR)
        out = m.clean(x, rng.integers(0, 9, size=20), _setattr=True)
        return [np.asarray(o).tolist() for o in out], _state(m)

    @case
    def clean_3d_inputs_squeezed(FR, rng):
        m = _model(FR)
        out = m.clean(rng.random((15, 2, 1)), rng.random((15, 1, 1)), _setattr=True)
        return [np.asarray(o).tolist() for o in out], _state(m)

    @case
    def clean_data_row_vector(FR, rng):
        m = _model(FR)
        out = m.clean(rng.random((12, 2)), rng.random((1, 12)), _setattr=True)
        return [np.asarray(o).shape for o in out], _state(m)

    @case
    def clean_data_matrix_raises(FR, rng):
        m = _model(FR)
        return m.clean(rng.random((12, 2)), rng.random((12, 2)))

    @case
    def clean_unknown_keyword_raises(FR, rng):
        m = _model(FR)
        return m.clean(rng.random((12, 2)), rng.random(12), nonsense=True)

    @case
    def clean_bad_bit_falls_back(FR, rng):
        m = _model(FR)
        out = m.clean(rng.random((12, 2)), rng.random(12), bit=8, _setattr=True)
        return [np.asarray(o).tolist() for o in out], _state(m)

    @case
    def clean_bit32(FR, rng):
        m = _model(FR)
        out = m.clean(rng.random((12, 2)), rng.random(12), bit=32, _setattr=True)
        return [np.asarray(o).tolist() for o in out], _state(m)

    # ---- normalisation keywords ------------------------------------------------------------------------------------
    @case
    def clean_normalize_off(FR, rng):
        m = _model(FR)
        out = m.clean(rng.random((12, 2)) * 5, rng.random(12), normalize=False, _setattr=True)
        return [np.asarray(o).tolist() for o in out], _state(m)

    @case
    def clean_normalize_strings(FR, rng):
        m = _model(FR)
        out = m.clean(rng.random((12, 2)) * 5, rng.random(12), normalize='off', _setattr=True)
        return [np.asarray(o).tolist() for o in out], _state(m)

    @case
    def clean_minmax_single_pair_one_input(FR, rng):
        m = _model(FR)
        out = m.clean(rng.random(12) * 4 + 1, rng.random(12), minmax=[0.0, 6.0], _setattr=True)
        return [np.asarray(o).tolist() for o in out], _state(m)

    @case
    def clean_minmax_wrong_length_raises(FR, rng):
        m = _model(FR)
        return m.clean(rng.random((12, 3)), rng.random(12), minmax=[[0, 1], [0, 1]])

    @case
    def clean_pillow_scalar(FR, rng):
        m = _model(FR)
        out = m.clean(rng.random((12, 3)) * 9, rng.random(12), pillow=0.2, _setattr=True)
        return [np.asarray(o).tolist() for o in out], _state(m)

    @case
    def clean_pillow_pair(FR, rng):
        m = _model(FR)
        out = m.clean(rng.random((12, 3)) * 9, rng.random(12), pillow=[0.1, 0.3], _setattr=True)
        return [np.asarray(o).tolist() for o in out], _state(m)

    @case
    def clean_pillow_absolute_per_input(FR, rng):
        m = _model(FR)
        out = m.clean(rng.random((12, 2)) * 9, rng.random(12), pillow=[[0.5, 1.0], [2.0, 0.0]], pillow_type='absolute',
                      _setattr=True)
        return [np.asarray(o).tolist() for o in out], _state(m)

    @case
    def clean_pillow_type_list(FR, rng):
        m = _model(FR)
        out = m.clean(rng.random((12, 2)) * 9, rng.random(12), pillow=[[0.1, 0.1], [1.0, 2.0]],
                      pillow_type=['percent', 'absolute'], _setattr=True)
        return [np.asarray(o).tolist() for o in out], _state(m)

    @case
    def clean_pillow_and_minmax(FR, rng):
        m = _model(FR)
        out = m.clean(rng.random((12, 2)) * 9, rng.random(12), pillow=0.1, minmax=[[0, 10], [0, 10]], _setattr=True)
        return [np.asarray(o).tolist() for o in out], _state(m)

    @case
    def clean_bad_pillow_type_raises(FR, rng):
        m = _model(FR)
        return m.clean(rng.random((12, 2)), rng.random(12), pillow=0.1, pillow_type='relative')

    @case
    def clean_second_call_keeps_first_attributes(FR, rng):
        m = _model(FR)
        m.clean(rng.random((12, 2)) * 3, rng.random(12))
        out = m.clean(rng.random((7, 2)) * 3, rng.random(7))
        return [np.asarray(o).tolist() for o in out], _state(m)

    # ---- train split -------------------------------------------------------------------------------------------------
    @case
    def clean_train_fraction(FR, rng):
        np.random.seed(5)
        m = _model(FR)
        m.clean(rng.random((50, 2)), rng.random(50), train=0.3, _setattr=True)
        ti, td = m.trainset()
        return np.asarray(ti).tolist(), np.asarray(td).tolist(), _state(m)

    @case
    def generate_trainlog_small(FR, rng):
        np.random.seed(6)
        m = _model(FR)
        m.clean(rng.random((9, 2)), rng.random(9), _setattr=True)
        return [np.asarray(m.generate_trainlog(t)).tolist() for t in (0.01, 0.5, 0.99)] + [m.generate_trainlog(1)]

    # ---- evaluate_basis -----------------------------------------------------------------------------------------------
    @case
    def evaluate_basis_bernoulli_all_derivatives(FR, rng):
        m = _model(FR)
        xs = [0.0, 1.0, 0.3, 0.999]
        return [[float(m.evaluate_basis(m.phis[n], x, d=d)) for x in xs for d in (0, 1, 2)] for n in (0, 1, 5, 19)]

    @case
    def evaluate_basis_cubic_form(FR, rng):
        m = _model(FR)
        c = [0.5, -1.0, 2.0, 3.0]
        return [float(m.evaluate_basis(c, x, kernel='Cubic Splines', d=d)) for x in (0.0, 0.25, 1.0) for d in (0, 1, 2)]

    @case
    def evaluate_basis_kernel_by_index(FR, rng):
        m = _model(FR)
        return [float(m.evaluate_basis([0.5, -1.0, 2.0, 3.0], 0.4, kernel=0)), float(m.evaluate_basis([1.0, 2.0], 0.4, kernel=1))]

    @case
    def evaluate_basis_bad_kernel_raises(FR, rng):
        m = _model(FR)
        return m.evaluate_basis([1.0, 2.0], 0.4, kernel='Splines')

    @case
    def evaluate_basis_bad_derivative_raises(FR, rng):
        m = _model(FR)
        return m.evaluate_basis([1.0, 2.0], 0.4, d=3)

    # ---- constructor / kwargs / clear ---------------------------------------------------------------------------------
    @case
    def constructor_unknown_keyword_raises(FR, rng):
        return FR.FoKL(kernel=1, nonsense=3)

    @case
    def constructor_bad_kernel_raises(FR, rng):
        return FR.FoKL(kernel='Wavelets')

    @case
    def constructor_string_booleans(FR, rng):
        m = FR.FoKL(kernel=1, way3='yes', aic='on', gimmie='no', UserWarnings='off', ConsoleOutput='false')
        return [m.way3, m.aic, m.gimmie, m.UserWarnings, m.ConsoleOutput]

    @case
    def clear_default_and_keep(FR, rng):
        m = _model(FR)
        m.clean(rng.random((12, 2)), rng.random(12), _setattr=True)
        m.betas, m.mtx, m.evs, m.extra = np.ones((3, 3)), np.ones((2, 2)), np.ones(2), 5
        m.clear(keep=['mtx'])
        left = sorted(k for k in vars(m) if not k.startswith('_'))
        m2 = _model(FR)
        m2.betas, m2.extra = 1, 2
        m2.clear(clear=['extra'])
        m3 = _model(FR)
        m3.betas = 1
        m3.clear(all=True)
        return left, sorted(k for k in vars(m2) if not k.startswith('_')), sorted(k for k in vars(m3) if not k.startswith('_'))

    @case
    def inputs_to_phind_cubic(FR, rng):
        m = _model(FR)
        x = np.array([[0.0, 1.0], [1 / 499, 0.5], [0.9999, 1e-12]])
        phis = tuple([[np.zeros(499)] * 4])
        X, phind, xsm = m._inputs_to_phind(x, phis=phis, kernel='Cubic Splines')
        return np.asarray(X).tolist(), np.asarray(phind).tolist(), np.asarray(xsm).tolist()

    @case
    def inputs_to_phind_out_of_range_raises(FR, rng):
        m = _model(FR)
        phis = tuple([[np.zeros(499)] * 4])
        return m._inputs_to_phind(np.array([[1.5, 0.2]]), phis=phis, kernel='Cubic Splines')

    @case
    def evaluate_without_minmax_raises(FR, rng):
        m = _model(FR)
        m.betas, m.mtx = np.ones((50, 2)), np.array([[1.0]])
        return m.evaluate(np.array([[0.5]]))

    @case
    def coverage3_more_draws_than_betas_raises(FR, rng):
        m = _model(FR)
        m.clean(rng.random((12, 1)), rng.random(12), _setattr=True)
        m.betas, m.mtx = np.ones((50, 2)), np.array([[1.0]])
        return m.coverage3(draws=51)

    @case
    def fit_relats_in_exclusion_raises(FR, rng):
        m = _model(FR, relats_in=[[1, 0], [0, 1]], draws=5, burnin=5)
        return m.fit(rng.random((20, 2)), rng.random(20), clean=True)

    @case
    def fit_without_data_raises(FR, rng):
        m = _model(FR)
        return m.fit()

    @case
    def save_and_load_roundtrip(FR, rng):
        import os
        import tempfile
        m = _model(FR, draws=7)
        m.betas = np.arange(6.0).reshape(3, 2)
        d = tempfile.mkdtemp()
        path = m.save(os.path.join(d, 'model'))
        m2 = FR.load(path)
        m3 = FR.load('model', directory=d)
        return os.path.basename(path), m2.draws, m2.betas.tolist(), m3.kernel, type(m2).__module__ + '.' + type(m2).__name__

    @case
    def clear_default(FR, rng):
        m = _model(FR)
        m.clean(rng.random((12, 2)), rng.random(12), _setattr=True)
        m.betas, m.mtx, m.evs, m.extra = np.ones((3, 3)), np.ones((2, 2)), np.ones(2), 5
        m.clear()
        return sorted(k for k in vars(m) if not k.startswith('_'))

    @case
    def clear_keep_list(FR, rng):
        m = _model(FR)
        m.betas, m.mtx, m.extra = np.ones((3, 3)), np.ones((2, 2)), 5
        m.clear(keep=['mtx'])
        return sorted(k for k in vars(m) if not k.startswith('_')), sorted(m.keep)

    @case
    def clear_hyper(FR, rng):
        m = _model(FR)
        m.betas = 1
        m.clear(clear=['kernel', 'phis'])
        return sorted(k for k in vars(m) if not k.startswith('_'))

    @case
    def clear_all(FR, rng):
        m = _model(FR)
        m.betas = 1
        m.clear(all='on')
        return sorted(vars(m))

    @case
    def clean_pillow_flat_list(FR, rng):
        m = _model(FR)
        out = m.clean(rng.random((12, 2)) * 9, rng.random(12), pillow=[0.1, 0.2, 0.3, 0.4], _setattr=True)
        return [np.asarray(o).tolist() for o in out], _state(m)

    @case
    def clean_pillow_integer(FR, rng):
        m = _model(FR)
        out = m.clean(rng.random((12, 2)) * 9, rng.random(12), pillow=1, pillow_type='absolute', _setattr=True)
        return [np.asarray(o).tolist() for o in out], _state(m)

    @case
    def clean_minmax_flat_list(FR, rng):
        m = _model(FR)
        out = m.clean(rng.random((12, 2)) * 9, rng.random(12), minmax=[0.0, 10.0, -1.0, 12.0], _setattr=True)
        return [np.asarray(o).tolist() for o in out], _state(m)

    @case
    def clean_minmax_of_another_model(FR, rng):
        m = _model(FR)
        m.clean(rng.random((12, 2)) * 9, rng.random(12), _setattr=True)
        other = _model(FR)
        out = other.clean(rng.random((5, 2)) * 9, minmax=m.minmax)
        return np.asarray(out).tolist(), _state(other)

    @case
    def warnings_when_enabled(FR, rng):
        import pandas as pd
        m = _model(FR, UserWarnings=True)
        df = pd.DataFrame({'a': rng.integers(0, 9, 20), 'b': rng.integers(0, 9, 20)})
        m.clean(df.T, pd.Series(rng.integers(0, 5, 20)), bit=7, _setattr=True)
        m.evaluate_basis(m.phis[0], 0.5, kernel=1)
        return _state(m)

    @case
    def fit_never_cleaned_vector_data_raises(FR, rng):
        # FR:1277-1297: clean=False with both arguments given skips `clean`; the raw 1-D data then fails at FR:1378
        m = _model(FR, UserWarnings=True, relats_in=[[1, 0], [0, 1]], draws=5, burnin=5)
        return m.fit(rng.random((20, 2)), rng.random(20), train=0.5)

    @case
    def fit_never_cleaned_column_data_goes_on(FR, rng):
        # ... while [n x 1] data passes that point (and stops at the relats_in exclusion, FR:1569, before any training)
        m = _model(FR, UserWarnings=True, relats_in=[[1, 0], [0, 1]], draws=5, burnin=5)
        return m.fit(rng.random((20, 2)), rng.random((20, 1)))

    @case
    def trainset_before_clean_raises(FR, rng):
        return _model(FR).trainset()

    @case
    def str_to_bool_table(FR, rng):
        vals = ['yes', 'y', 'on', 'all', 'true', 'both', 'no', 'n', 'off', 'none', 'n/a', 'false', 1, 0, 2.5, None, [], [0]]
        return [repr(FR._str_to_bool(v)) for v in vals]

    @case
    def process_kwargs(FR, rng):
        a = FR._process_kwargs({'a': 1, 'b': 2}, {'b': 3})
        b = FR._process_kwargs(['a', 'b'], {'b': 3})
        return a, b

    @case
    def process_kwargs_bad_default_raises(FR, rng):
        return FR._process_kwargs(3, {'b': 3})

    # ---- evaluate / coverage3 on a hand-made model (FR:851-1200); on this package the device product is replaced ----
    def _fitted(FR, rng, kernel=1, m=2, rows=60):
        if hasattr(FR, 'eng_predict'):                 # this package: numpy stand-in for K1 + fokl_predict_draws
            import fokl_oracle as fo

            def predict(eng, phis, kern, normputs, terms, betas_sel):
                k = fo.CUBIC if kern == 'Cubic Splines' else fo.BERNOULLI
                x = np.asarray(normputs, dtype=np.float64)
                X = np.hstack([np.ones((x.shape[0], 1)), fo.basis_columns(x, np.asarray(terms, dtype=np.int64), phis, k)])
                return X @ np.asarray(betas_sel, dtype=np.float64).T
            FR.eng_predict = predict
            FR._engine = lambda device=None: None
        model = _model(FR, draws=50)
        x = rng.random((25, m))
        model.clean(x * 4 - 1, rng.standard_normal(25), _setattr=True)
        model.mtx = np.array([[1, 0], [0, 2], [1, 1], [3, 0]], dtype=np.float64)[:, :m]
        model.betas = rng.standard_normal((rows, 5))
        return model

    def _pred(out):
        if isinstance(out, tuple):
            return [np.asarray(o).tolist() if not isinstance(o, list) else o for o in out]
        return np.asarray(out).tolist()

    @case
    def evaluate_defaults(FR, rng):
        np.random.seed(3)
        m = _fitted(FR, rng)
        return _pred(m.evaluate()), np.asarray(m.setnos).tolist()

    @case
    def evaluate_bounds_and_draws(FR, rng):
        np.random.seed(4)
        m = _fitted(FR, rng)
        return _pred(m.evaluate(m.inputs, draws=45, ReturnBounds=1)), _pred(m.evaluate(m.inputs, draws=45, ReturnBounds='yes'))

    @case
    def evaluate_few_draws_with_bounds_warns(FR, rng):
        np.random.seed(5)
        m = _fitted(FR, rng)
        m.UserWarnings = True
        return _pred(m.evaluate(m.inputs, draws=45, ReturnBounds=0)), _pred(m.evaluate(rng.random((4, 2)), draws=45))

    @case
    def evaluate_one_draw(FR, rng):
        np.random.seed(6)
        m = _fitted(FR, rng)
        return _pred(m.evaluate(m.inputs, draws=1))

    @case
    def evaluate_raw_inputs_clean(FR, rng):
        np.random.seed(7)
        m = _fitted(FR, rng)
        return _pred(m.evaluate(rng.random((6, 2)) * 4 - 1, clean=True)), _pred(m.evaluate(rng.random(2) * 4 - 1, clean=True,
                                                                                           SingleInstance=True))

    @case
    def evaluate_clean_on_default_inputs_warns(FR, rng):
        np.random.seed(8)
        m = _fitted(FR, rng)
        return _pred(m.evaluate(clean=True))

    @case
    def evaluate_user_betas_and_mtx(FR, rng):
        np.random.seed(9)
        m = _fitted(FR, rng)
        b = rng.standard_normal((70, 3))
        return _pred(m.evaluate(rng.random((5, 2)), betas=b, mtx=[[1, 0], [0, 1]], draws=50))

    @case
    def evaluate_mtx_single_row(FR, rng):
        np.random.seed(10)
        m = _fitted(FR, rng)
        return _pred(m.evaluate(rng.random((5, 2)), betas=rng.standard_normal((60, 2)), mtx=[2, 1]))

    @case
    def evaluate_one_input_mtx_int(FR, rng):
        np.random.seed(11)
        m = _fitted(FR, rng, m=1)
        return _pred(m.evaluate(rng.random((5, 1)), betas=rng.standard_normal((60, 2)), mtx=3))

    @case
    def evaluate_out_of_range_inputs(FR, rng):
        np.random.seed(12)
        m = _fitted(FR, rng)
        return _pred(m.evaluate(np.array([[1.5, 0.2], [-0.1, 0.5]])))

    @case
    def evaluate_too_many_draws_raises(FR, rng):
        m = _fitted(FR, rng)
        return m.evaluate(m.inputs, betas=m.betas[:20], draws=30)

    @case
    def coverage3_defaults(FR, rng):
        np.random.seed(13)
        m = _fitted(FR, rng)
        return _pred(m.coverage3())

    @case
    def coverage3_no_bounds(FR, rng):
        np.random.seed(14)
        m = _fitted(FR, rng)
        return _pred(m.coverage3(ReturnBounds=False, draws=30))

    @case
    def coverage3_inputs_without_data_warns(FR, rng):
        np.random.seed(15)
        m = _fitted(FR, rng)
        m.UserWarnings = True
        return _pred(m.coverage3(inputs=rng.random((7, 2))))

    @case
    def coverage3_user_inputs_and_data(FR, rng):
        np.random.seed(16)
        m = _fitted(FR, rng)
        return _pred(m.coverage3(inputs=rng.random((7, 2)), data=rng.standard_normal(7), draws=40))

    @case
    def coverage3_unknown_keyword_raises(FR, rng):
        m = _fitted(FR, rng)
        return m.coverage3(colour='red')

    # ---- bss_derivatives (FR:594-805) on a hand-made model; on this package through the host "device" -----------------
    def _deriv_model(FR, rng, kernel=1, m=3):
        if hasattr(FR, 'eng_derivative_draws'):
            from fake_device import FakeDeviceEngine
            eng = FakeDeviceEngine()
            FR._engine = lambda device=None: eng
        model = _model(FR, draws=5) if kernel == 1 else FR.FoKL(phis=_cubic(), UserWarnings=False, ConsoleOutput=False, draws=5)
        x = rng.random((11, m))
        x[0, 0], x[1, m - 1] = 0.0, 1.0
        mtx = np.array([[1, 0, 0], [0, 2, 0], [0, 0, 3], [1, 1, 0], [2, 0, 1], [1, 2, 1]], dtype=np.float64)[:, :m]
        mtx = mtx[np.any(mtx != 0, axis=1)]
        betas = rng.standard_normal((7, mtx.shape[0] + 1))
        minmax = [[-1.0, 3.0], [0.0, 1.0], [10.0, 10.7]][:m]
        model.mtx, model.minmax, model.betas, model.inputs = mtx, minmax, betas, x
        return model, dict(inputs=x, betas=betas, mtx=mtx, minmax=minmax, draws=5)

    def _cubic():
        import os
        import spline_table
        here = os.path.dirname(os.path.abspath(__file__))
        return spline_table.to_phis(np.load(os.path.join(here, '..', 'golden', 'phis_cubic_48.npy')))

    def _arr(out):
        if out is None:
            return None
        if isinstance(out, tuple):
            return [np.asarray(o).tolist() for o in out]
        return np.asarray(out).tolist()

    @case
    def derivs_defaults_from_attributes(FR, rng):
        m, kw = _deriv_model(FR, rng)
        return _arr(m.bss_derivatives())

    @case
    def derivs_d1_index_d2_true(FR, rng):
        m, kw = _deriv_model(FR, rng)
        return _arr(m.bss_derivatives(d1=1, d2=True, **kw))

    @case
    def derivs_d1_strings(FR, rng):
        m, kw = _deriv_model(FR, rng)
        return _arr(m.bss_derivatives(d1='on', d2='off', **kw)), _arr(m.bss_derivatives(d1='off', d2='on', **kw))

    @case
    def derivs_list_of_one(FR, rng):
        m, kw = _deriv_model(FR, rng)
        return _arr(m.bss_derivatives(d1=[2], d2=[0], **kw))

    @case
    def derivs_bad_list_length_raises(FR, rng):
        m, kw = _deriv_model(FR, rng)
        return m.bss_derivatives(d1=[1, 0], **kw)

    @case
    def derivs_bad_type_raises(FR, rng):
        m, kw = _deriv_model(FR, rng)
        return m.bss_derivatives(d1=1.5, **kw)

    @case
    def derivs_none_requested_warns(FR, rng):
        m, kw = _deriv_model(FR, rng)
        m.UserWarnings = True
        return _arr(m.bss_derivatives(d1=False, d2=False, **kw))

    @case
    def derivs_betas_transposed_and_list(FR, rng):
        m, kw = _deriv_model(FR, rng)
        kw['betas'] = kw['betas'].T.tolist()
        return _arr(m.bss_derivatives(**kw))

    @case
    def derivs_betas_wrong_shape_raises(FR, rng):
        m, kw = _deriv_model(FR, rng)
        kw['betas'] = kw['betas'][:, :3]
        return m.bss_derivatives(**kw)

    @case
    def derivs_individual_draws_full_array(FR, rng):
        m, kw = _deriv_model(FR, rng)
        return _arr(m.bss_derivatives(d1=True, d2=[1, 0, 1], IndividualDraws='yes', ReturnFullArray=1, **kw))

    @case
    def derivs_one_draw(FR, rng):
        m, kw = _deriv_model(FR, rng)
        kw['draws'] = 1
        return _arr(m.bss_derivatives(IndividualDraws=True, **kw))

    @case
    def derivs_one_input_vector(FR, rng):
        m, kw = _deriv_model(FR, rng, m=1)
        kw['inputs'] = kw['inputs'][:, 0]
        kw['minmax'] = [-1.0, 3.0]
        return _arr(m.bss_derivatives(d2=0, **kw))

    @case
    def derivs_mtx_int(FR, rng):
        m, kw = _deriv_model(FR, rng, m=1)
        kw['mtx'], kw['betas'] = 2, kw['betas'][:, :2]
        return _arr(m.bss_derivatives(**kw))

    @case
    def derivs_return_basis(FR, rng):
        m, kw = _deriv_model(FR, rng)
        return _arr(m.bss_derivatives(d1=[0, 1, 0], ReturnBasis=True, **kw))

    @case
    def derivs_cubic_kernel(FR, rng):
        m, kw = _deriv_model(FR, rng, kernel=0)
        return _arr(m.bss_derivatives(d1=True, d2=True, **kw))

    @case
    def derivs_cubic_return_basis(FR, rng):
        m, kw = _deriv_model(FR, rng, kernel=0)
        return _arr(m.bss_derivatives(d1=2, ReturnBasis=True, **kw))

    @case
    def derivs_kernel_by_index(FR, rng):
        m, kw = _deriv_model(FR, rng)
        return _arr(m.bss_derivatives(kernel=1, **kw))

    @case
    def derivs_unsupported_kernel_raises(FR, rng):
        m, kw = _deriv_model(FR, rng)
        return m.bss_derivatives(kernel='Wavelets', **kw)

    @case
    def derivs_inputs_outside_unit_interval_warn(FR, rng):
        m, kw = _deriv_model(FR, rng)
        m.UserWarnings = True
        kw['inputs'] = kw['inputs'] * 1.0
        kw['inputs'][3, 1] = 1.0 + 1e-9
        return _arr(m.bss_derivatives(**kw))

    @case
    def derivs_unknown_keyword_raises(FR, rng):
        m, kw = _deriv_model(FR, rng)
        return m.bss_derivatives(order=2, **kw)

    # ---- getKernels / module helpers --------------------------------------------------------------------------------
    @case
    def getkernels_bernoulli_default_and_named(FR, rng):
        from FoKL import getKernels
        a, b = getKernels.bernoulli(), getKernels.bernoulli('orthogonal_Bn_scaled.txt')
        return [list(map(float, r)) for r in a], [len(r) for r in b], a == b

    @case
    def getkernels_sp500_unknown_keyword_raises(FR, rng):
        from FoKL import getKernels
        return getKernels.sp500(Smoothing=1)

    @case
    def getkernels_bss_anova_writes_the_eigenvalue_file(FR, rng):
        import os
        import tempfile
        from FoKL import getKernels
        cwd = os.getcwd()
        os.chdir(tempfile.mkdtemp())
        try:
            out = getKernels.bss_anova(n=60)
            vals = np.loadtxt("BSS-ANOVA__sqrt-eigvals__K-500x500.txt", delimiter=",")
        finally:
            os.chdir(cwd)
        return out, len(vals), [round(float(v), 9) for v in vals[:12]]

    @case
    def set_attributes_helper(FR, rng):
        m = _model(FR)
        FR._set_attributes(m, {'alpha': 1, 'beta': [2]})
        m.UserWarnings = True
        FR._set_attributes(m, ['not', 'a', 'dict'])
        return m.alpha, m.beta

    return c


def run(FR):
    out = {}
    for name, fn in cases().items():
        rng = np.random.default_rng(abs(hash(name)) % 1000 if False else sum(map(ord, name)))
        with warnings.catch_warnings(record=True) as w:
            warnings.simplefilter('always')
            try:
                res = ('ok', fn(FR, rng))
            except Exception as exc:  # noqa: BLE001 -- the exception TYPE is the outcome under comparison
                res = ('raised', type(exc).__name__)
        out[name] = dict(result=res, warnings=sorted({x.category.__name__ + ': ' + str(x.message)[:60] for x in w
                                                      if issubclass(x.category, UserWarning)}))
    return out


if __name__ == '__main__':
    from FoKL import FoKLRoutines as FR
    res = run(FR)
    with open(sys.argv[1], 'wb') as f:
        pickle.dump(dict(file=FR.__file__, outcomes=res), f)
