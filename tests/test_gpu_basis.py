"""K1 (fokl_basis_build) against the oracle: bit-exact for cubic splines, stated tolerance for Bernoulli."""
import numpy as np
import pytest

import fokl_oracle as fo

pytestmark = pytest.mark.gpu


def build_on_device(engine, phis, kernel, x, terms):
    import torch
    engine.set_phis(phis, kernel)
    ds = engine.upload(x, np.zeros(x.shape[0]))
    terms = np.ascontiguousarray(terms, dtype=np.int16)
    c = terms.shape[0]
    X = torch.full((c, ds.ldx), np.nan, dtype=torch.float64, device=engine.device)
    engine._ck(engine.lib.fokl_basis_build(engine.ctx, engine.kernel_id, ds.x.data_ptr(), ds.n, ds.ldx, ds.m,
                                           terms.ctypes.data, c, X.data_ptr(), ds.ldx))
    engine.synchronize()
    return X[:, :ds.n].t().cpu().numpy()


def random_terms(rng, m, count, max_order, max_way=3):
    seen, out = set(), []
    while len(out) < count:
        t = np.zeros(m, dtype=int)
        ks = rng.choice(m, rng.integers(1, min(max_way, m) + 1), replace=False)
        t[ks] = rng.integers(1, max_order + 1, len(ks))
        if tuple(t) not in seen:
            seen.add(tuple(t))
            out.append(t)
    return np.array(out)


@pytest.mark.parametrize('n', [1, 2, 255, 256, 257, 511, 513, 4099])
def test_cubic_bit_exact_ragged_sizes(engine, phis_cubic, n):
    rng = np.random.default_rng(n)
    x = rng.random((n, 3))
    terms = random_terms(rng, 3, 17, 12)
    got = build_on_device(engine, phis_cubic, fo.CUBIC, x, terms)
    ref = fo.basis_columns(x, terms, phis_cubic, fo.CUBIC)
    assert np.array_equal(got, ref)


def test_cubic_edge_inputs(engine, phis_cubic):
    x = np.array([0.0, 1.0, 0.5, 1 / 499, 2 / 499, 1e-300, 1 - 2 ** -53, 498 / 499, 0.25])[:, None]
    x = np.hstack([x, x[::-1]])
    terms = np.array([[1, 0], [0, 1], [2, 3], [48, 48], [5, 0]])
    got = build_on_device(engine, phis_cubic, fo.CUBIC, x, terms)
    ref = fo.basis_columns(x, terms, phis_cubic, fo.CUBIC)
    assert np.array_equal(got, ref)


def test_cubic_way3_substage_shapes(engine, phis_cubic):
    """The term lists fit() really produces: all distinct permutations of (2,1,1,0,...) and (3,2,1,...) over 8 inputs."""
    rng = np.random.default_rng(8)
    x = rng.random((3000, 8))
    for part in ([2, 1, 1, 0, 0, 0, 0, 0], [3, 2, 1, 0, 0, 0, 0, 0], [1, 0, 0, 0, 0, 0, 0, 0]):
        terms = fo.distinct_perms(part).astype(int)
        got = build_on_device(engine, phis_cubic, fo.CUBIC, x, terms)
        ref = fo.basis_columns(x, terms, phis_cubic, fo.CUBIC)
        assert np.array_equal(got, ref), part


def test_cubic_many_factors_chunked(engine, phis_cubic):
    """More than 96 distinct (input, order) factors forces several launches / unstaged tables."""
    rng = np.random.default_rng(5)
    x = rng.random((700, 6))
    terms = random_terms(rng, 6, 400, 40)
    got = build_on_device(engine, phis_cubic, fo.CUBIC, x, terms)
    ref = fo.basis_columns(x, terms, phis_cubic, fo.CUBIC)
    assert np.array_equal(got, ref)


def test_cubic_out_of_range_raises(engine, phis_cubic):
    x = np.array([[0.2], [1.5], [0.1]])
    with pytest.raises(ValueError):
        build_on_device(engine, phis_cubic, fo.CUBIC, x, np.array([[1]]))
    # the context stays usable
    got = build_on_device(engine, phis_cubic, fo.CUBIC, np.array([[0.2]]), np.array([[1]]))
    assert np.isfinite(got).all()


def test_bernoulli_tolerance_by_order(engine, phis_bern):
    """glibc pow() is not correctly rounded in ~0.1 % of calls and the monomial sum amplifies one ulp by up to
    1e14 at order 20, so exact agreement is impossible; tolerance: 1e-9 of the column scale for orders <= 10."""
    rng = np.random.default_rng(2)
    x = rng.random((5000, 2))
    for order, tol in ((1, 0.0), (2, 1e-14), (4, 1e-12), (6, 1e-10), (10, 1e-9)):
        terms = np.array([[order, 0], [0, order], [order, order]])
        got = build_on_device(engine, phis_bern, fo.BERNOULLI, x, terms)
        ref = fo.basis_columns(x, terms, phis_bern, fo.BERNOULLI)
        scale = np.abs(ref).max(axis=0)
        assert np.all(np.abs(got - ref).max(axis=0) <= tol * scale), order
    # most cells are bit-exact even at moderate order
    terms = np.array([[3, 0], [5, 2]])
    got = build_on_device(engine, phis_bern, fo.BERNOULLI, x, terms)
    ref = fo.basis_columns(x, terms, phis_bern, fo.BERNOULLI)
    assert np.mean(got == ref) > 0.98


def test_bernoulli_orders_11_to_20(engine, phis_bern):
    """The upper half of the shipped table (phis has 20 orders and the selection loop may reach ind = len(phis),
    FR:1747) with the per-order tolerance of tests/test_host_emu.py::bernoulli_order_tolerance -- the reference's own
    rounding noise against the exact polynomial at that order, measured there."""
    from test_host_emu import bernoulli_order_tolerance
    rng = np.random.default_rng(12)
    x = rng.random((5000, 2))
    for order in range(11, 21):
        terms = np.array([[order, 0], [0, order], [order, 1]])
        got = build_on_device(engine, phis_bern, fo.BERNOULLI, x, terms)
        ref = fo.basis_columns(x, terms, phis_bern, fo.BERNOULLI)
        scale = np.abs(ref).max(axis=0)
        assert np.all(np.abs(got - ref).max(axis=0) <= bernoulli_order_tolerance(order) * scale), order


def test_bernoulli_way3(engine, phis_bern):
    rng = np.random.default_rng(3)
    x = rng.random((1234, 4))
    terms = fo.distinct_perms([2, 1, 1, 0]).astype(int)
    got = build_on_device(engine, phis_bern, fo.BERNOULLI, x, terms)
    ref = fo.basis_columns(x, terms, phis_bern, fo.BERNOULLI)
    assert np.allclose(got, ref, rtol=1e-12, atol=1e-15)


# ---- bss_derivatives (FR:594-805; a SURVEY section-8f "next" row) ------------------------------------------------

def _deriv_model(FR, phis, kernel, g):
    model = FR.FoKL(phis=phis, kernel=kernel, UserWarnings=False)
    model.mtx, model.minmax, model.draws, model.betas = g['mtx'], g['minmax'].tolist(), int(g['draws']), g['betas']
    return model, dict(inputs=g['x'], betas=g['betas'], mtx=g['mtx'], minmax=g['minmax'].tolist(), draws=int(g['draws']))


@pytest.mark.gpu
@pytest.mark.parametrize('name', ['cubic', 'bern'])
def test_bss_derivatives_against_reference_golden(name, phis_cubic, phis_bern):
    """FoKL.bss_derivatives through the derivative basis kernel vs the outputs of the unmodified reference
    (tests/golden/bss_derivatives.npz, oracle/gen_golden.py).  Tolerance: rtol 1e-9 of the column scale (the design
    columns are bit-exact for cubic splines; the product with the draws sums in a different order)."""
    import warnings
    from FoKL import FoKLRoutines as FR
    from conftest import load_golden
    g = load_golden('bss_derivatives')
    phis, kernel = (phis_cubic, 'Cubic Splines') if name == 'cubic' else (phis_bern, 'Bernoulli Polynomials')
    model, kw = _deriv_model(FR, phis, kernel, g)

    def close(got, want):
        assert np.shape(got) == np.shape(want)
        scale = np.max(np.abs(want), axis=0, keepdims=True) + 1e-300
        assert np.max(np.abs(got - want) / scale) < 1e-9

    with warnings.catch_warnings():
        warnings.simplefilter('ignore')
        close(model.bss_derivatives(**kw), g[name + '_grad'])
        close(model.bss_derivatives(d1=[1, 0, 1], d2=[0, 1, 1], **kw), g[name + '_d1d2'])
        close(model.bss_derivatives(d1=True, d2=True, IndividualDraws=True, ReturnFullArray=True, **kw),
              g[name + '_full_draws'])
        close(model.bss_derivatives(d1=False, d2=1, **kw), g[name + '_d2_only'])


@pytest.mark.gpu
@pytest.mark.parametrize('n', [1, 255, 4097])
def test_derivative_columns_bit_exact_cubic(engine, phis_cubic, n):
    """fokl_basis_build_deriv vs the oracle's literal loop (FR:757-790): bit-exact for cubic splines, ragged sizes,
    first and second derivatives, 1- to 3-way terms."""
    import torch
    import fokl_oracle as fo
    rng = np.random.default_rng(n)
    m = 3
    x = rng.random((n, m))
    x[0, 0] = 1.0
    terms = np.array([[1, 0, 0], [0, 2, 0], [3, 1, 0], [2, 0, 5], [1, 2, 3], [0, 0, 7]])
    span = [2.5, 1.0, 0.3]
    engine.set_phis(phis_cubic, 'Cubic Splines')
    ds = engine.upload(x, np.zeros(n))
    for wrt in range(m):
        for order in (1, 2):
            idx = np.nonzero(terms[:, wrt])[0]
            want = fo.derivative_columns(x, terms, phis_cubic, fo.CUBIC, wrt, order, span[wrt])[:, idx]
            t16 = np.ascontiguousarray(terms[idx], dtype=np.int16)
            d8 = np.zeros((len(idx), m), dtype=np.uint8)
            d8[:, wrt] = order
            dv = np.ones((m, 3))
            for k in range(m):
                s_l = span[k] / 499
                dv[k] = [1, s_l, s_l ** 2]
            X = torch.empty((len(idx), ds.ldx), dtype=torch.float64, device=engine.device)
            engine._ck(engine.lib.fokl_basis_build_deriv(engine.ctx, engine.kernel_id, ds.x.data_ptr(), n, ds.ldx, m,
                                                         t16.ctypes.data, d8.ctypes.data, dv.ctypes.data, len(idx),
                                                         X.data_ptr(), ds.ldx))
            engine.synchronize()
            got = X[:, :n].cpu().numpy().T
            assert np.array_equal(got, want), (wrt, order)


# ---- evaluate / coverage3 (FR:851-1200) ------------------------------------------------------

@pytest.mark.gpu
@pytest.mark.parametrize('name', ['cubic', 'bern'])
def test_evaluate_and_coverage3_against_reference_golden(name, phis_cubic, phis_bern):
    """FoKL.evaluate / coverage3 on the device (K1 + fokl_predict_draws) vs outputs of the unmodified reference
    (tests/golden/evaluate.npz, oracle/gen_golden.py): same `setnos` from the seeded global RNG (FR:933-937), mean and
    95 % bounds to rtol 1e-9 of the output scale, the coverage3 'rmse' as the reference computes it (FR:1193)."""
    import warnings
    from FoKL import FoKLRoutines as FR
    from conftest import load_golden
    g = load_golden('evaluate')
    phis, kernel = (phis_cubic, 'Cubic Splines') if name == 'cubic' else (phis_bern, 'Bernoulli Polynomials')
    model = FR.FoKL(phis=phis, kernel=kernel, UserWarnings=False)
    model.mtx, model.minmax, model.draws, model.betas = g['mtx'], g['minmax'].tolist(), 100, g['betas']
    model.inputs, model.data = g['x'], g['data']
    scale = np.max(np.abs(g[name + '_mean'])) + 1e-300

    def close(got, want):
        assert np.shape(got) == np.shape(want)
        assert np.max(np.abs(np.asarray(got) - want)) < 1e-9 * scale

    with warnings.catch_warnings():
        warnings.simplefilter('ignore')
        np.random.seed(int(g['seed']))
        mean, bounds = model.evaluate(g['x'], ReturnBounds=1)
        assert np.array_equal(model.setnos, g[name + '_setnos'])
        close(mean, g[name + '_mean'])
        close(bounds, g[name + '_bounds'])
        close(model.evaluate(g['raw'], clean=True), g[name + '_mean_raw'])
        close(model.evaluate(g['x'], draws=20), g[name + '_mean_20'])
        cm, cb, rmse = model.coverage3(inputs=g['x'], data=g['data'], draws=100)
        close(cm, g[name + '_cov_mean'])
        close(cb, g[name + '_cov_bounds'])
        assert abs(float(rmse) - float(g[name + '_cov_rmse'])) < 1e-9 * scale


@pytest.mark.gpu
def test_model_saved_by_the_reference_predicts_the_same():
    """tests/golden/ref_model.fokl was trained, evaluated and pickled by the unmodified reference; loaded here, the
    device `evaluate` must return the reference's own mean and bounds (rtol 1e-9 of the output scale)."""
    import os
    import warnings
    from FoKL import FoKLRoutines as FR
    from conftest import GOLD
    g = np.load(os.path.join(GOLD, 'ref_model_expect.npz'))
    model = FR.load(os.path.join(GOLD, 'ref_model.fokl'))
    with warnings.catch_warnings():
        warnings.simplefilter('ignore')
        mean, bounds = model.evaluate(ReturnBounds=1)
    scale = np.max(np.abs(g['mean']))
    assert np.max(np.abs(mean - g['mean'])) < 1e-9 * scale
    assert np.max(np.abs(bounds - g['bounds'])) < 1e-9 * scale
