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


def test_bernoulli_way3(engine, phis_bern):
    rng = np.random.default_rng(3)
    x = rng.random((1234, 4))
    terms = fo.distinct_perms([2, 1, 1, 0]).astype(int)
    got = build_on_device(engine, phis_bern, fo.BERNOULLI, x, terms)
    ref = fo.basis_columns(x, terms, phis_bern, fo.BERNOULLI)
    assert np.allclose(got, ref, rtol=1e-12, atol=1e-15)
