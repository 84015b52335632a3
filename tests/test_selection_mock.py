"""Host selection loop (FoKL/_selection.py) on the CPU stand-in engine (tests/mock_engine.py): the fast path (device
kill loop, batched and speculative chains) must equal the literal path (one spectral evaluation per proposal, the
reference's order FR:1669-1690) bit for bit."""
import numpy as np
import pytest

import fokl_oracle as fo
from FoKL import _selection
from mock_engine import MockEngine


def _run(x, y, phis, kernel, eager, seed, way3=True, draws=80, tolerance=3, aic=False, **hy_kw):
    eng = MockEngine(x, y, phis, kernel)
    a, atau = 4.0, 4.0
    b, btau = fo.default_b_btau(y, a, atau)
    hy = dict(a=a, b=b, atau=atau, btau=btau, tolerance=tolerance, total_draws=draws, gimmie=False, way3=way3,
              threshav=0.05, threshstda=0.5, threshstdb=2.0, aic=aic)
    hy.update(hy_kw)
    np.random.seed(seed)
    out = _selection.forward_select(eng, hy, x.shape[1], len(phis), console=False, rng='philox', eager=eager)
    return out, eng


def _data(n, m, seed, noise=0.1):
    rng = np.random.default_rng(seed)
    x = rng.random((n, m))
    y = np.sin(2 * np.pi * x[:, 0]) + 0.1 * noise * rng.standard_normal(n)
    if m > 1:
        y = y + x[:, 0] * x[:, 1]
    if m > 2:
        y = y + 0.5 * x[:, 2] ** 2 + noise * rng.standard_normal(n)
    return x, y


@pytest.mark.parametrize('n,m,seed,kw', [
    (300, 3, 1, {}), (500, 2, 2, dict(way3=False)), (200, 4, 3, {}), (400, 3, 4, dict(aic=True)),
    (250, 3, 5, dict(threshav=0.6, threshstda=0.05)), (150, 1, 6, {}), (300, 3, 7, dict(tolerance=1)),
    (300, 3, 8, dict(gimmie=True))])
def test_fast_path_equals_literal_path(phis_cubic, n, m, seed, kw):
    x, y = _data(n, m, seed)
    fast, eng_f = _run(x, y, phis_cubic, fo.CUBIC, False, seed, **kw)
    slow, eng_s = _run(x, y, phis_cubic, fo.CUBIC, True, seed, **kw)
    assert np.array_equal(fast['mtx'], slow['mtx'])
    assert np.array_equal(fast['evs'], slow['evs'])
    assert np.array_equal(fast['betas'], slow['betas'])
    assert fast['n_gibbs'] == slow['n_gibbs']
