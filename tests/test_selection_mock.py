"""Host selection loop (FoKL/_selection.py) on the CPU stand-in engine (tests/mock_engine.py): the fast path (device
kill loop, batched and speculative chains) must equal the literal path (one spectral evaluation per proposal, the
reference's order FR:1669-1690) bit for bit."""
import numpy as np
import pytest

import fokl_oracle as fo
from FoKL import _selection
from mock_engine import MockEngine


def select_on_mock(x, y, phis, dist=None, pipeline=None, draws=80):
    """forward_select on a (possibly row-sharded) stand-in engine with fixed hyper-parameters; the caller seeds numpy."""
    eng = MockEngine(x, y, phis, fo.CUBIC, dist=dist)
    hy = dict(a=4.0, b=0.5, atau=4.0, btau=10.0, tolerance=3, total_draws=draws, gimmie=False, way3=True,
              threshav=0.05, threshstda=0.5, threshstdb=2.0, aic=False)
    out = _selection.forward_select(eng, hy, x.shape[1], len(phis), console=False, rng='philox', pipeline=pipeline)
    return out, eng


def synthetic(n, m, seed, noise=0.1):
    return _data(n, m, seed, noise)


def _run(x, y, phis, kernel, eager, seed, way3=True, draws=80, tolerance=3, aic=False, pipeline=None, **hy_kw):
    eng = MockEngine(x, y, phis, kernel)
    a, atau = 4.0, 4.0
    b, btau = fo.default_b_btau(y, a, atau)
    hy = dict(a=a, b=b, atau=atau, btau=btau, tolerance=tolerance, total_draws=draws, gimmie=False, way3=way3,
              threshav=0.05, threshstda=0.5, threshstdb=2.0, aic=aic)
    hy.update(hy_kw)
    np.random.seed(seed)
    out = _selection.forward_select(eng, hy, x.shape[1], len(phis), console=False, rng='philox', eager=eager,
                                    pipeline=pipeline)
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
    (300, 3, 8, dict(gimmie=True)),
    # every proposal threshold-dependent and the threshold in the middle of the candidates: the true chains change
    # the proposal mask, i.e. the speculation fails and the substage is rolled back (checked below)
    (210, 3, 22, dict(threshav=1.0, threshstda=0.01, threshstdb=1e9, noise=0.5, rollback=True)),
    (230, 4, 38, dict(threshav=1.0, threshstda=0.01, threshstdb=1e9, noise=0.5, rollback=True)),
    (250, 4, 47, dict(threshav=2.0, threshstda=0.01, threshstdb=1e9, noise=0.5, rollback=True))])
def test_fast_path_equals_literal_path(phis_cubic, n, m, seed, kw):
    kw = dict(kw)
    x, y = _data(n, m, seed, noise=kw.pop('noise', 0.1))
    rollback = kw.pop('rollback', False)
    fast, eng_f = _run(x, y, phis_cubic, fo.CUBIC, False, seed, **kw)
    slow, eng_s = _run(x, y, phis_cubic, fo.CUBIC, True, seed, **kw)
    assert np.array_equal(fast['mtx'], slow['mtx'])
    assert np.array_equal(fast['evs'], slow['evs'])
    assert np.array_equal(fast['betas'], slow['betas'])
    assert fast['n_gibbs'] == slow['n_gibbs']
    # ... and the sequential form of the fast path (no speculation across substages) as well
    seq, eng_q = _run(x, y, phis_cubic, fo.CUBIC, False, seed, pipeline=False, **kw)
    assert np.array_equal(fast['mtx'], seq['mtx']) and np.array_equal(fast['evs'], seq['evs'])
    assert np.array_equal(fast['betas'], seq['betas']) and fast['n_gibbs'] == seq['n_gibbs']
    assert ('launch', 'side') not in eng_q.calls
    if 'tolerance' not in kw and m > 1:
        assert ('launch', 'side') in eng_f.calls      # the pipelined form did speculate
    rolled = sum(1 for i, c in enumerate(eng_f.calls[:-1]) if c[0] == 'truncate' and eng_f.calls[i + 1][0] == 'append')
    assert (rolled > 0) == rollback
    # the engine ends on the last substage's model in every form
    assert eng_f.P == eng_s.P == eng_q.P


def test_speculation_discarded_when_the_fit_finishes(phis_cubic):
    """pipeline='always' speculates even when the stopping rule is predicted to fire: the last substage's speculative
    successor must be dropped again and the result must not change."""
    x, y = _data(300, 3, 1)
    ref, eng_r = _run(x, y, phis_cubic, fo.CUBIC, False, 1, pipeline=False)
    alw, eng_a = _run(x, y, phis_cubic, fo.CUBIC, False, 1, pipeline='always')
    assert np.array_equal(ref['mtx'], alw['mtx']) and np.array_equal(ref['evs'], alw['evs'])
    assert np.array_equal(ref['betas'], alw['betas']) and ref['n_gibbs'] == alw['n_gibbs']
    assert eng_a.calls[-1][0] == 'truncate' and eng_a.P == eng_r.P


def test_interpolating_model_takes_the_literal_path(phis_cubic):
    """A (nearly) interpolating model: the Gram-only BIC loses its digits to cancellation and is recomputed from the
    residual pass over X (Engine.refine_mask); the device kill loop scores with the same Gram-only form, so such a
    substage must run the literal loop.  Same result in every form."""
    rng = np.random.default_rng(12)
    x = rng.random((250, 3))
    y = 1.0 + 2.0 * fo.basis_columns(x, np.array([[1, 0, 0]]), phis_cubic, fo.CUBIC)[:, 0]
    fast, eng_f = _run(x, y, phis_cubic, fo.CUBIC, False, 12)
    seq, eng_q = _run(x, y, phis_cubic, fo.CUBIC, False, 12, pipeline=False)
    slow, eng_s = _run(x, y, phis_cubic, fo.CUBIC, True, 12)
    for other in (seq, slow):
        assert np.array_equal(fast['mtx'], other['mtx']) and np.array_equal(fast['evs'], other['evs'])
        assert np.array_equal(fast['betas'], other['betas']) and fast['n_gibbs'] == other['n_gibbs']
    assert any(c[0] == 'refine' for c in eng_f.calls) and not any(c[0] == 'kill_loop' for c in eng_f.calls)


def _digest():
    import hashlib
    st = np.random.get_state()
    return hashlib.sha256(st[1].tobytes() + bytes(str((st[2], st[3], repr(st[4]))), 'ascii')).hexdigest()


@pytest.mark.parametrize('n,m,seed,way3,aic', [(120, 2, 3, False, False), (150, 3, 4, True, False), (100, 2, 5, False, True)])
def test_parity_mode_equals_the_oracle_fit(phis_cubic, n, m, seed, way3, aic):
    """The whole host loop in parity mode (numpy variates injected in the reference's order, eigenvector signs aligned
    with LAPACK) on the stand-in engine against the oracle's `fit` -- which is pinned to the unmodified reference: same
    term matrix, same number of `gibbs` calls, same RNG end state, BIC trace and draws to rtol 1e-9 / 1e-7."""
    x, y = _data(n, m, seed)
    D = 60
    a, atau = 4.0, 4.0
    b, btau = fo.default_b_btau(y, a, atau)
    np.random.seed(seed)
    want = fo.fit(x, y, phis_cubic, kernel=fo.CUBIC, a=a, b=b, atau=atau, btau=btau, tolerance=3, burnin=D // 2,
                  draws=D // 2, way3=way3, aic=aic)
    want_digest = _digest()
    eng = MockEngine(x, y, phis_cubic, fo.CUBIC)
    hy = dict(a=a, b=b, atau=atau, btau=btau, tolerance=3, total_draws=D, gimmie=False, way3=way3, threshav=0.05,
              threshstda=0.5, threshstdb=2.0, aic=aic)
    np.random.seed(seed)
    got = _selection.forward_select(eng, hy, m, len(phis_cubic), console=False, rng='numpy')
    assert np.array_equal(got['mtx'], want.mtx)
    assert got['n_gibbs'] == want.n_gibbs and _digest() == want_digest
    assert np.allclose(got['evs'], want.evs, rtol=1e-9, atol=0)
    full = np.asarray(want.betas_full)
    assert got['betas'].shape == full.shape
    if full.shape[1] < n // 3:       # the draws of a nearly saturated model (p ~ N) hinge on degenerate eigen-directions
        assert np.allclose(got['betas'], full, rtol=1e-7, atol=1e-7 * np.max(np.abs(full)))


def test_the_host_loops_call_only_what_the_real_engine_defines():
    """Guard against drift between tests/mock_engine.py and FoKL._engine.Engine: every `engine.<name>` the host loops
    (FoKL/_selection.py, FoKL/_update.py, FoKL/FoKLRoutines.py) touch must exist on the real Engine (or be probed with hasattr / getattr there),
    and the keyword names the loops pass must be parameters of the real method -- so a loop that runs on the stand-in
    cannot be calling something the device engine does not have."""
    import ast
    import inspect
    import os
    from FoKL import FoKLRoutines, _engine, _selection, _update
    real = _engine.Engine
    init_src = inspect.getsource(real)
    real_attrs = {n for n, _ in inspect.getmembers(real)} | set(__import__('re').findall(r'self\.([A-Za-z_]\w*)\s*=', init_src))
    for mod in (_selection, _update, FoKLRoutines):
        tree = ast.parse(inspect.getsource(mod))
        src = inspect.getsource(mod)
        for node in ast.walk(tree):
            if isinstance(node, ast.Attribute) and isinstance(node.value, ast.Name) and node.value.id in ('engine', 'e', 'eng'):
                name = node.attr
                if name in real_attrs:
                    continue
                probed = ("hasattr(engine, '%s')" % name in src) or ("getattr(engine, '%s'" % name in src)
                assert probed, '%s uses engine.%s, which FoKL._engine.Engine does not define' % (os.path.basename(mod.__file__), name)
            if isinstance(node, ast.Call) and isinstance(node.func, ast.Attribute) and isinstance(node.func.value, ast.Name) \
                    and node.func.value.id in ('engine', 'e', 'eng') and hasattr(real, node.func.attr):
                params = inspect.signature(getattr(real, node.func.attr)).parameters
                if any(p.kind == p.VAR_KEYWORD for p in params.values()):
                    continue
                for kw in node.keywords:
                    if kw.arg is not None:
                        assert kw.arg in params, 'Engine.%s has no keyword %r' % (node.func.attr, kw.arg)
