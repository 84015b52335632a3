"""K3/K4 (fokl_candidates_eval) against the oracle's gibbs (FR:1492-1558).

Tolerances: BIC rtol 1e-9; eigenvalues 1e-12 of the largest; betahat / betas rtol 1e-9 of the largest entry on
well-conditioned Grams when the same numpy variates are injected and eigenvector signs are aligned with
scipy.linalg.eigh of the same Gram bits."""
import numpy as np
import pytest

import fokl_oracle as fo
from FoKL import _lib

pytestmark = pytest.mark.gpu


def make_problem(phis, n, m, seed, way3=False, ind_max=2):
    rng = np.random.default_rng(seed)
    x = rng.random((n, m))
    y = np.sin(2 * np.pi * x[:, 0]) + 2 * (x[:, 1] - 0.5) ** 2 + 0.05 * rng.standard_normal(n)
    if m > 2:
        y = y + x[:, 1] * x[:, 2]
    parts = [[1] + [0] * (m - 1), [1, 1] + [0] * (m - 2), [2] + [0] * (m - 1)]
    if way3 and m > 2:
        parts.append([1, 1, 1] + [0] * (m - 3))
    if ind_max > 2:
        parts += [[2, 1] + [0] * (m - 2), [3] + [0] * (m - 1)]
    terms = np.vstack([fo.distinct_perms(p).astype(int) for p in parts])
    X = np.hstack([np.ones((n, 1)), fo.basis_columns(x, terms, phis, fo.CUBIC)])
    return x, y[:, None], terms, X


def setup_engine(engine, phis, x, y, terms):
    engine.set_phis(phis, fo.CUBIC)
    ds = engine.upload(x, y)
    engine.begin_fit(ds)
    engine.append_terms(terms)
    return ds


def oracle_with_device_gram(engine, X, y, cols, a, b, atau, btau, D, seed):
    import torch
    idx = torch.as_tensor(np.asarray(cols), device=engine.device)
    G = engine.G.index_select(0, idx).index_select(1, idx).cpu().numpy()
    Xty = engine.Xty.index_select(0, idx).cpu().numpy()
    dtd = y.T.dot(y)
    np.random.seed(seed)
    r = fo.gibbs_from_X(X[:, cols], y, a, b, atau, btau, D, b / (1 + a), btau / (1 + atau), dtd, literal=True,
                        gram=(G, Xty))
    np.random.seed(seed)
    p = len(cols)
    z, g1, g2 = fo.draw_variates(p, D, a + 1 + len(y) / 2 + p / 2, atau + (p - 1) / 2)
    return r, np.hstack([z, g1[:, None], g2[:, None]])


@pytest.mark.parametrize('n,m,seed', [(441, 2, 1), (5000, 3, 2), (20000, 4, 3)])
def test_bic_betahat_eig_match_oracle(engine, phis_cubic, n, m, seed):
    x, y, terms, X = make_problem(phis_cubic, n, m, seed)
    setup_engine(engine, phis_cubic, x, y, terms)
    a = atau = 4
    b, btau = fo.default_b_btau(y, a, atau)
    hyp = engine.make_hypers(a, b, atau, btau, b / (1 + a), btau / (1 + atau), 50)
    P = engine.P
    sets = [list(range(P)), list(range(P - 2)), [0] + list(range(2, P)), [0, 1], [0]]
    res = engine.evaluate(sets, hyp, rng_mode=_lib.RNG_NONE, want_eig=True, refine_tol=None)
    dtd = y.T.dot(y)
    for c, cols in enumerate(sets):
        r = fo.gibbs_from_X(X[:, cols], y, a, b, atau, btau, 1, 1.0, 1.0, dtd, literal=False,
                            variates=(np.zeros((1, len(cols))), np.ones(1), np.ones(1)))
        assert abs(res.ev[c] - r['ev']) <= 1e-9 * abs(r['ev']), (c, res.ev[c], r['ev'])
        p = len(cols)
        lam = res.lamb[res.vec_off[c]:res.vec_off[c] + p].cpu().numpy()
        assert np.all(np.abs(lam - r['Lamb']) <= 1e-12 * r['Lamb'][-1])
        bh = res.betahat[res.vec_off[c]:res.vec_off[c] + p].cpu().numpy()
        cond = r['Lamb'][-1] / r['Lamb'][0]
        assert np.max(np.abs(bh - r['betahat'][:, 0])) <= max(1e-9, 1e-15 * cond) * np.max(np.abs(bh))
        Q = res.Q[res.mat_off[c]:res.mat_off[c] + p * p].view(p, p).cpu().numpy().T
        assert np.allclose(Q.T @ Q, np.eye(p), atol=1e-13)
        assert (res.info[c] >> 8) < 40      # Jacobi converged before the sweep cap


@pytest.mark.parametrize('n,m,seed,D', [(441, 2, 4, 300), (3000, 3, 5, 200)])
def test_injected_chain_matches_oracle(engine, phis_cubic, n, m, seed, D):
    x, y, terms, X = make_problem(phis_cubic, n, m, seed)
    setup_engine(engine, phis_cubic, x, y, terms)
    a = atau = 4
    b, btau = fo.default_b_btau(y, a, atau)
    hyp = engine.make_hypers(a, b, atau, btau, b / (1 + a), btau / (1 + atau), D)
    P = engine.P
    sets = [list(range(P)), [0] + list(range(3, P))]
    pre = engine.evaluate(sets, hyp, rng_mode=_lib.RNG_NONE, want_eig=True, refine_tol=None)
    refs, variates, signs = [], [], []
    for c, cols in enumerate(sets):
        r, v = oracle_with_device_gram(engine, X, y, cols, a, b, atau, btau, D, seed + c)
        p = len(cols)
        Q = pre.Q[pre.mat_off[c]:pre.mat_off[c] + p * p].view(p, p).cpu().numpy().T
        signs.append(np.sign(np.sum(Q * r['Q'], axis=0)))
        refs.append(r)
        variates.append(v.reshape(-1))
    res = engine.evaluate(sets, hyp, rng_mode=_lib.RNG_INJECTED, variates=np.concatenate(variates),
                          sign_fix=np.concatenate(signs), want_betas=True, refine_tol=None)
    for c, cols in enumerate(sets):
        r = refs[c]
        bt = res.betas_of(c).cpu().numpy()
        assert np.max(np.abs(bt - r['betas'])) <= 1e-9 * np.max(np.abs(r['betas'])), c
        sg = res.sigs[D * c:D * (c + 1)].cpu().numpy()
        tu = res.taus[D * c:D * (c + 1)].cpu().numpy()
        assert np.allclose(sg, r['sigs'][:, 0], rtol=1e-9, atol=0)
        assert np.allclose(tu, r['taus'][:, 0], rtol=1e-9, atol=0)
        # column statistics used by the selection loop (FR:1656-1658)
        st = res.stats_of(c).cpu().numpy()
        h0, h1 = int(np.ceil(D / 2)), int(np.ceil(D / 2 + 1))
        assert np.allclose(st[0], bt[h1:].mean(axis=0), rtol=1e-12, atol=1e-15)
        assert np.allclose(st[1], bt[h1:].std(axis=0), rtol=1e-10, atol=1e-15)
        assert np.allclose(st[2], bt[h0:].mean(axis=0), rtol=1e-12, atol=1e-15)
        assert (res.info[c] & 1) == 0


def test_philox_chain_posterior_within_monte_carlo_error(engine, phis_cubic):
    """Free-running mode: posterior means of beta and sigma^2 agree with a long oracle chain within 5 MC standard
    errors; different streams give different draws, same stream is reproducible."""
    x, y, terms, X = make_problem(phis_cubic, 2000, 3, 7)
    setup_engine(engine, phis_cubic, x, y, terms)
    a = atau = 4
    b, btau = fo.default_b_btau(y, a, atau)
    D = 4000
    hyp = engine.make_hypers(a, b, atau, btau, b / (1 + a), btau / (1 + atau), D)
    P = engine.P
    cols = list(range(P))
    res = engine.evaluate([cols, cols], hyp, rng_mode=_lib.RNG_PHILOX, seed=1234, stream_ids=[1, 2], want_betas=True)
    res2 = engine.evaluate([cols], hyp, rng_mode=_lib.RNG_PHILOX, seed=1234, stream_ids=[1], want_betas=True)
    b1, b2 = res.betas_of(0).cpu().numpy(), res.betas_of(1).cpu().numpy()
    assert np.array_equal(b1, res2.betas_of(0).cpu().numpy())
    assert not np.array_equal(b1, b2)
    np.random.seed(3)
    r = fo.gibbs_from_X(X, y, a, b, atau, btau, D, b / (1 + a), btau / (1 + atau), y.T.dot(y), literal=False)
    burn = 500
    for got in (b1, b2):
        se = r['betas'][burn:].std(axis=0) / np.sqrt((D - burn) / 4.0)     # generous: assume ESS = N/4
        assert np.all(np.abs(got[burn:].mean(axis=0) - r['betas'][burn:].mean(axis=0)) < 5 * se + 1e-12)
        assert np.allclose(got[burn:].std(axis=0), r['betas'][burn:].std(axis=0), rtol=0.15)
    sg = res.sigs[:D].cpu().numpy()
    assert abs(sg[burn:].mean() - r['sigs'][burn:].mean()) < 0.05 * r['sigs'][burn:].mean()


def test_large_candidate_uses_global_workspace(engine, phis_cubic):
    """p > 119 does not fit W|V in shared memory: the global-memory Jacobi path must give the same answers."""
    rng = np.random.default_rng(9)
    n, m = 6000, 6
    x = rng.random((n, m))
    y = (np.sin(2 * np.pi * x[:, 0]) + x[:, 1] * x[:, 2] + 0.1 * rng.standard_normal(n))[:, None]
    terms = np.vstack([fo.distinct_perms(p).astype(int) for p in
                       ([1, 0, 0, 0, 0, 0], [1, 1, 0, 0, 0, 0], [2, 0, 0, 0, 0, 0], [2, 1, 0, 0, 0, 0],
                        [1, 1, 1, 0, 0, 0], [3, 0, 0, 0, 0, 0], [2, 2, 0, 0, 0, 0], [2, 1, 1, 0, 0, 0])])
    X = np.hstack([np.ones((n, 1)), fo.basis_columns(x, terms, phis_cubic, fo.CUBIC)])
    setup_engine(engine, phis_cubic, x, y, terms)
    P = engine.P
    assert P > 150
    a = atau = 4
    b, btau = fo.default_b_btau(y, a, atau)
    hyp = engine.make_hypers(a, b, atau, btau, b / (1 + a), btau / (1 + atau), 10)
    sets = [list(range(P)), list(range(100))]
    res = engine.evaluate(sets, hyp, rng_mode=_lib.RNG_NONE, want_eig=True, refine_tol=None)
    dtd = y.T.dot(y)
    for c, cols in enumerate(sets):
        r = fo.gibbs_from_X(X[:, cols], y, a, b, atau, btau, 1, 1.0, 1.0, dtd, literal=False,
                            variates=(np.zeros((1, len(cols))), np.ones(1), np.ones(1)))
        assert abs(res.ev[c] - r['ev']) <= 1e-9 * abs(r['ev'])
        lam = res.lamb[res.vec_off[c]:res.vec_off[c] + len(cols)].cpu().numpy()
        assert np.all(np.abs(lam - r['Lamb']) <= 1e-12 * r['Lamb'][-1])


def test_residual_bic_cross_check(engine, phis_cubic):
    """Size-independent property: the Gram-only BIC equals the BIC from an explicit N-length residual pass."""
    x, y, terms, X = make_problem(phis_cubic, 50000, 4, 12, way3=True, ind_max=3)
    setup_engine(engine, phis_cubic, x, y, terms)
    a = atau = 4
    b, btau = fo.default_b_btau(y, a, atau)
    hyp = engine.make_hypers(a, b, atau, btau, b / (1 + a), btau / (1 + atau), 10)
    cols = list(range(engine.P))
    res = engine.evaluate([cols], hyp, rng_mode=_lib.RNG_NONE, refine_tol=None)
    ev2 = engine.residual_bic(cols, res.betahat[:len(cols)])
    assert abs(res.ev[0] - ev2) <= 1e-10 * abs(ev2)


def test_kill_scores_equal_spectral_bic(engine, phis_cubic):
    """fokl_kill_scores (one Cholesky, SSE_{-q} = SSE + b_q^2 / (A^-1)_qq) against the per-candidate spectral path."""
    x, y, terms, X = make_problem(phis_cubic, 8000, 4, 21, way3=True, ind_max=3)
    setup_engine(engine, phis_cubic, x, y, terms)
    a = atau = 4
    b, btau = fo.default_b_btau(y, a, atau)
    hyp = engine.make_hypers(a, b, atau, btau, b / (1 + a), btau / (1 + atau), 10)
    P = engine.P
    model = [0] + list(range(2, P))
    pos = list(range(1, len(model)))
    ev, ok = engine.kill_scores(model, pos, hyp)
    assert ok
    sets = [[c for j, c in enumerate(model) if j != q] for q in pos] + [model]
    ref = engine.evaluate(sets, hyp, rng_mode=_lib.RNG_NONE, refine_tol=None).ev
    assert np.max(np.abs(ev - ref) / np.abs(ref)) < 1e-11
    # a singular Gram (duplicated column) is reported, not silently scored
    engine.append_terms(terms[:1])
    dup = list(range(engine.P))
    _, ok = engine.kill_scores(dup, [1], hyp)
    assert not ok
