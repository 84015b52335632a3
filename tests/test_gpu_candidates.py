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


def test_kill_loop_kernel_matches_host_emulation(engine, phis_cubic):
    """fokl_kill_loop (one CTA, sweep operator) against the single-thread host build of the same math, which
    tests/test_host_emu.py pins to the literal sequential loop of the oracle."""
    import torch
    import emu
    rng = np.random.default_rng(21)
    n, m = 5000, 4
    x = rng.random((n, m))
    y = np.sin(2 * np.pi * x[:, 0]) + x[:, 1] * x[:, 2] + 0.1 * rng.standard_normal(n)
    terms = np.vstack([fo.distinct_perms(p).astype(int) for p in
                       ([1, 0, 0, 0], [1, 1, 0, 0], [2, 0, 0, 0], [2, 1, 0, 0], [1, 1, 1, 0], [2, 1, 1, 0])])
    engine.set_phis(phis_cubic, fo.CUBIC)
    ds = engine.upload(x, y)
    engine.begin_fit(ds)
    engine.append_terms(terms)
    P = engine.P
    G = engine.G[:P, :P].cpu().numpy()
    Xty = engine.Xty[:P].cpu().numpy()
    hyp = engine.make_hypers(4, 1, 4, 1, 1, 1, 10)
    hd = dict(a=4, b=1, atau=4, btau=1, sigsqd0=1, tausqd0=1, yty=engine.yty, sum_y=engine.sum_y, n=n, draws=10)
    for vm, aic in ((12, False), (40, True)):
        cols = list(range(P))
        cand = list(rng.permutation(np.arange(P - vm, P)))
        bv0 = np.sort(rng.random(vm))
        bv1 = rng.random(vm) * 3
        aic_adj = (2 - np.log(n)) if aic else 0.0
        full = float(engine.evaluate([cols], hyp).ev[0]) + aic_adj * P
        pos = [cols.index(c) for c in cand]
        want = emu.kill_loop(G, Xty, cols, pos, bv0, bv1, hd, threshav=0.3, icpt=2.0, evmin=full, aic_adj=aic_adj)
        got = engine.kill_loop(cols, pos, bv0, bv1, hyp, 0.3, 0.5, 2.0, 2.0, full, aic_adj, 0)
        assert got['bad'] == 0 and got['n_acc'] == want['n_acc'] > 0
        assert list(got['acc']) == list(want['acc']) and list(got['calls']) == list(want['calls'])
        assert got['tested'] == want['tested']
        assert np.allclose(got['ev'], want['ev'], rtol=1e-12, atol=0)


@pytest.mark.parametrize('sizes', [(1, 2, 3, 5, 31, 64), (65, 100, 128), (129, 200, 256), (300, 440), (470, 600), (700,)])
def test_eigensolver_all_cluster_classes(engine, sizes):
    """Cholesky + cluster Jacobi (cluster sizes 1, 2, 4, 8, 16) and the global-memory fallback against
    scipy.linalg.eigh on the same Gram bits: eigenvalues to 1e-12 of the largest, orthonormal eigenvectors,
    reconstruction, betahat."""
    import torch
    from scipy.linalg import eigh
    rng = np.random.default_rng(sum(sizes))
    pmax = max(sizes)
    n = 4 * pmax + 50
    X = rng.standard_normal((n, pmax)) * (1.0 + 3.0 * rng.random(pmax))
    X[:, 0] = 1.0
    y = X[:, :min(5, pmax)] @ rng.standard_normal(min(5, pmax)) + 0.1 * rng.standard_normal(n)
    G = X.T @ X
    Xty = X.T @ y
    cap = max(pmax, 64)
    engine.G = torch.zeros((cap, cap), dtype=torch.float64, device=engine.device)
    engine.Xty = torch.zeros(cap, dtype=torch.float64, device=engine.device)
    engine.G[:pmax, :pmax] = torch.from_numpy(G).to(engine.device)
    engine.Xty[:pmax] = torch.from_numpy(Xty).to(engine.device)
    engine.Gcap = cap
    engine.n_global, engine.sum_y, engine.yty = n, float(y.sum()), float(y @ y)
    hyp = engine.make_hypers(4, 1, 4, 1, 1, 1, 10)
    sets = [list(range(p)) for p in sizes]
    res = engine.evaluate(sets, hyp, want_eig=True, refine_tol=None)
    info = res.info
    for c, p in enumerate(sizes):
        lam = res.lamb[res.vec_off[c]:res.vec_off[c] + p].cpu().numpy()
        Q = res.Q[res.mat_off[c]:res.mat_off[c] + p * p].view(p, p).cpu().numpy().T      # columns = eigenvectors
        lam_ref, _ = eigh(G[:p, :p])
        assert (info[c] & 2) == 0, (p, info[c])          # positive definite: no Cholesky failure
        print('eig p', p, 'info', info[c] & 7, 'sweeps', info[c] >> 8)
        assert (info[c] >> 8) < 40
        assert np.all(np.diff(lam) >= 0)
        assert np.max(np.abs(lam - lam_ref)) <= 1e-12 * lam_ref[-1], p
        assert np.max(np.abs(Q.T @ Q - np.eye(p))) < 1e-11, p
        assert np.max(np.abs((Q * lam) @ Q.T - G[:p, :p])) <= 1e-11 * lam_ref[-1], p
        bh = res.betahat[res.vec_off[c]:res.vec_off[c] + p].cpu().numpy()
        bh_ref = np.linalg.solve(G[:p, :p], Xty[:p])
        assert np.max(np.abs(bh - bh_ref)) <= 1e-8 * np.max(np.abs(bh_ref)), p
        r = y - X[:, :p] @ bh_ref
        ev_ref = p * np.log(n) - 2 * (-(n / 2) * np.log(np.var(r)) - (n - 1) / 2)
        assert abs(res.ev[c] - ev_ref) <= 1e-9 * abs(ev_ref), p


def test_eigensolver_not_positive_definite_falls_back(engine):
    """A singular Gram (duplicated column) fails the Cholesky pre-pass and is solved by the two-matrix Jacobi."""
    import torch
    from scipy.linalg import eigh
    rng = np.random.default_rng(3)
    X = rng.standard_normal((50, 8))
    X[:, 5] = X[:, 2]
    G = X.T @ X
    engine.G = torch.zeros((64, 64), dtype=torch.float64, device=engine.device)
    engine.Xty = torch.zeros(64, dtype=torch.float64, device=engine.device)
    engine.G[:8, :8] = torch.from_numpy(G).to(engine.device)
    engine.Gcap = 64
    engine.n_global, engine.sum_y, engine.yty = 50, 1.0, 60.0
    hyp = engine.make_hypers(4, 1, 4, 1, 1, 1, 10)
    res = engine.evaluate([list(range(8))], hyp, want_eig=True, refine_tol=None)
    lam = res.lamb[:8].cpu().numpy()
    lam_ref, _ = eigh(G)
    assert np.max(np.abs(lam - lam_ref)) <= 1e-12 * lam_ref[-1]


def _load_gram(engine, G, Xty, n, y):
    import torch
    pmax = G.shape[0]
    cap = max(pmax, 64)
    engine.G = torch.zeros((cap, cap), dtype=torch.float64, device=engine.device)
    engine.Xty = torch.zeros(cap, dtype=torch.float64, device=engine.device)
    engine.G[:pmax, :pmax] = torch.from_numpy(G).to(engine.device)
    engine.Xty[:pmax] = torch.from_numpy(Xty).to(engine.device)
    engine.Gcap = cap
    engine.n_global, engine.sum_y, engine.yty = n, float(y.sum()), float(y @ y)


def _check_eig(res, c, p, G, Xty, X, y, n):
    from scipy.linalg import eigh
    lam = res.lamb[res.vec_off[c]:res.vec_off[c] + p].cpu().numpy()
    Q = res.Q[res.mat_off[c]:res.mat_off[c] + p * p].view(p, p).cpu().numpy().T      # columns = eigenvectors
    lam_ref, _ = eigh(G[:p, :p])
    assert np.all(np.diff(lam) >= 0)
    assert np.max(np.abs(lam - lam_ref)) <= 1e-12 * lam_ref[-1], p
    assert np.max(np.abs(Q.T @ Q - np.eye(p))) < 1e-11, p
    assert np.max(np.abs((Q * lam) @ Q.T - G[:p, :p])) <= 1e-11 * lam_ref[-1], p
    bh = res.betahat[res.vec_off[c]:res.vec_off[c] + p].cpu().numpy()
    bh_ref = np.linalg.solve(G[:p, :p], Xty[:p])
    assert np.max(np.abs(bh - bh_ref)) <= 1e-8 * np.max(np.abs(bh_ref)), p
    r = y - X[:, :p] @ bh_ref
    ev_ref = p * np.log(n) - 2 * (-(n / 2) * np.log(np.var(r)) - (n - 1) / 2)
    assert abs(res.ev[c] - ev_ref) <= 1e-9 * abs(ev_ref), p


@pytest.mark.parametrize('sizes', [(1024,), (1700,), (2072, 705, 1300)])
def test_blocked_eigensolver_wide_models(engine, sizes):
    """Models too wide for the cluster solver (BASELINE configs[4]: the 3-way substages of a 16-input problem reach
    ~2000 columns): blocked Cholesky + one-sided block Jacobi on the FP64 tensor pipe (csrc/eigbig.cuh) against
    scipy.linalg.eigh on the same Gram bits -- eigenvalues to 1e-12 of the largest, orthonormal eigenvectors,
    reconstruction, betahat and BIC; several models side by side go through the team dispenser."""
    rng = np.random.default_rng(sum(sizes))
    pmax = max(sizes)
    n = 3 * pmax + 50
    X = rng.standard_normal((n, pmax)) * (1.0 + 3.0 * rng.random(pmax))
    X[:, 0] = 1.0
    y = X[:, :5] @ rng.standard_normal(5) + 0.1 * rng.standard_normal(n)
    G, Xty = X.T @ X, X.T @ y
    _load_gram(engine, G, Xty, n, y)
    hyp = engine.make_hypers(4, 1, 4, 1, 1, 1, 10)
    res = engine.evaluate([list(range(p)) for p in sizes], hyp, want_eig=True, refine_tol=None)
    for c, p in enumerate(sizes):
        assert (res.info[c] & 0xff) == 0, (p, res.info[c])
        assert 0 < (res.info[c] >> 8) < 40
        _check_eig(res, c, p, G, Xty, X, y, n)


def test_blocked_eigensolver_small_models_and_graded_columns(engine, monkeypatch):
    """The blocked solver forced onto small models (every block-count / padding case: p = 2 ... 300, ragged last
    block, odd block counts) with strongly graded, correlated columns (condition ~1e8); a subset model (scattered
    column list) as well."""
    monkeypatch.setenv('FOKL_EIGB_MIN_P', '2')
    rng = np.random.default_rng(11)
    pmax = 300
    n = 2000
    u = rng.random((n, 6))
    X = np.empty((n, pmax))
    for j in range(pmax):
        k = rng.choice(6, size=2, replace=False)
        X[:, j] = (u[:, k[0]] ** (1 + j % 4)) * np.cos(3 * u[:, k[1]] * (1 + j % 3)) * 10.0 ** (-4 * rng.random()) \
            + 1e-3 * rng.standard_normal(n)
    X[:, 0] = 1.0
    y = X[:, :5] @ rng.standard_normal(5) + 0.1 * rng.standard_normal(n)
    G, Xty = X.T @ X, X.T @ y
    _load_gram(engine, G, Xty, n, y)
    hyp = engine.make_hypers(4, 1, 4, 1, 1, 1, 10)
    sizes = (2, 3, 8, 9, 16, 17, 31, 33, 100, 129, 300)
    res = engine.evaluate([list(range(p)) for p in sizes], hyp, want_eig=True, refine_tol=None)
    from scipy.linalg import eigh
    for c, p in enumerate(sizes):
        assert (res.info[c] & 0xff) == 0, (p, res.info[c])
        lam = res.lamb[res.vec_off[c]:res.vec_off[c] + p].cpu().numpy()
        Q = res.Q[res.mat_off[c]:res.mat_off[c] + p * p].view(p, p).cpu().numpy().T
        lam_ref, _ = eigh(G[:p, :p])
        # one-sided Jacobi on the Cholesky factor: small eigenvalues to high RELATIVE accuracy
        assert np.max(np.abs(lam - lam_ref) / lam_ref) <= 1e-8, p
        assert np.max(np.abs(lam - lam_ref)) <= 1e-12 * lam_ref[-1], p
        assert np.max(np.abs(Q.T @ Q - np.eye(p))) < 1e-11, p
    cols = [0] + sorted(rng.choice(np.arange(1, pmax), size=150, replace=False).tolist())
    res = engine.evaluate([cols], hyp, want_eig=True, refine_tol=None)
    lam = res.lamb[:len(cols)].cpu().numpy()
    lam_ref, _ = eigh(G[np.ix_(cols, cols)])
    assert np.max(np.abs(lam - lam_ref)) <= 1e-12 * lam_ref[-1]


def test_blocked_eigensolver_not_positive_definite_falls_back(engine, monkeypatch):
    """A singular Gram fails the blocked Cholesky (status -1) and is handed to the two-matrix Jacobi."""
    monkeypatch.setenv('FOKL_EIGB_MIN_P', '2')
    from scipy.linalg import eigh
    rng = np.random.default_rng(3)
    X = rng.standard_normal((200, 70))
    X[:, 45] = X[:, 2]
    y = rng.standard_normal(200)
    G = X.T @ X
    # (an exactly duplicated column leaves a pivot of +-1 ulp, which may pass the factorisation: make the matrix
    # indefinite beyond rounding so that the fall-back is what runs)
    G[45, 45] *= 1.0 - 1e-6
    _load_gram(engine, G, X.T @ y, 200, y)
    hyp = engine.make_hypers(4, 1, 4, 1, 1, 1, 10)
    res = engine.evaluate([list(range(70)), list(range(40))], hyp, want_eig=True, refine_tol=None)
    assert res.info[0] & 2 and not (res.info[1] & 2)
    for c, p in enumerate((70, 40)):
        lam = res.lamb[res.vec_off[c]:res.vec_off[c] + p].cpu().numpy()
        lam_ref, _ = eigh(G[:p, :p])
        assert np.max(np.abs(lam - lam_ref)) <= 1e-12 * lam_ref[-1]


@pytest.mark.parametrize('p', [60, 150, 400, 1100])
def test_whole_device_kill_loop_equals_single_cta_kernel(engine, monkeypatch, p):
    """fokl_kill_loop for models whose tableau does not fit one SM (csrc/killbig.cuh: rows dealt to one CTA per SM,
    look-ahead pivot row, one grid barrier per pivot) takes the single-CTA kernel's decisions and BICs bit for bit."""
    rng = np.random.default_rng(p)
    n = 3 * p + 50
    X = rng.standard_normal((n, p)) * (1.0 + 3.0 * rng.random(p))
    X[:, 0] = 1.0
    y = X[:, :min(40, p)] @ rng.standard_normal(min(40, p)) + 0.5 * rng.standard_normal(n)
    _load_gram(engine, X.T @ X, X.T @ y, n, y)
    hyp = engine.make_hypers(4, 1, 4, 1, 1, 1, 10)
    cols = list(range(p))
    vm = (3 * p) // 4
    pos = [int(c) for c in rng.permutation(np.arange(p - vm, p))]
    bv0 = np.sort(rng.random(vm))
    bv1 = rng.random(vm) * 3
    full = float(engine.evaluate([cols], hyp, refine_tol=None).ev[0])
    out = {}
    for name, knob in (('big', '2'), ('one', '0')):
        monkeypatch.setenv('FOKL_KILL_BIG_MIN_P', knob)
        out[name] = engine.kill_loop(cols, pos, bv0, bv1, hyp, 0.3, 0.5, 2.0, 2.0, full, 0.0, 3)
    big, one = out['big'], out['one']
    assert big['bad'] == one['bad'] == 0 and big['n_acc'] == one['n_acc'] > 0
    assert big['tested'] == one['tested']
    assert np.array_equal(big['acc'], one['acc']) and np.array_equal(big['calls'], one['calls'])
    assert np.array_equal(big['ev'], one['ev'])
    # a singular Gram is reported by both
    G2 = X.T @ X
    G2[:, 5] = G2[:, 2]
    G2[5, :] = G2[2, :]
    _load_gram(engine, G2, X.T @ y, n, y)
    for knob in ('2', '0'):
        monkeypatch.setenv('FOKL_KILL_BIG_MIN_P', knob)
        assert engine.kill_loop(cols, pos, bv0, bv1, hyp, 0.3, 0.5, 2.0, 2.0, full, 0.0, 0)['bad'] != 0


@pytest.mark.parametrize('p,big', [(60, '0'), (100, '0'), (226, '0'), (150, '2'), (700, '2')])
def test_kill_loop_from_eigendecomposition_equals_sequential_pivots(engine, monkeypatch, p, big):
    """fokl_kill_params.lamb / Qt: the tableau the loop starts from is formed from the model's eigendecomposition
    (csrc/candidates.cu kill_tableau_kernel) instead of by p sequential pivots -- same decisions, BICs to 1e-9; an
    ill-conditioned model takes the sequential form (and reports a singular Gram as before)."""
    rng = np.random.default_rng(100 + p)
    n = 3 * p + 50
    X = rng.standard_normal((n, p)) * (1.0 + 3.0 * rng.random(p))
    X[:, 0] = 1.0
    y = X[:, :min(40, p)] @ rng.standard_normal(min(40, p)) + 0.5 * rng.standard_normal(n)
    _load_gram(engine, X.T @ X, X.T @ y, n, y)
    hyp = engine.make_hypers(4, 1, 4, 1, 1, 1, 10)
    cols = list(range(p))
    vm = (3 * p) // 4
    pos = [int(c) for c in rng.permutation(np.arange(p - vm, p))]
    bv0 = np.sort(rng.random(vm))
    bv1 = rng.random(vm) * 3
    res = engine.evaluate([cols], hyp, want_eig=True, refine_tol=None)
    full = float(res.ev[0])
    monkeypatch.setenv('FOKL_KILL_BIG_MIN_P', big)
    seq = engine.kill_loop(cols, pos, bv0, bv1, hyp, 0.3, 0.5, 2.0, 2.0, full, 0.0, 1)
    eig = engine.kill_loop_launch(cols, pos, bv0, bv1, hyp, 0.3, 0.5, 2.0, 2.0, full, 0.0, 1,
                                  eig=(res.lamb, res.Q)).finish()
    assert seq['bad'] == eig['bad'] == 0 and seq['n_acc'] == eig['n_acc'] > 0
    assert seq['tested'] == eig['tested']
    assert np.array_equal(seq['acc'], eig['acc']) and np.array_equal(seq['calls'], eig['calls'])
    # (north_star: BIC within rtol 1e-9 of the reference's float64 path)
    assert np.allclose(seq['ev'], eig['ev'], rtol=1e-9, atol=0), np.max(np.abs(seq['ev'] - eig['ev']) / np.abs(seq['ev']))
    # the environment switch really selects the sequential form (bit-identical to it)
    monkeypatch.setenv('FOKL_KILL_NO_EIG', '1')
    off = engine.kill_loop_launch(cols, pos, bv0, bv1, hyp, 0.3, 0.5, 2.0, 2.0, full, 0.0, 1,
                                  eig=(res.lamb, res.Q)).finish()
    assert np.array_equal(off['ev'], seq['ev'])
    monkeypatch.delenv('FOKL_KILL_NO_EIG')
    # nearly singular model: lam_min <= 1e-10 max diag -> the sequential form runs, flags the Gram as it did before
    X2 = X.copy()
    X2[:, 5] = X2[:, 2] * (1.0 + 1e-9)
    _load_gram(engine, X2.T @ X2, X2.T @ y, n, y)
    res2 = engine.evaluate([cols], hyp, want_eig=True, refine_tol=None)
    seq2 = engine.kill_loop(cols, pos, bv0, bv1, hyp, 0.3, 0.5, 2.0, 2.0, full, 0.0, 0)
    if res2.lamb is not None and not (int(res2.info[0]) & 6):
        eig2 = engine.kill_loop_launch(cols, pos, bv0, bv1, hyp, 0.3, 0.5, 2.0, 2.0, full, 0.0, 0,
                                       eig=(res2.lamb, res2.Q)).finish()
        assert eig2['bad'] == seq2['bad'] and eig2['n_acc'] == seq2['n_acc']
        assert np.array_equal(eig2['ev'], seq2['ev'])


@pytest.mark.parametrize('p0,n_models,kind', [(40, 12, 'gauss'), (300, 25, 'graded'), (900, 30, 'gauss')])
def test_nested_chains_match_one_eigensolver_per_model(engine, p0, n_models, kind):
    """csrc/nested.cu: the spectral decomposition of a model carried to its sub-models by secular-equation steps
    (bisection on the offset from the nearer pole, Gu / Eisenstat weights) + the intercept-only chains: the posterior
    mean of the intercept of every nested model equals the ordinary path's (one eigensolver + chain + betas + column
    statistics per model, same Philox streams) to 1e-10; the eigenvalues of a step equal scipy's."""
    import torch
    from scipy.linalg import eigh
    rng = np.random.default_rng(p0)
    n = 4 * p0 + 50
    X = rng.standard_normal((n, p0)) * (1.0 + 3.0 * rng.random(p0))
    if kind == 'graded':
        X = X * 10.0 ** (-3 * rng.random(p0)) + 0.3 * X[:, [1]]
    X[:, 0] = 1.0
    y = 2.0 + X[:, 1:6] @ rng.standard_normal(5) + 0.3 * rng.standard_normal(n)
    G, Xty = X.T @ X, X.T @ y
    _load_gram(engine, G, Xty, n, y)
    hyp = engine.make_hypers(4, 1, 4, 1, 1, 1, 400)
    drop = rng.permutation(np.arange(1, p0))[:2 * n_models]
    sets = [np.array([c for c in range(p0) if c not in set(drop[:2 * k + 1].tolist())], dtype=np.int32)
            for k in range(n_models)]                       # one, then two more columns removed per model
    ids = np.arange(7, 7 + n_models, dtype=np.uint64)
    got = engine.nested_chains_launch(sets, hyp, 99, ids).finish()
    assert got['ok'] and got['status'] == 0
    cold = engine.evaluate(sets, hyp, rng_mode=_lib.RNG_PHILOX, seed=99, stream_ids=ids, refine_tol=None)
    st = cold.stats.cpu().numpy()
    want = np.array([st[3 * cold.vec_off[c] + 2 * cold.p[c]] for c in range(n_models)])
    assert np.max(np.abs(got['mean0'] - want) / np.abs(want)) < 1e-10
    # one step against LAPACK: eigenvalues of the compressed matrix, orthonormal rows, and Z' Q diagonalises it
    lam, Q = eigh(G)
    m = int(drop[0])
    f64 = dict(dtype=torch.float64, device=engine.device)
    lam_d, u_d = torch.from_numpy(lam).to(engine.device), torch.from_numpy(np.ascontiguousarray(Q[m, :])).to(engine.device)
    mu, zt = torch.empty(p0 - 1, **f64), torch.empty((p0 - 1, p0), **f64)
    work, status = torch.empty(4 * p0, **f64), torch.zeros(1, dtype=torch.int32, device=engine.device)
    engine._ck(engine.lib.fokl_secular_step(engine.ctx, lam_d.data_ptr(), u_d.data_ptr(), p0, mu.data_ptr(), zt.data_ptr(),
                                            p0, work.data_ptr(), status.data_ptr()))
    keep = [c for c in range(p0) if c != m]
    mu_ref = eigh(G[np.ix_(keep, keep)], eigvals_only=True)
    assert int(status.item()) == 0
    assert np.max(np.abs(mu.cpu().numpy() - mu_ref)) <= 1e-12 * mu_ref[-1]
    Z = zt.cpu().numpy()
    assert np.max(np.abs(Z @ Z.T - np.eye(p0 - 1))) < 1e-11
    Qn = Z @ Q.T                                            # rows = new eigenvectors over the p0 variables
    assert np.max(np.abs(Qn[:, m])) < 1e-11                 # no component on the removed variable
    Qk = Qn[:, keep]
    assert np.max(np.abs(Qk @ G[np.ix_(keep, keep)] @ Qk.T - np.diag(mu_ref))) <= 1e-10 * mu_ref[-1]


def test_nested_chains_report_equal_eigenvalues(engine):
    """Two exactly equal eigenvalues leave no interval for a root of the secular equation: the step must say so
    (status 1, ok False) -- the selection loop then evaluates the batch by the ordinary path."""
    n, p0 = 1000, 12
    G = np.diag([float(n)] + [3.0] * (p0 - 1))          # orthogonal columns of equal norm: an 11-fold eigenvalue
    Xty = np.arange(1, p0 + 1, dtype=np.float64)
    y = np.zeros(n)
    y[0] = 30.0
    _load_gram(engine, G, Xty, n, y)
    engine.sum_y, engine.yty = 1.0, 900.0
    hyp = engine.make_hypers(4, 1, 4, 1, 1, 1, 50)
    sets = [np.array([c for c in range(p0) if c not in (3, 5, 7, 9)[:k + 1]], dtype=np.int32) for k in range(4)]
    got = engine.nested_chains_launch(sets, hyp, 1, np.arange(4, dtype=np.uint64)).finish()
    assert not got['ok'] and (got['status'] & 1)
