"""K2 (fokl_gram_update + scatter/compact) against numpy: rtol 1e-12 (summation order differs from OpenBLAS)."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def gram_block(engine, Xh, yh, p_old, c, joint=False):
    """joint: y stored in the row in front of X (the engine's [y | X] buffer -> tensor-map TMA kernel); otherwise y is
    a separate buffer (mbarrier / cp.async kernels)."""
    import torch
    n, p = Xh.shape
    ld = ((n + 15) // 16) * 16
    full = torch.zeros((p + 1, ld), dtype=torch.float64, device=engine.device)
    X = full[1:]
    X[:, :n] = torch.from_numpy(np.ascontiguousarray(Xh.T)).to(engine.device)
    if joint:
        y = full[0]
        # rows n .. ld of the buffer are never read as data: poison them
        full[:, n:] = float('nan')
    else:
        y = torch.zeros(ld, dtype=torch.float64, device=engine.device)
    y[:n] = torch.from_numpy(yh).to(engine.device)
    block = torch.full(((p_old + c + 1) * c,), np.nan, dtype=torch.float64, device=engine.device)
    engine._ck(engine.lib.fokl_gram_update(engine.ctx, X.data_ptr(), ld, n, p_old, c, y.data_ptr(), block.data_ptr()))
    engine.synchronize()
    return block.cpu().numpy().reshape(p_old + c + 1, c)


@pytest.mark.parametrize('joint', [False, True])
@pytest.mark.parametrize('n,p_old,c', [(1, 1, 1), (10, 1, 2), (33, 3, 5), (1000, 9, 8), (4097, 63, 1), (4097, 64, 64),
                                        (2500, 65, 67), (100000, 130, 40), (300001, 20, 168)])
def test_gram_block_matches_numpy(engine, n, p_old, c, joint, monkeypatch):
    if joint:       # every shape through the tensor-map kernel (the library would pick it for wide tiles only)
        monkeypatch.setenv('FOKL_GRAM_KERNEL', 'tma')
    rng = np.random.default_rng(n + p_old)
    Xh = rng.standard_normal((n, p_old + c))
    Xh[:, 0] = 1.0
    yh = rng.standard_normal(n)
    got = gram_block(engine, Xh, yh, p_old, c, joint)
    A = np.hstack([Xh, yh[:, None]])
    ref = A.T @ Xh[:, p_old:]
    scale = np.sqrt(np.outer(np.sum(A * A, axis=0), np.sum(Xh[:, p_old:] ** 2, axis=0)))
    ok = np.abs(got - ref) <= 1e-12 * scale + 1e-300
    # entries strictly below the diagonal of the symmetric X_new' X_new part are optional: the kernel skips 8 x 8
    # fragments that lie entirely there (they read back as 0) and fokl_gram_scatter mirrors the upper triangle
    a, j = np.meshgrid(np.arange(p_old + c + 1) - p_old, np.arange(c), indexing='ij')
    below = (a > j) & (a < c)
    assert np.all(ok | (below & (got == 0.0)))
    assert np.all(ok[~below])
    # deterministic: a second run gives identical bits
    again = gram_block(engine, Xh, yh, p_old, c, joint)
    assert np.array_equal(got, again)
    # scattered into the master Gram the block is exactly symmetric and complete
    import torch
    P = p_old + c
    G = torch.zeros((P, P), dtype=torch.float64, device=engine.device)
    Xty = torch.zeros(P, dtype=torch.float64, device=engine.device)
    blk = torch.from_numpy(got.reshape(-1)).to(engine.device)
    engine._ck(engine.lib.fokl_gram_scatter(engine.ctx, blk.data_ptr(), p_old, c, G.data_ptr(), P, Xty.data_ptr()))
    Gh = G.cpu().numpy()
    full = Xh.T @ Xh
    assert np.array_equal(Gh[:, p_old:], Gh[p_old:, :].T)
    sc = np.sqrt(np.outer(np.diag(full), np.diag(full)))
    assert np.all(np.abs(Gh[:, p_old:] - full[:, p_old:]) <= 1e-12 * sc[:, p_old:])
    assert np.allclose(Xty.cpu().numpy()[p_old:], Xh[:, p_old:].T @ yh, rtol=1e-10, atol=1e-9)


def test_engine_gram_state_append_and_compact(engine, phis_cubic):
    """Engine-level: G, Xty after appends and a compaction equal numpy on the oracle's X."""
    import fokl_oracle as fo
    rng = np.random.default_rng(11)
    n, m = 3000, 3
    x = rng.random((n, m))
    y = rng.standard_normal(n)
    engine.set_phis(phis_cubic, fo.CUBIC)
    ds = engine.upload(x, y)
    engine.begin_fit(ds)
    t1 = fo.distinct_perms([1, 0, 0]).astype(int)
    t2 = fo.distinct_perms([1, 1, 0]).astype(int)
    engine.append_terms(t1)
    engine.append_terms(t2)
    terms = np.vstack([t1, t2])
    Xh = np.hstack([np.ones((n, 1)), fo.basis_columns(x, terms, phis_cubic, fo.CUBIC)])
    P = engine.P
    assert P == Xh.shape[1]
    G = engine.G[:P, :P].cpu().numpy()
    Xty = engine.Xty[:P].cpu().numpy()
    assert np.allclose(G, Xh.T @ Xh, rtol=1e-12, atol=1e-9)
    assert np.allclose(Xty, Xh.T @ y, rtol=1e-10, atol=1e-9)
    assert np.array_equal(G, G.T)
    assert abs(engine.sum_y - y.sum()) < 1e-9 and abs(engine.yty - y @ y) < 1e-9 and engine.n_global == n
    keep = [0, 1, 3, 4, 6]
    engine.compact(keep)
    G2 = engine.G[:5, :5].cpu().numpy()
    assert np.array_equal(G2, G[np.ix_(keep, keep)])
    assert np.array_equal(engine.Xty[:5].cpu().numpy(), Xty[keep])
    Xd = engine.X[:5, :n].t().cpu().numpy()
    assert np.array_equal(Xd, Xh[:, keep])
