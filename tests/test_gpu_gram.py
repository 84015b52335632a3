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


@pytest.mark.parametrize('joint', [False, True])
@pytest.mark.parametrize('n,p_old,c,gap', [(10, 1, 2, 3), (2500, 65, 67, 9), (100000, 30, 40, 168), (300001, 20, 168, 56)])
def test_gram_block_with_gap_and_cross_only(engine, n, p_old, c, gap, joint, monkeypatch):
    """fokl_gram_update_ex: the new block stored `gap` columns above the old ones (built ahead of time, above columns
    the previous substage may still delete -- here poisoned with NaN) gives the SAME BITS as the contiguous layout;
    FOKL_GRAM_CROSS_ONLY leaves the new x new rows zero and reproduces the (old | y) x new rows bit for bit."""
    import torch
    from FoKL import _lib
    if joint:
        monkeypatch.setenv('FOKL_GRAM_KERNEL', 'tma')
    rng = np.random.default_rng(n + p_old + gap)
    Xh = rng.standard_normal((n, p_old + c))
    Xh[:, 0] = 1.0
    yh = rng.standard_normal(n)
    ref = gram_block(engine, Xh, yh, p_old, c, joint)
    ld = ((n + 15) // 16) * 16
    full = torch.full((p_old + gap + c + 1, ld), float('nan'), dtype=torch.float64, device=engine.device)
    full[:, :n] = 0.0
    X = full[1:]
    X[:p_old, :n] = torch.from_numpy(np.ascontiguousarray(Xh[:, :p_old].T)).to(engine.device)
    X[p_old:p_old + gap, :n] = float('nan')                      # dead columns: never read as data
    X[p_old + gap:, :n] = torch.from_numpy(np.ascontiguousarray(Xh[:, p_old:].T)).to(engine.device)
    if joint:
        y = full[0]
    else:
        y = torch.zeros(ld, dtype=torch.float64, device=engine.device)
    y[:n] = torch.from_numpy(yh).to(engine.device)
    for flags in (0, _lib.GRAM_CROSS_ONLY):
        block = torch.full(((p_old + c + 1) * c,), np.nan, dtype=torch.float64, device=engine.device)
        engine._ck(engine.lib.fokl_gram_update_ex(engine.ctx, X.data_ptr(), ld, n, p_old, c, p_old + gap, flags,
                                                  y.data_ptr(), block.data_ptr()))
        engine.synchronize()
        got = block.cpu().numpy().reshape(p_old + c + 1, c)
        if flags == 0:
            assert np.array_equal(got, ref)
        else:
            # another plan (fewer blocks per tile, another k-split): same values to rounding, not the same bits
            A = np.hstack([Xh[:, :p_old], yh[:, None]])
            scale = np.sqrt(np.outer(np.sum(A * A, axis=0), np.sum(Xh[:, p_old:] ** 2, axis=0)))
            rows = list(range(p_old)) + [p_old + c]
            assert np.all(np.abs(got[rows] - ref[rows]) <= 1e-12 * scale)
            assert np.all(got[p_old:p_old + c] == 0.0)


def test_engine_build_ahead_layout(engine, phis_cubic):
    """Engine-level: a block started ahead of time above a gap, committed after a compaction, with the cross block
    against the survivors: G / Xty equal numpy and equal the synchronous two-part build bit for bit; residual pass and
    later compactions see the right physical columns."""
    import fokl_oracle as fo
    rng = np.random.default_rng(12)
    n, m = 5000, 3
    x = rng.random((n, m))
    y = rng.standard_normal(n)
    engine.set_phis(phis_cubic, fo.CUBIC)
    ds = engine.upload(x, y)
    t1 = fo.distinct_perms([1, 0, 0]).astype(int)
    t2 = fo.distinct_perms([1, 1, 0]).astype(int)
    t3 = fo.distinct_perms([2, 1, 0]).astype(int)
    t4 = fo.distinct_perms([1, 1, 1]).astype(int)
    results = []
    for ahead in (True, False):
        engine.begin_fit(ds)
        engine.append_terms(t1)                        # P = 4
        p0 = engine.P
        engine.append_terms(t2, p_stable=p0)           # P = 7, synchronous two-part build
        if ahead:
            assert engine.prefetch_terms(t3, p0) is not None      # columns 0 .. 3 are final whatever happens to t2
        keep = [0, 1, 2, 3, 5]                         # t2 loses two columns
        engine.compact(keep)
        p1 = engine.P
        engine.append_terms(t3, p_stable=p0)           # commits the block built ahead (or builds it now)
        if ahead:
            assert engine.gap > 0
            assert engine.prefetch_terms(t4, p1) is not None
        keep2 = list(range(p1)) + [p1 + 1, p1 + 4]
        cols = np.array(keep2, dtype=np.int32)
        bh = engine.evaluate([keep2], engine.make_hypers(4, 1, 4, 1, 1, 1, 10), refine_tol=None).betahat
        ev_res = engine.residual_bic(cols, bh[:len(keep2)])
        engine.compact(keep2)
        engine.append_terms(t4, p_stable=p1)
        P = engine.P
        results.append((engine.G[:P, :P].cpu().numpy(), engine.Xty[:P].cpu().numpy(),
                        engine.X[:P, :n].t().cpu().numpy() if engine.gap == 0 else
                        engine.X[torch_idx(engine, P), :n].t().cpu().numpy(), ev_res))
    terms = np.vstack([t1, t2[[1]], t3[[1, 4]], t4])
    Xh = np.hstack([np.ones((n, 1)), fo.basis_columns(x, terms, phis_cubic, fo.CUBIC)])
    for G, Xty, Xd, _ in results:
        assert np.array_equal(Xd, Xh)
        assert np.allclose(G, Xh.T @ Xh, rtol=1e-12, atol=1e-9) and np.array_equal(G, G.T)
        assert np.allclose(Xty, Xh.T @ y, rtol=1e-10, atol=1e-9)
    assert np.array_equal(results[0][0], results[1][0]) and np.array_equal(results[0][1], results[1][1])
    assert results[0][3] == results[1][3]
    engine.drop_prefetch()


def torch_idx(engine, P):
    import torch
    return torch.as_tensor(engine.phys_cols(np.arange(P)).astype(np.int64), device=engine.device)
