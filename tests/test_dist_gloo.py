"""World-size-2 `gloo` tests (CPU) of the host-side multi-rank logic: row sharding of the
synthetic workloads, the SUM allreduce of partial Gram blocks / data moments, the MIN/MAX allreduce behind
`clean`'s normalisation bounds and the seed broadcast.  The device kernels are not involved; the partial Gram
of a shard is formed with numpy exactly as `fokl_gram_update` defines its output block."""
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, out_dir):
    import torch
    import torch.distributed as dist
    for p in (os.path.join(ROOT, 'fokl-gpy_b200'), os.path.join(ROOT, 'oracle'), ROOT):
        if p not in sys.path:
            sys.path.insert(0, p)
    os.environ['MASTER_ADDR'] = '127.0.0.1'
    os.environ['MASTER_PORT'] = str(port)
    dist.init_process_group('gloo', rank=rank, world_size=world)
    try:
        import bench_data
        import fokl_oracle as fo
        from FoKL import FoKLRoutines as FR
        from FoKL import _selection
        n_total = 3001                      # not divisible by the world size: ragged last shard
        per = -(-n_total // world)
        lo, hi = rank * per, min((rank + 1) * per, n_total)
        x, y = bench_data.make_rows('cfg3', lo, hi, n_total=n_total)

        # (1) normalisation bounds: every rank must see the global per-column min / max
        raw = 3.0 * x - 1.0
        bounds = np.array(FR._column_minmax_host(raw))

        # (2) Gram block of the shard: [X_old X_new y]' X_new, then one SUM allreduce (Engine._append_built)
        from FoKL import getKernels
        phis = getKernels.bernoulli()
        terms_old = np.array([[1, 0, 0, 0], [0, 1, 0, 0]])
        terms_new = np.array([[0, 0, 1, 0], [0, 0, 0, 1], [1, 1, 0, 0]])
        Xo = np.hstack([np.ones((hi - lo, 1)), fo.basis_columns(x, terms_old, phis, fo.BERNOULLI)])
        Xn = fo.basis_columns(x, terms_new, phis, fo.BERNOULLI)
        block = np.vstack([Xo.T @ Xn, Xn.T @ Xn, y[None, :] @ Xn])
        mom = np.array([float(hi - lo), y.sum(), y @ y])
        tb, tm = torch.from_numpy(block.copy()), torch.from_numpy(mom.copy())
        dist.all_reduce(tb, op=dist.ReduceOp.SUM)
        dist.all_reduce(tm, op=dist.ReduceOp.SUM)

        # (3) the Philox seed is rank 0's (PhiloxVariates broadcast)
        class _Eng:
            pass
        eng = _Eng()
        eng.torch, eng.dist, eng.group, eng.device = torch, dist, None, torch.device('cpu')
        np.random.seed(100 + rank)          # deliberately different host RNG state per rank
        seed = _selection.PhiloxVariates(eng).seed
        np.savez(os.path.join(out_dir, 'rank%d.npz' % rank), bounds=bounds, block=tb.numpy(), mom=tm.numpy(),
                 seed=np.array([seed], dtype=np.int64), rows=np.array([lo, hi]))
    finally:
        dist.destroy_process_group()


@pytest.mark.timeout(300)
def test_two_rank_row_sharding_allreduce(tmp_path):
    import torch.multiprocessing as mp
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, 'oracle'))
    sys.path.insert(0, os.path.join(ROOT, 'fokl-gpy_b200'))
    import bench_data
    import fokl_oracle as fo
    from FoKL import getKernels
    world = 2
    port = 29500 + (os.getpid() % 2000)
    mp.spawn(_worker, args=(world, port, str(tmp_path)), nprocs=world, join=True)
    r = [np.load(os.path.join(str(tmp_path), 'rank%d.npz' % k)) for k in range(world)]

    n_total = 3001
    x, y = bench_data.make_rows('cfg3', 0, n_total, n_total=n_total)
    # shards tile the dataset exactly
    assert r[0]['rows'][0] == 0 and r[0]['rows'][1] == r[1]['rows'][0] and r[1]['rows'][1] == n_total
    raw = 3.0 * x - 1.0
    want = np.array([[raw[:, k].min(), raw[:, k].max()] for k in range(raw.shape[1])])
    for k in range(world):
        assert np.array_equal(r[k]['bounds'], want)
    phis = getKernels.bernoulli()
    Xo = np.hstack([np.ones((n_total, 1)),
                    fo.basis_columns(x, np.array([[1, 0, 0, 0], [0, 1, 0, 0]]), phis, fo.BERNOULLI)])
    Xn = fo.basis_columns(x, np.array([[0, 0, 1, 0], [0, 0, 0, 1], [1, 1, 0, 0]]), phis, fo.BERNOULLI)
    block = np.vstack([Xo.T @ Xn, Xn.T @ Xn, y[None, :] @ Xn])
    for k in range(world):
        assert np.allclose(r[k]['block'], block, rtol=1e-12, atol=1e-12 * np.abs(block).max())
        assert np.allclose(r[k]['mom'], [n_total, y.sum(), y @ y], rtol=1e-13)
    # allreduce results are bit-identical on both ranks (replicated candidate evaluation relies on it)
    assert np.array_equal(r[0]['block'], r[1]['block'])
    assert r[0]['seed'][0] == r[1]['seed'][0]


def _select_worker(rank, world, port, out_dir):
    import torch.distributed as dist
    for p in (os.path.join(ROOT, 'fokl-gpy_b200'), os.path.join(ROOT, 'oracle'), os.path.join(ROOT, 'tests'), ROOT):
        if p not in sys.path:
            sys.path.insert(0, p)
    os.environ['MASTER_ADDR'] = '127.0.0.1'
    os.environ['MASTER_PORT'] = str(port)
    dist.init_process_group('gloo', rank=rank, world_size=world)
    try:
        import spline_table
        from test_selection_mock import select_on_mock, synthetic
        phis = spline_table.to_phis(np.load(os.path.join(ROOT, 'tests', 'golden', 'phis_cubic_48.npy')))
        x, y = synthetic(401, 3, 9)                       # ragged shards
        per = -(-len(y) // world)
        lo, hi = rank * per, min((rank + 1) * per, len(y))
        for tag, kw in (('pipe', {}), ('seq', dict(pipeline=False))):
            np.random.seed(50 + rank)                     # the Philox seed must come from rank 0 all the same
            out, eng = select_on_mock(x[lo:hi], y[lo:hi], phis, dist=dist, **kw)
            np.savez(os.path.join(out_dir, '%s_rank%d.npz' % (tag, rank)), mtx=out['mtx'], evs=out['evs'],
                     betas=out['betas'], n_gibbs=out['n_gibbs'],
                     side=sum(1 for c in eng.calls if c == ('launch', 'side')),
                     evaluated=sum(len(c[1]) for c in eng.calls if c[0] == 'evaluate'))
    finally:
        dist.destroy_process_group()


@pytest.mark.timeout(300)
def test_two_rank_selection_loop_on_the_stand_in_engine(tmp_path, phis_cubic):
    """The whole selection loop on two gloo ranks (row shards; tests/mock_engine.py): Gram blocks summed over the
    ranks, the chains of the accepted models dealt between them (one allreduce of the scalars that drive the loop, the
    accepted model's draws broadcast from their owner), in the pipelined and in the sequential form.  Both ranks must
    hold the same fit, it must be the single-rank fit of the whole dataset, and each rank must have evaluated only
    its share of the models."""
    import torch.multiprocessing as mp
    from test_selection_mock import select_on_mock, synthetic
    world = 2
    port = 31500 + (os.getpid() % 2000)
    mp.spawn(_select_worker, args=(world, port, str(tmp_path)), nprocs=world, join=True)
    x, y = synthetic(401, 3, 9)
    np.random.seed(50)
    single, eng1 = select_on_mock(x, y, phis_cubic)
    single_evaluated = sum(len(c[1]) for c in eng1.calls if c[0] == 'evaluate')
    for tag in ('pipe', 'seq'):
        r = [np.load(os.path.join(str(tmp_path), '%s_rank%d.npz' % (tag, k))) for k in range(world)]
        for key in ('mtx', 'evs', 'betas', 'n_gibbs'):
            assert np.array_equal(r[0][key], r[1][key]), (tag, key)
        assert np.array_equal(r[0]['mtx'], single['mtx']) and int(r[0]['n_gibbs']) == single['n_gibbs']
        assert np.allclose(r[0]['evs'], single['evs'], rtol=1e-9, atol=0)
        assert np.allclose(r[0]['betas'].mean(axis=0), single['betas'].mean(axis=0), rtol=1e-6, atol=1e-9)
        assert min(int(r[0]['evaluated']), int(r[1]['evaluated'])) < single_evaluated      # the chains were dealt
    assert int(np.load(os.path.join(str(tmp_path), 'pipe_rank0.npz'))['side']) > 0
    assert int(np.load(os.path.join(str(tmp_path), 'seq_rank0.npz'))['side']) == 0


def test_shards_are_independent_of_world_size():
    sys.path.insert(0, ROOT)
    import bench_data
    full_x, full_y = bench_data.make_rows('cfg4', 0, 1000, n_total=1000)
    for world in (2, 4, 8):
        per = -(-1000 // world)
        xs, ys = zip(*[bench_data.make_rows('cfg4', k * per, min((k + 1) * per, 1000), n_total=1000)
                       for k in range(world)])
        assert np.array_equal(np.concatenate(xs), full_x)
        assert np.array_equal(np.concatenate(ys), full_y)


def _public_api_worker(rank, world, port, out_dir):
    import types

    import torch.distributed as dist
    for p in (os.path.join(ROOT, 'fokl-gpy_b200'), os.path.join(ROOT, 'oracle'), os.path.join(ROOT, 'tests'), ROOT):
        if p not in sys.path:
            sys.path.insert(0, p)
    os.environ['MASTER_ADDR'] = '127.0.0.1'
    os.environ['MASTER_PORT'] = str(port)
    dist.init_process_group('gloo', rank=rank, world_size=world)
    try:
        import spline_table
        from FoKL import FoKLRoutines as FR
        from mock_engine import MockEngine
        from test_public_api_stand_in import StandInEngine
        from test_selection_mock import synthetic

        class ShardedStandIn(StandInEngine):
            def begin_fit(self, ds):
                MockEngine.__init__(self, ds.x, ds.y, self._phis_arg, self._kernel_arg, dist=dist)

        eng = ShardedStandIn()
        FR._engine = lambda device=None: eng
        phis = spline_table.to_phis(np.load(os.path.join(ROOT, 'tests', 'golden', 'phis_cubic_48.npy')))
        x, y = synthetic(333, 3, 12)
        raw = 5.0 * x - 2.0                               # un-normalised: the bounds must come from all ranks
        per = -(-len(y) // world)
        lo, hi = rank * per, min((rank + 1) * per, len(y))
        np.random.seed(70 + rank)
        model = FR.FoKL(phis=phis, way3=True, draws=40, burnin=40, UserWarnings=False, ConsoleOutput=False)
        betas, mtx, evs = model.fit(raw[lo:hi], y[lo:hi], clean=True)
        np.savez(os.path.join(out_dir, 'api_rank%d.npz' % rank), betas=betas, mtx=mtx, evs=evs, b=model.b, btau=model.btau,
                 minmax=np.asarray(model.minmax), rows=np.asarray(model.inputs).shape[0], n=eng.n_global)
    finally:
        dist.destroy_process_group()


@pytest.mark.timeout(300)
def test_two_rank_fit_through_the_public_api(tmp_path, phis_cubic, monkeypatch):
    """`FoKL.fit(raw_row_shard, data_shard, clean=True)` on two gloo ranks (stand-in engine): `clean` normalises every
    shard with the GLOBAL per-column bounds (MIN / MAX allreduce), the b / btau defaults come from the all-reduced data
    moments (FR:1322-1348 on the whole dataset), and both ranks return the single-rank fit of the whole dataset."""
    import torch.multiprocessing as mp
    from FoKL import FoKLRoutines as FR
    from test_public_api_stand_in import StandInEngine
    from test_selection_mock import synthetic
    world = 2
    port = 33500 + (os.getpid() % 2000)
    mp.spawn(_public_api_worker, args=(world, port, str(tmp_path)), nprocs=world, join=True)
    r = [np.load(os.path.join(str(tmp_path), 'api_rank%d.npz' % k)) for k in range(world)]
    x, y = synthetic(333, 3, 12)
    raw = 5.0 * x - 2.0
    eng = StandInEngine()
    monkeypatch.setattr(FR, '_engine', lambda device=None: eng)
    np.random.seed(70)
    model = FR.FoKL(phis=phis_cubic, way3=True, draws=40, burnin=40, UserWarnings=False, ConsoleOutput=False)
    betas, mtx, evs = model.fit(raw, y, clean=True)
    assert int(r[0]['rows']) + int(r[1]['rows']) == 333 and int(r[0]['n']) == int(r[1]['n']) == 333
    for k in range(world):
        assert np.array_equal(r[k]['minmax'], np.asarray(model.minmax))
        assert np.isclose(float(r[k]['b']), model.b, rtol=1e-12) and np.isclose(float(r[k]['btau']), model.btau, rtol=1e-12)
        assert np.array_equal(r[k]['mtx'], mtx)
        assert np.allclose(r[k]['evs'], evs, rtol=1e-9, atol=0)
        assert r[k]['betas'].shape == betas.shape
    for key in ('betas', 'mtx', 'evs', 'b', 'btau'):
        assert np.array_equal(r[0][key], r[1][key]), key
