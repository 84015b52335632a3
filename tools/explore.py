"""Exploration script (not part of the product): run one fit of a synthetic config with per-stage timing."""
import argparse
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, 'fokl-gpy_b200'))
sys.path.insert(0, ROOT)
import bench_data  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--cfg', default='cfg4')
    ap.add_argument('--n', type=int, default=0)
    ap.add_argument('--eager', type=int, default=0)
    ap.add_argument('--draws', type=int, default=1000)
    ap.add_argument('--cprofile', type=int, default=0)
    ap.add_argument('--repeat', type=int, default=1)
    ap.add_argument('--resident', type=int, default=0, help='fit a DeviceDataset already in HBM (bench.py `value` path)')
    a = ap.parse_args()
    import torch
    from FoKL import FoKLRoutines as FR
    cfg = bench_data.CONFIGS[a.cfg]
    n = a.n or cfg['n']
    t0 = time.time()
    x, y = bench_data.make_rows(a.cfg, 0, n, n_total=n)
    print('data', x.shape, time.time() - t0, flush=True)
    model = bench_data.make_model(FR, a.cfg, draws=a.draws)
    FR.B200_CONFIG['eager_chains'] = bool(a.eager)
    eng = FR._engine()
    eng.profile = {}
    if a.resident:
        eng.set_phis(model.phis, cfg['kernel'])
        ds = eng.upload(x, y)
        fit_args = (ds, None)
        fit_kw = {}
    else:
        fit_args = (x, y)
        fit_kw = dict(clean=True)
    np.random.seed(cfg['seed'])
    subs = []
    from FoKL import _selection
    orig = _selection.forward_select

    def patched(*args, **kw):
        def _on(ind, ev, terms):
            subs.append((ind, ev, terms.shape[0], time.time() - t1))
            print('substage', subs[-1], flush=True)
        kw['on_substage'] = _on
        return orig(*args, **kw)
    _selection.forward_select = patched
    t1 = time.time()
    for rep in range(a.repeat - 1):          # warm-up fits (allocator, module loading)
        np.random.seed(cfg['seed'])
        bench_data.make_model(FR, a.cfg, draws=a.draws).fit(*fit_args, **fit_kw)
        subs.clear()
        eng.profile = {}
    np.random.seed(cfg['seed'])
    t1 = time.time()
    if a.cprofile:
        import cProfile
        import pstats
        pr = cProfile.Profile()
        pr.enable()
    betas, mtx, evs = model.fit(*fit_args, **fit_kw)
    torch.cuda.synchronize()
    if a.cprofile:
        pr.disable()
        pstats.Stats(pr).sort_stats('cumulative').print_stats(35)
    print('fit wall', time.time() - t1, FR.LAST_FIT_INFO, flush=True)
    for s in subs:
        print(s)
    for k, v in eng.profile_summary().items():
        print(k, v)
    print('terms', mtx.shape, 'evs', len(evs))


if __name__ == '__main__':
    main()
