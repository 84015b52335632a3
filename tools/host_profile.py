"""Exploration script (not part of the product): where the host spends its time during one resident cfg fit.
Prints (a) the device-idle gaps between consecutive profiled stages (CUDA events), (b) cProfile of one fit, top
entries by own time and by cumulative time."""
import argparse
import cProfile
import io
import os
import pstats
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
    ap.add_argument('--top', type=int, default=45)
    a = ap.parse_args()
    import torch
    from FoKL import FoKLRoutines as FR
    cfg = bench_data.CONFIGS[a.cfg]
    n = a.n or cfg['n']
    x, y = bench_data.make_rows(a.cfg, 0, n, n_total=n)
    model = bench_data.make_model(FR, a.cfg)
    eng = FR._engine()
    eng.set_phis(model.phis, cfg['kernel'])
    ds = eng.upload(x, y)

    def one(profile_events):
        np.random.seed(cfg['seed'])
        eng.profile = {} if profile_events else None
        m = bench_data.make_model(FR, a.cfg)
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        m.fit(ds, None)
        torch.cuda.synchronize()
        return (time.perf_counter() - t0) * 1e3

    for _ in range(3):
        one(False)
    print('plain fits (ms):', ['%.1f' % one(False) for _ in range(4)])
    print('fit with stage events (ms): %.1f' % one(True))
    ev = eng.profile.get('_events', [])
    ev = sorted(ev, key=lambda r: ev[0][1].elapsed_time(r[1]))
    t0 = ev[0][1]
    gaps = {}
    busy = 0.0
    prev_end, prev_name = None, None
    for name, s, e, extra in ev:
        if name == 'side_batch':
            continue
        if prev_end is not None:
            g = prev_end.elapsed_time(s)
            key = prev_name + ' -> ' + name
            k = gaps.setdefault(key, [0, 0.0])
            k[0] += 1
            k[1] += g
        busy += s.elapsed_time(e)
        prev_end, prev_name = e, name
    print('timeline (ms from the first event; side_batch runs on the side stream):')
    for name, s_, e_, extra in ev:
        print('%9.3f  %-18s %8.3f ms  %s' % (t0.elapsed_time(s_), name, s_.elapsed_time(e_),
                                            ' '.join('%s=%s' % kv for kv in extra.items() if kv[0] in ('cands', 'pmax', 'cols'))))
    print('main-stream stages busy %.1f ms; gaps between them:' % busy)
    for key, (cnt, tot) in sorted(gaps.items(), key=lambda kv: -kv[1][1]):
        print('   %-40s %3d x  %7.3f ms total  %6.3f avg' % (key, cnt, tot, tot / cnt))
    print('   total gap %.1f ms' % sum(v[1] for v in gaps.values()))
    eng.profile = None
    pr = cProfile.Profile()
    np.random.seed(cfg['seed'])
    m = bench_data.make_model(FR, a.cfg)
    torch.cuda.synchronize()
    pr.enable()
    m.fit(ds, None)
    torch.cuda.synchronize()
    pr.disable()
    for key in ('tottime', 'cumtime'):
        s = io.StringIO()
        pstats.Stats(pr, stream=s).strip_dirs().sort_stats(key).print_stats(a.top)
        print(s.getvalue()[:9000])
    print(FR.LAST_FIT_INFO)


if __name__ == '__main__':
    main()
