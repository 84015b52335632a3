"""Exploration script (not part of the product): per-call CUDA-event times of one resident cfg fit, in call order."""
import argparse
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, 'fokl-gpy_b200'))
sys.path.insert(0, ROOT)
import bench_data  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--cfg', default='cfg4')
    ap.add_argument('--n', type=int, default=0)
    ap.add_argument('--repeat', type=int, default=3)
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
    for rep in range(a.repeat):
        np.random.seed(cfg['seed'])
        eng.profile = {}
        model = bench_data.make_model(FR, a.cfg)
        model.fit(ds, None)
        torch.cuda.synchronize()
    ev = eng.profile.get('_events', [])
    t0 = ev[0][1]
    for name, s, e, extra in ev:
        ms = s.elapsed_time(e)
        line = '%9.3f  %-18s %8.3f ms' % (t0.elapsed_time(s), name, ms)
        if name == 'gram':
            line += '  cols=%d  %.1f TF/s  %.0f GB/s' % (extra['cols'], extra['flops'] / ms / 1e9, extra['bytes'] / ms / 1e6)
        elif name == 'basis':
            line += '  C=%d  %.0f GB/s' % (round(extra['cells'] / n), extra['bytes'] / ms / 1e6)
        else:
            line += '  ' + ' '.join('%s=%s' % kv for kv in extra.items())
        print(line)
    for k, v in eng.profile_summary().items():
        print(k, v)
    print(FR.LAST_FIT_INFO)


if __name__ == '__main__':
    main()
