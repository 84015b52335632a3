"""Exploration script (not part of the product): time fokl_gram_update (K2) alone on a random design matrix for the
(P_old, C) shapes of the cfg4 fit, under the tuning knobs of csrc/gram.cu (FOKL_GRAM_KERNEL / _KB / _STAGES)."""
import argparse
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, 'fokl-gpy_b200'))

SHAPES = [(1, 8), (8, 28), (12, 8), (14, 56), (20, 56), (21, 8), (28, 168), (37, 28), (37, 56), (49, 8), (51, 168),
          (58, 168), (64, 56)]


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--n', type=int, default=10_000_000)
    ap.add_argument('--reps', type=int, default=3)
    ap.add_argument('--variants', default='auto;FOKL_GRAM_KERNEL=cpasync;FOKL_GRAM_KERNEL=mb')
    ap.add_argument('--shapes', default='')
    a = ap.parse_args()
    import torch
    from FoKL import _engine
    eng = _engine.Engine()
    n = a.n
    ld = _engine._round_up(n, 16)
    shapes = [tuple(int(v) for v in s.split(',')) for s in a.shapes.split(';')] if a.shapes else SHAPES
    pmax = max(p + c for p, c in shapes)
    g = torch.Generator(device='cuda').manual_seed(1)
    full = torch.rand((pmax + 1, ld), dtype=torch.float64, device='cuda', generator=g) - 0.5
    X, y = full[1:], full[0]           # [y | X]: the engine's layout (tensor-map TMA kernel unless a knob says otherwise)
    block = torch.empty(((pmax + 1) * max(c for _, c in shapes) + 64,), dtype=torch.float64, device='cuda')
    variants = a.variants.split(';')
    knobs = ('FOKL_GRAM_KERNEL', 'FOKL_GRAM_KB', 'FOKL_GRAM_STAGES', 'FOKL_GRAM_PLACE', 'FOKL_GRAM_WARPS', 'FOKL_GRAM_KSPLIT')
    print('%-10s' % 'P_old,C', ' '.join('%-34s' % v[:34] for v in variants))
    for p_old, c in shapes:
        flops = 2.0 * n * (p_old * c + c * (c + 1) / 2 + c)
        cells = []
        ref = None
        for v in variants:
            for k in knobs:
                os.environ.pop(k, None)
            if v != 'auto':
                for kv in v.split(','):
                    k, val = kv.split('=')
                    os.environ[k] = val
            best = 1e9
            times = []
            err = 0.0
            try:
                for r in range(a.reps + 1):
                    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                    e0.record()
                    eng._ck(eng.lib.fokl_gram_update(eng.ctx, X.data_ptr(), ld, n, p_old, c, y.data_ptr(), block.data_ptr()))
                    e1.record()
                    torch.cuda.synchronize()
                    if r:
                        best = min(best, e0.elapsed_time(e1))
                        times.append(e0.elapsed_time(e1))
                out = block[:(p_old + c + 1) * c].view(p_old + c + 1, c).clone()
                if ref is None:
                    ref = out
                else:
                    # new x new part is only defined on and above the diagonal
                    m = torch.ones_like(out, dtype=torch.bool)
                    m[p_old:p_old + c] = torch.triu(torch.ones((c, c), dtype=torch.bool, device='cuda'))
                    err = float(((out - ref).abs() * m).max() / ref.abs().max())
                cells.append('%7.3f/%7.3f ms %5.1f TF/s %.0e' % (best, float(np.median(times)), flops / best / 1e9, err))
            except Exception as ex:  # noqa: BLE001
                cells.append('ERR ' + str(ex)[:28])
        print('%-10s' % ('%d,%d' % (p_old, c)), ' '.join('%-34s' % s for s in cells), flush=True)


if __name__ == '__main__':
    main()
