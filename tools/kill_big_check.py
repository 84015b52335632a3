"""Diagnostic (not product): the whole-device kill loop (csrc/killbig.cuh) against the single-CTA kernel -- identical
decisions and BICs bit for bit -- with timings.   usage: python tools/kill_big_check.py 150,300,900[,2072] [slow_max_p]"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, 'fokl-gpy_b200'))
import torch  # noqa: E402
from FoKL import FoKLRoutines as FR  # noqa: E402

eng = FR._engine()
sizes = [int(s) for s in (sys.argv[1].split(',') if len(sys.argv) > 1 else '150,300,900'.split(','))]
slow_max = int(sys.argv[2]) if len(sys.argv) > 2 else 1300
rng = np.random.default_rng(0)
pmax = max(sizes)
n = 4 * pmax + 50
X = rng.standard_normal((n, pmax)) * (1.0 + 3.0 * rng.random(pmax))
X[:, 0] = 1.0
y = X[:, :40] @ rng.standard_normal(40) + 0.5 * rng.standard_normal(n)
G = X.T @ X
cap = max(pmax, 64)
eng.G = torch.zeros((cap, cap), dtype=torch.float64, device=eng.device)
eng.Xty = torch.zeros(cap, dtype=torch.float64, device=eng.device)
eng.G[:pmax, :pmax] = torch.from_numpy(G).to(eng.device)
eng.Xty[:pmax] = torch.from_numpy(X.T @ y).to(eng.device)
eng.Gcap = cap
eng.n_global, eng.sum_y, eng.yty = n, float(y.sum()), float(y @ y)
hyp = eng.make_hypers(4, 1, 4, 1, 1, 1, 10)


def timed(fn):
    torch.cuda.synchronize()
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    r = fn()
    e.record()
    torch.cuda.synchronize()
    return r, s.elapsed_time(e)


for p in sizes:
    cols = list(range(p))
    vm = (3 * p) // 4
    cand = list(rng.permutation(np.arange(p - vm, p)))
    pos = [int(c) for c in cand]
    bv0 = np.sort(rng.random(vm))
    bv1 = rng.random(vm) * 3
    full = float(eng.evaluate([cols], hyp, refine_tol=None).ev[0])
    args = (cols, pos, bv0, bv1, hyp, 0.3, 0.5, 2.0, 2.0, full, 0.0, 0)
    os.environ['FOKL_KILL_BIG_MIN_P'] = '2'
    eng.kill_loop(*args)
    big, t_big = timed(lambda: eng.kill_loop(*args))
    line = 'p=%4d vm=%4d  big: %8.2f ms  n_acc=%d tested=%d bad=%d' % (p, vm, t_big, big['n_acc'], big['tested'], big['bad'])
    if p <= slow_max:
        os.environ['FOKL_KILL_BIG_MIN_P'] = '0'
        one, t_one = timed(lambda: eng.kill_loop(*args))
        same = (one['n_acc'] == big['n_acc'] and one['tested'] == big['tested'] and one['bad'] == big['bad'] and
                np.array_equal(one['acc'], big['acc']) and np.array_equal(one['calls'], big['calls']) and
                np.array_equal(one['ev'], big['ev']))
        line += '   single CTA: %9.2f ms  identical=%s' % (t_one, same)
    print(line, flush=True)
os.environ.pop('FOKL_KILL_BIG_MIN_P', None)
