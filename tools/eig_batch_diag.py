"""Diagnostic (not product): time of fokl_candidates_eval (eig + BIC only) over model width p, forced cluster size
(FOKL_EIGJ_CS) and batch size -- the data behind the cluster-size heuristic in candidates.cu."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, 'fokl-gpy_b200'))
import torch  # noqa: E402
from FoKL import FoKLRoutines as FR, _lib  # noqa: E402

eng = FR._engine()
sizes = [60, 100, 140, 180, 220]
batches = [1, 16, 64, 160]
rng = np.random.default_rng(0)
pmax = max(sizes)
n = 4 * pmax + 50
X = rng.random((n, pmax)) - 0.5
X[:, 0] = 1.0
y = X[:, :5] @ rng.standard_normal(5) + 0.1 * rng.standard_normal(n)
G = X.T @ X
cap = max(pmax, 64)
eng.G = torch.zeros((cap, cap), dtype=torch.float64, device=eng.device)
eng.Xty = torch.zeros(cap, dtype=torch.float64, device=eng.device)
eng.G[:pmax, :pmax] = torch.from_numpy(G).to(eng.device)
eng.Xty[:pmax] = torch.from_numpy(X.T @ y).to(eng.device)
eng.Gcap = cap
eng.n_global, eng.sum_y, eng.yty = n, float(y.sum()), float(y @ y)
hyp = eng.make_hypers(4, 1, 4, 1, 1, 1, 2000)
print('p batch ' + ' '.join('cs=%-2d' % c for c in (1, 2, 4, 8, 16)))
for p in sizes:
    for b in batches:
        row = []
        for cs in (1, 2, 4, 8, 16):
            os.environ['FOKL_EIGJ_CS'] = str(cs)
            # batch of b models of width ~p (distinct column subsets so nothing is cached between candidates)
            sets = [[0] + sorted(rng.choice(np.arange(1, pmax), p - 1, replace=False).tolist()) for _ in range(b)]
            try:
                eng.evaluate(sets, hyp, rng_mode=_lib.RNG_NONE, refine_tol=None)
                torch.cuda.synchronize()
                s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                s.record()
                res = eng.evaluate(sets, hyp, rng_mode=_lib.RNG_NONE, refine_tol=None)
                e.record()
                torch.cuda.synchronize()
                row.append('%6.2f' % s.elapsed_time(e))
            except Exception as ex:  # noqa: BLE001
                row.append('  fail')
        print('%3d %4d  %s' % (p, b, ' '.join(row)), flush=True)
