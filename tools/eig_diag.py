"""Diagnostic (not product): sweeps and time of the eigensolver / chain / kill loop at several model sizes."""
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, 'fokl-gpy_b200'))
import torch  # noqa: E402
from FoKL import FoKLRoutines as FR, _lib  # noqa: E402

eng = FR._engine()
sizes = [int(s) for s in (sys.argv[1].split(',') if len(sys.argv) > 1 else '8,30,60,100,128,200,230,300,440'.split(','))]
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
for p in sizes:
    for mode, name in ((_lib.RNG_NONE, 'eig'), (_lib.RNG_PHILOX, 'eig+chain')):
        eng.evaluate([list(range(p))], hyp, rng_mode=mode, refine_tol=None)
        torch.cuda.synchronize()
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record()
        res = eng.evaluate([list(range(p))], hyp, rng_mode=mode, refine_tol=None)
        e.record()
        torch.cuda.synchronize()
        print('p=%4d %-10s %.3f ms  sweeps=%d flags=%d' % (p, name, s.elapsed_time(e), res.info[0] >> 8, res.info[0] & 7), flush=True)
