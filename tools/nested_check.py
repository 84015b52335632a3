"""Diagnostic (not product): chains of nested models by secular-equation updates (csrc/nested.cu) against the cold path
(one eigensolver per model): the intercept's posterior mean of every model must agree; timings of both.
usage: python tools/nested_check.py p0 n_models [flavour]"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, 'fokl-gpy_b200'))
import torch  # noqa: E402
from FoKL import FoKLRoutines as FR, _lib  # noqa: E402

eng = FR._engine()
p0 = int(sys.argv[1]) if len(sys.argv) > 1 else 400
n_models = int(sys.argv[2]) if len(sys.argv) > 2 else 20
kind = sys.argv[3] if len(sys.argv) > 3 else 'gauss'
rng = np.random.default_rng(1)
n = 4 * p0 + 50
if kind == 'gauss':
    X = rng.standard_normal((n, p0)) * (1.0 + 3.0 * rng.random(p0))
else:
    u = rng.random((n, 16))
    X = np.empty((n, p0))
    for j in range(p0):
        k = rng.choice(16, size=3, replace=False)
        o = rng.integers(1, 4, size=3)
        X[:, j] = np.cos(np.pi * o[0] * u[:, k[0]]) * np.cos(np.pi * o[1] * u[:, k[1]]) * (
            np.cos(np.pi * o[2] * u[:, k[2]]) if j % 3 else 1.0) * (0.2 + rng.random()) + 0.02 * rng.standard_normal(n)
X[:, 0] = 1.0
y = 2.0 + X[:, 1:6] @ rng.standard_normal(5) + 0.3 * rng.standard_normal(n)
G = X.T @ X
cap = max(p0, 64)
eng.G = torch.zeros((cap, cap), dtype=torch.float64, device=eng.device)
eng.Xty = torch.zeros(cap, dtype=torch.float64, device=eng.device)
eng.G[:p0, :p0] = torch.from_numpy(G).to(eng.device)
eng.Xty[:p0] = torch.from_numpy(X.T @ y).to(eng.device)
eng.Gcap = cap
eng.n_global, eng.sum_y, eng.yty = n, float(y.sum()), float(y @ y)
hyp = eng.make_hypers(4, 1, 4, 1, 1, 1, 2000)
drop = rng.permutation(np.arange(1, p0))[:n_models]
sets = [np.array([c for c in range(p0) if c not in set(drop[:k + 1].tolist())], dtype=np.int32) for k in range(n_models)]
ids = np.arange(100, 100 + n_models, dtype=np.uint64)


def timed(fn):
    torch.cuda.synchronize()
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    r = fn()
    e.record()
    torch.cuda.synchronize()
    return r, s.elapsed_time(e)


nest = lambda: eng.nested_chains_launch(sets, hyp, 12345, ids).finish()      # noqa: E731
nest()
res, t_nest = timed(nest)
cold_fn = lambda: eng.evaluate(sets, hyp, rng_mode=_lib.RNG_PHILOX, seed=12345, stream_ids=ids, refine_tol=None)   # noqa: E731
cold, t_cold = timed(cold_fn)
st = cold.stats.cpu().numpy()
m_cold = np.array([st[3 * cold.vec_off[c] + 2 * cold.p[c]] for c in range(n_models)])
rel = np.abs(res['mean0'] - m_cold) / np.abs(m_cold)
print('%s p0=%d models=%d  nested %.1f ms (ok=%s status=%d)  cold %.1f ms   max rel diff of the intercept mean %.2e  (mean %.6f)'
      % (kind, p0, n_models, t_nest, res['ok'], res['status'], t_cold, rel.max(), m_cold[0]), flush=True)
print('   per model rel diff:', ' '.join('%.1e' % v for v in rel[:12]))
