"""Diagnostic (not product): the blocked eigensolver (csrc/eigbig.cuh) against scipy.linalg.eigh, with timings.

usage: python tools/eig_big_check.py 705,1024,1700,2072 [batch]
  `batch` > 1 also evaluates that many nested models of each size side by side (throughput of a verification batch).
Design-matrix flavours: 'gauss' (scaled Gaussian columns) and 'spline' (products of smooth functions of uniform inputs:
graded, strongly correlated columns like the BSS-ANOVA design)."""
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, 'fokl-gpy_b200'))
import torch  # noqa: E402
from scipy.linalg import eigh  # noqa: E402
from FoKL import FoKLRoutines as FR, _lib  # noqa: E402

eng = FR._engine()
sizes = [int(s) for s in (sys.argv[1].split(',') if len(sys.argv) > 1 else '705,1024,1700'.split(','))]
batch = int(sys.argv[2]) if len(sys.argv) > 2 else 1
flavours = (sys.argv[3].split(',') if len(sys.argv) > 3 else ['gauss', 'spline'])
rng = np.random.default_rng(0)
pmax = max(sizes)


def design(kind, n, p):
    if kind == 'gauss':
        X = rng.standard_normal((n, p)) * (1.0 + 3.0 * rng.random(p))
    else:
        m = 16
        u = rng.random((n, m))
        X = np.empty((n, p))
        for j in range(p):
            k = rng.choice(m, size=3, replace=False)
            o = rng.integers(1, 4, size=3)
            X[:, j] = np.cos(np.pi * o[0] * u[:, k[0]]) * np.cos(np.pi * o[1] * u[:, k[1]]) * (
                np.cos(np.pi * o[2] * u[:, k[2]]) if j % 3 else 1.0) * (0.2 + rng.random()) + 0.02 * rng.standard_normal(n)
    X[:, 0] = 1.0
    return X


for kind in flavours:
    n = 4 * pmax + 50
    X = design(kind, n, pmax)
    y = X[:, :5] @ rng.standard_normal(5) + 0.1 * rng.standard_normal(n)
    G = X.T @ X
    Xty = X.T @ y
    cap = max(pmax, 64)
    eng.G = torch.zeros((cap, cap), dtype=torch.float64, device=eng.device)
    eng.Xty = torch.zeros(cap, dtype=torch.float64, device=eng.device)
    eng.G[:pmax, :pmax] = torch.from_numpy(G).to(eng.device)
    eng.Xty[:pmax] = torch.from_numpy(Xty).to(eng.device)
    eng.Gcap = cap
    eng.n_global, eng.sum_y, eng.yty = n, float(y.sum()), float(y @ y)
    hyp = eng.make_hypers(4, 1, 4, 1, 1, 1, 2000)
    for p in sizes:
        sets = [list(range(p))]
        eng.evaluate(sets, hyp, want_eig=True, refine_tol=None)
        torch.cuda.synchronize()
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record()
        res = eng.evaluate(sets, hyp, want_eig=True, refine_tol=None)
        e.record()
        torch.cuda.synchronize()
        ms = s.elapsed_time(e)
        lam = res.lamb[:p].cpu().numpy()
        Q = res.Q[:p * p].view(p, p).cpu().numpy().T
        t0 = time.time()
        lam_ref, _ = eigh(G[:p, :p])
        t_ref = time.time() - t0
        bh = res.betahat[:p].cpu().numpy()
        bh_ref = np.linalg.solve(G[:p, :p], Xty[:p])
        print('%-6s p=%4d  %8.2f ms  sweeps=%2d flags=%d  dlam/lmax=%.1e  rel-dlam(min)=%.1e  orth=%.1e  recon=%.1e  dbh=%.1e  '
              'cond=%.1e  (scipy eigh %.0f ms)' % (
                  kind, p, ms, res.info[0] >> 8, res.info[0] & 0xff, np.max(np.abs(lam - lam_ref)) / lam_ref[-1],
                  abs(lam[0] - lam_ref[0]) / lam_ref[0], np.max(np.abs(Q.T @ Q - np.eye(p))),
                  np.max(np.abs((Q * lam) @ Q.T - G[:p, :p])) / lam_ref[-1],
                  np.max(np.abs(bh - bh_ref)) / np.max(np.abs(bh_ref)), lam_ref[-1] / lam_ref[0], 1e3 * t_ref), flush=True)
        # eig + chain + betas + stats (what the selection loop asks for)
        s.record()
        res = eng.evaluate(sets, hyp, rng_mode=_lib.RNG_PHILOX, refine_tol=None)
        e.record()
        torch.cuda.synchronize()
        print('       p=%4d  eig+chain+betas %8.2f ms' % (p, s.elapsed_time(e)), flush=True)
        if batch > 1:
            drop = rng.permutation(np.arange(1, p))[:batch]
            sets = [[c for c in range(p) if c not in set(drop[:k + 1].tolist())] for k in range(batch)]
            eng.evaluate(sets, hyp, rng_mode=_lib.RNG_PHILOX, refine_tol=None)
            torch.cuda.synchronize()
            s.record()
            res = eng.evaluate(sets, hyp, rng_mode=_lib.RNG_PHILOX, refine_tol=None)
            e.record()
            torch.cuda.synchronize()
            print('       p=%4d  batch of %d nested models, eig+chain+betas %8.2f ms  sweeps %s' % (
                p, batch, s.elapsed_time(e), sorted(set((res.info >> 8).tolist()))), flush=True)
