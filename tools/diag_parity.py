"""Diagnose a parity divergence: device fit (parity mode) vs oracle replay on the device's Gram bits, call by call."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (os.path.join(ROOT, 'fokl-gpy_b200'), os.path.join(ROOT, 'oracle'), os.path.join(ROOT, 'tests')):
    sys.path.insert(0, p)
import fokl_oracle as fo  # noqa: E402
from conftest import load_golden  # noqa: E402
from FoKL import FoKLRoutines as FR, _lib  # noqa: E402
import test_gpu_fit as T  # noqa: E402

name = sys.argv[1] if len(sys.argv) > 1 else 'isotherm_gp'
g = load_golden(name)
bern = np.load(os.path.join(ROOT, 'tests', 'golden', 'bernoulli_table.npy'))
phis_bern = tuple(list(bern[n, :n + 2]) for n in range(bern.shape[0]))
import spline_table  # noqa: E402
phis_cubic = spline_table.to_phis(np.load(os.path.join(ROOT, 'tests', 'golden', 'phis_cubic_48.npy')))
phis = phis_cubic if str(g['kernel']) == fo.CUBIC else phis_bern

eng = FR._engine()
orig = eng.evaluate
dev_log = []


def wrap(col_sets, hyp, **kw):
    r = orig(col_sets, hyp, **kw)
    if kw.get('rng_mode', _lib.RNG_NONE) != _lib.RNG_NONE:
        dev_log.append((len(col_sets[0]), float(r.ev[0]), int(r.info[0])))
    return r


eng.evaluate = wrap
grams = {}
keys = []


def rec(k, G, xty):
    grams[k] = (G, xty)
    keys.append(k)


model, betas, mtx, evs, info, dig = T.fit_device(FR, g, phis, recorder=rec)
print('device: calls', len(dev_log), 'substages', len(evs), 'terms', mtx.shape[0])
or_log = []


def on_gibbs(d):
    or_log.append((d['discmtx'].shape[0] + 1, float(d['ev'])))


def hook(discmtx):
    key = tuple(map(tuple, np.asarray(discmtx, dtype=np.int64)))
    if key not in grams:
        raise KeyError('missing')
    return grams[key]


np.random.seed(int(g['seed']))
try:
    r = fo.fit(g['inputs'], g['data'], phis, kernel=str(g['kernel']), a=float(g['a']), b=float(g['b']),
               atau=float(g['atau']), btau=float(g['btau']), tolerance=int(g['tolerance']), burnin=int(g['burnin']),
               draws=int(g['draws']), way3=bool(g['way3']), aic=bool(g['aic']), gram_hook=hook, on_gibbs=on_gibbs)
    print('oracle finished', len(or_log))
except KeyError:
    print('oracle diverged at call', len(or_log) + 1)
for i in range(max(0, len(or_log) - 12), min(len(dev_log), len(or_log) + 3)):
    d = dev_log[i]
    o = or_log[i] if i < len(or_log) else None
    rel = abs(d[1] - o[1]) / abs(o[1]) if o else float('nan')
    print(i + 1, 'dev p=%d ev=%.12g info=%d' % d, 'oracle', o, 'rel %.2e' % rel)
worst = max((abs(d[1] - o[1]) / abs(o[1]), i) for i, (d, o) in enumerate(zip(dev_log, or_log)))
print('worst rel diff before divergence', worst)
# conditioning of the Gram at the divergence point
k = keys[min(len(or_log), len(keys) - 1)]
G = grams[k][0]
w = np.linalg.eigvalsh(G)
print('gram at divergence: p', G.shape[0], 'eig min %.3e max %.3e' % (w[0], w[-1]))
