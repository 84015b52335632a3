"""Synthetic workloads of BASELINE.json `configs`, generated in 8 fixed row blocks so the
data are identical however many ranks share them."""
import numpy as np

N_BLOCKS = 8

CONFIGS = {
    # cfg3: N = 1M, 4 inputs, 2-way, Bernoulli polynomials, 1000 + 1000 draws
    'cfg3': dict(n=1_000_000, m=4, seed=3, kernel='Bernoulli Polynomials', way3=False),
    # cfg4: N = 10M, 8 inputs, 3-way, cubic splines (the north_star target)
    'cfg4': dict(n=10_000_000, m=8, seed=4, kernel='Cubic Splines', way3=True),
    # cfg5: N = 1M, 16 inputs, 3-way
    'cfg5': dict(n=1_000_000, m=16, seed=5, kernel='Cubic Splines', way3=True),
}


def target(cfg, x, eps):
    if cfg == 'cfg3':
        return np.sin(2 * np.pi * x[:, 0]) + 2 * (x[:, 1] - 0.5) ** 2 + x[:, 2] * x[:, 3] + 0.05 * eps
    return (np.sin(2 * np.pi * x[:, 0]) + x[:, 1] * x[:, 2] + x[:, 3] * x[:, 4] * x[:, 5] + 0.5 * x[:, 6] ** 2
            + 0.05 * eps)


def make_rows(cfg, lo, hi, n_total=None):
    """Rows [lo, hi) of the synthetic dataset of `cfg` with n_total rows in all (default: the config's N)."""
    c = CONFIGS[cfg]
    n_total = n_total or c['n']
    if hi - lo <= 0:
        return np.zeros((0, c['m'])), np.zeros(0)
    per = -(-n_total // N_BLOCKS)
    xs, ys = [], []
    for b in range(N_BLOCKS):
        b_lo, b_hi = b * per, min((b + 1) * per, n_total)
        if b_hi <= lo or b_lo >= hi or b_hi <= b_lo:
            continue
        rng = np.random.default_rng([c['seed'], b, n_total])
        xb = rng.random((b_hi - b_lo, c['m']))
        eb = rng.standard_normal(b_hi - b_lo)
        s, e = max(lo, b_lo) - b_lo, min(hi, b_hi) - b_lo
        xs.append(xb[s:e])
        ys.append(target(cfg, xb[s:e], eb[s:e]))
    return np.concatenate(xs, axis=0), np.concatenate(ys, axis=0)


def make_model(FR, cfg, draws=1000, burnin=None, phis=None, **kw):
    c = CONFIGS[cfg]
    kwargs = dict(kernel=c['kernel'], way3=c['way3'], draws=draws, burnin=draws if burnin is None else burnin,
                  UserWarnings=False, ConsoleOutput=False)
    if phis is not None:
        kwargs['phis'] = phis
    kwargs.update(kw)
    return FR.FoKL(**kwargs)
