"""Exploration script (not part of the product; build container only): random small whole fits on the LIVE unmodified
reference (/root/reference, subprocess) and, through this package's public API in parity mode, on the CPU stand-in engine
(tests/mock_engine.py).  Reports every case whose normalised inputs, b / btau, term matrix, BIC trace (1e-9) or numpy RNG
end state differ, with the width and the conditioning of the model at the first differing substage.  A difference that
starts where the model has p >= n columns, or where lambda_min / lambda_max of its Gram is below 1e-12, is the
reference's own unclamped 1 / Lamb on a singular Gram (FR:1502; SURVEY 8d: "rank-deficient rounds are reported, not
parity-checked"); anything earlier is a bug.

    python tools/diff_fuzz.py <seed> <n_cases>          # ~13 min for 60 cases on 8 cores

Results at the end of round 2 (profiles/r02_diff_fuzz.txt): 190 random fits, 175 identical, 15 part from the reference at
a singular Gram, 0 bugs.  The fixed list tests/diff/fit_cases.py (28 cases, all well posed) is what the test suite runs."""
import os
import pickle
import subprocess
import sys
import tempfile

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF_SRC = '/root/reference/src'


def random_cases(seed, count):
    rng = np.random.default_rng(seed)
    cases = {}
    for i in range(count):
        m, n = int(rng.integers(1, 6)), int(rng.integers(40, 220))
        ckw = {}
        if rng.random() < 0.4:
            ckw['cubic'] = True
        if m >= 3 and rng.random() < 0.5:
            ckw['way3'] = True
        if rng.random() < 0.3:
            ckw['aic'] = True
        if rng.random() < 0.2:
            ckw['gimmie'] = True
        if rng.random() < 0.6:
            ckw['tolerance'] = int(rng.integers(1, 6))
        if rng.random() < 0.4:
            ckw['threshav'] = float(rng.choice([0.0, 0.01, 0.05, 0.2, 0.6, 1.5]))
            ckw['threshstda'] = float(rng.choice([0.01, 0.1, 0.5, 1.0, 3.0]))
            ckw['threshstdb'] = float(rng.choice([0.5, 1.0, 2.0, 10.0, 1e9]))
        if rng.random() < 0.3:
            ckw['a'] = float(rng.choice([1.1, 2, 4, 20, 100]))
            ckw['atau'] = float(rng.choice([0.1, 1, 4, 30]))
        if rng.random() < 0.2:
            ckw['b'] = float(rng.choice([0.001, 0.1, 5.0]))
        if rng.random() < 0.2:
            ckw['btau'] = float(rng.choice([0.01, 10.0, 5000.0]))
        ckw['draws'], ckw['burnin'] = int(rng.integers(8, 60)), int(rng.integers(0, 40))
        fkw = {}
        if rng.random() < 0.15:
            fkw['train'] = float(rng.choice([0.4, 0.7, 0.9]))
        if rng.random() < 0.15:
            fkw['pillow'] = float(rng.choice([0.01, 0.2]))
        cases['fz%d_%03d' % (seed, i)] = (n, m, int(rng.integers(0, 10 ** 6)), ckw, fkw)
    return cases


def main():
    seed, count = int(sys.argv[1]), int(sys.argv[2])
    tmp = tempfile.mkdtemp()
    case_file, ref_file = os.path.join(tmp, 'cases.pkl'), os.path.join(tmp, 'ref.pkl')
    with open(case_file, 'wb') as f:
        pickle.dump(random_cases(seed, count), f)
    os.environ['FOKL_DIFF_CASES'] = case_file
    env = dict(os.environ, PYTHONPATH=os.pathsep.join([os.path.join(ROOT, 'oracle', '_stubs'), REF_SRC]))
    r = subprocess.run([sys.executable, os.path.join(ROOT, 'tests', 'diff', 'fit_cases.py'), ref_file], env=env,
                       capture_output=True, text=True)
    assert r.returncode == 0, r.stderr[-3000:]
    with open(ref_file, 'rb') as f:
        ref = pickle.load(f)['fits']
    for p in ('fokl-gpy_b200', 'oracle', 'tests', os.path.join('tests', 'diff')):
        sys.path.insert(0, os.path.join(ROOT, p))
    import fit_cases
    import FoKL._selection as sel
    from FoKL import FoKLRoutines as FR
    from test_public_api_stand_in import StandInEngine
    eng = StandInEngine()
    FR._engine = lambda device=None: eng
    FR.B200_CONFIG['rng'] = 'numpy'
    real_select, widths, rconds = sel.forward_select, [], []

    def select(*a, **k):
        def on_substage(ind, ev, terms):
            widths.append(terms.shape[0] + 1)
            w = np.linalg.eigvalsh(eng._G)                    # Gram of the model the substage ends with
            rconds.append(float(w[0] / w[-1]))
        k['on_substage'] = on_substage
        return real_select(*a, **k)
    sel.forward_select = select
    n_bad = n_sat = 0
    for name, cfg in fit_cases.CASES.items():
        del widths[:]
        del rconds[:]
        want, got = ref[name], fit_cases.run_case(FR, name)
        why = []
        if 'raised' in want or 'raised' in got:
            if want.get('raised') != got.get('raised'):
                why.append(('raised', want.get('raised'), got.get('raised')))
        else:
            why += [k for k in ('inputs', 'data', 'minmax', 'mtx') if not np.array_equal(want[k], got[k])]
            if want['b'] != got['b'] or want['btau'] != got['btau']:
                why.append('b/btau')
            if want['evs'].shape != got['evs'].shape or not np.allclose(want['evs'], got['evs'], rtol=1e-9, atol=0):
                why.append('evs')
            if want['digest'] != got['digest']:
                why.append('rng')
        if not why:
            continue
        k = 0
        if 'evs' in want and 'evs' in got:
            while k < min(len(got['evs']), len(want['evs'])) and \
                    abs(got['evs'][k] - want['evs'][k]) <= 1e-9 * abs(want['evs'][k]):
                k += 1
        n_rows = int(np.asarray(got.get('inputs', np.zeros((cfg[0], 1)))).shape[0])
        # singular for the reference's unclamped 1 / Lamb (FR:1502): p >= n, or lambda_min / lambda_max of a model on the way
        # there below 1e-12 (the substage's own candidates, which hold more columns, are worse conditioned still)
        saturated = k < len(widths) and (widths[k] >= n_rows or min(rconds[max(0, k - 1):k + 1]) < 1e-12)
        n_sat += saturated
        n_bad += not saturated
        print('%s %s\n   differs in %s; BIC traces agree for %d substages; model width at the first differing one: %s of n = %d'
              ' rows, lambda_min / lambda_max %.1e -> %s' % (name, cfg, why, k, widths[k] if k < len(widths) else '?', n_rows,
                               rconds[k] if k < len(rconds) else float('nan'),
                               'singular Gram (not parity-checkable)' if saturated else 'BUG'), flush=True)
    print('cases %d: identical %d, differ from a saturated substage on %d, BUGS %d'
          % (len(fit_cases.CASES), len(fit_cases.CASES) - n_bad - n_sat, n_sat, n_bad))


if __name__ == '__main__':
    main()
