"""Differential test of the host-only API: ~40 calls (`clean` with every keyword and input container, train split,
`evaluate_basis`, constructor / `clear` / `save` / `load`, the error paths) run on the UNMODIFIED reference and on this
package, outcome by outcome -- returned values bit-identical, same attributes, same exception type, same warning
texts.  The reference lives only in the build container (/root/reference); elsewhere the test is skipped."""
import os
import pickle
import subprocess
import sys

import numpy as np
import pytest

from conftest import ROOT

REF_SRC = '/root/reference/src'
CASES = os.path.join(ROOT, 'tests', 'diff', 'host_api_cases.py')


def _run(pythonpath, out):
    env = dict(os.environ, PYTHONPATH=os.pathsep.join(pythonpath))
    r = subprocess.run([sys.executable, CASES, out], env=env, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stderr[-3000:]
    with open(out, 'rb') as f:
        return pickle.load(f)


def _same(a, b):
    if isinstance(a, (list, tuple)) and isinstance(b, (list, tuple)):
        return len(a) == len(b) and all(_same(u, v) for u, v in zip(a, b))
    if isinstance(a, dict) and isinstance(b, dict):
        return a.keys() == b.keys() and all(_same(a[k], b[k]) for k in a)
    if isinstance(a, float) and isinstance(b, float):
        return a == b or (np.isnan(a) and np.isnan(b))
    return type(a) is type(b) and a == b


@pytest.mark.skipif(not os.path.isdir(REF_SRC), reason='the reference is only present in the build container')
def test_host_api_outcomes_equal_the_reference(tmp_path):
    ref = _run([os.path.join(ROOT, 'oracle', '_stubs'), REF_SRC], str(tmp_path / 'ref.pkl'))
    mine = _run([os.path.join(ROOT, 'fokl-gpy_b200')], str(tmp_path / 'mine.pkl'))
    assert ref['file'].startswith(REF_SRC) and mine['file'].startswith(ROOT)
    assert ref['outcomes'].keys() == mine['outcomes'].keys() and len(ref['outcomes']) >= 40
    bad = []
    for name, want in ref['outcomes'].items():
        got = mine['outcomes'][name]
        if not _same(want['result'], got['result']):
            bad.append((name, 'result', want['result'], got['result']))
        elif want['warnings'] != got['warnings']:
            bad.append((name, 'warnings', want['warnings'], got['warnings']))
    assert not bad, '\n'.join('%s [%s]\n  reference: %.300r\n  here:      %.300r' % b for b in bad)
