"""Import the UNMODIFIED reference from /root/reference/src (build container only).

TEST INFRASTRUCTURE.  Used by oracle/gen_golden.py to produce tests/golden/*; never imported by the
product, by `-m gpu` tests, by smoke() or by bench.py (the reference does not exist on the GPU box).

Shims (no edits to reference source):
  * a matplotlib stub on sys.path (oracle/_stubs)
  * `phis=` injection of a regenerated cubic table (the upstream data file is a missing blob)
  * optional `itertools.permutations` replacement on the FoKLRoutines module that yields distinct
    permutations only -- np.unique(..., axis=0) at FoKLRoutines.py:1616 then returns identical `vecs`
    while avoiding the M! enumeration (FoKLRoutines.py:1350-1354) for M > 9.
"""
import os
import sys
import types

_HERE = os.path.dirname(os.path.abspath(__file__))
REF_SRC = os.environ.get('FOKL_REFERENCE_SRC', '/root/reference/src')


def load_reference(distinct_perms_shim=False):
    if not os.path.isdir(REF_SRC):
        raise RuntimeError('reference not available at %s' % REF_SRC)
    for p in (os.path.join(_HERE, '_stubs'), REF_SRC):
        if p not in sys.path:
            sys.path.insert(0, p)
    for name in list(sys.modules):
        if name == 'FoKL' or name.startswith('FoKL.'):
            mod = sys.modules[name]
            if not getattr(mod, '__file__', '').startswith(REF_SRC):
                del sys.modules[name]
    from FoKL import FoKLRoutines
    assert FoKLRoutines.__file__.startswith(REF_SRC)
    if distinct_perms_shim:
        import itertools as _it
        from fokl_oracle import distinct_perms

        shim = types.SimpleNamespace(**{k: getattr(_it, k) for k in dir(_it) if not k.startswith('_')})
        shim.permutations = lambda x: [tuple(r) for r in distinct_perms(x)]
        FoKLRoutines.itertools = shim
    return FoKLRoutines
