"""The C-ABI library loads and exports every symbol include/fokl_b200.h declares (no compute calls: no GPU here)."""
import os
import re

from conftest import ROOT


def declared_symbols():
    text = open(os.path.join(ROOT, 'include', 'fokl_b200.h')).read()
    text = re.sub(r'/\*.*?\*/', '', text, flags=re.S)
    return sorted(set(re.findall(r'\b(fokl_[a-z_0-9]+)\s*\(', text)))


def test_header_declares_the_documented_entry_points():
    syms = declared_symbols()
    for must in ('fokl_ctx_create', 'fokl_basis_build', 'fokl_gram_update', 'fokl_candidates_eval', 'fokl_last_error'):
        assert must in syms


def test_library_exports_every_declared_symbol_and_binding_matches():
    from FoKL import _lib
    _lib.build_library()
    lib = _lib.load()
    syms = declared_symbols()
    for s in syms:
        assert hasattr(lib, s), s
    assert sorted(_lib.PROTOTYPES) == syms
    assert lib.fokl_abi_version() == _lib.ABI_VERSION


def test_hypers_struct_layout_matches_header():
    import ctypes
    from FoKL import _lib
    assert ctypes.sizeof(_lib.Hypers) == 8 * 8 + 8 + 4 * 4
    assert _lib.Hypers.n.offset == 64 and _lib.Hypers.draws.offset == 72


def test_product_code_never_imports_the_oracle():
    pkg = os.path.join(ROOT, 'fokl-gpy_b200')
    for base, _, files in os.walk(pkg):
        for f in files:
            if f.endswith(('.py', '.cu', '.cuh')):
                src = open(os.path.join(base, f)).read()
                assert 'fokl_oracle' not in src and 'import emu' not in src, f


def test_every_call_site_passes_the_number_of_arguments_the_prototype_declares():
    """Static check of the ctypes call sites in FoKL/_engine.py, FoKL/FoKLRoutines.py and FoKL/_update.py: each
    `...lib.fokl_xxx(a, b, ...)` passes exactly as many positional arguments as `_lib.PROTOTYPES['fokl_xxx']` (=
    include/fokl_b200.h) declares -- ctypes would only say so at run time, on a GPU box."""
    import ast
    import inspect
    from FoKL import FoKLRoutines, _engine, _lib, _update
    seen = set()
    for mod in (_engine, FoKLRoutines, _update):
        for node in ast.walk(ast.parse(inspect.getsource(mod))):
            if isinstance(node, ast.Call) and isinstance(node.func, ast.Attribute) and node.func.attr.startswith('fokl_') \
                    and isinstance(node.func.value, ast.Attribute) and node.func.value.attr == 'lib':
                name = node.func.attr
                assert name in _lib.PROTOTYPES, name
                assert not node.keywords and not any(isinstance(a, ast.Starred) for a in node.args), name
                assert len(node.args) == len(_lib.PROTOTYPES[name][1]), (mod.__name__, name, node.lineno, len(node.args),
                                                                        len(_lib.PROTOTYPES[name][1]))
                seen.add(name)
    assert len(seen) >= 20, sorted(seen)
