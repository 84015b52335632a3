"""A host stand-in for the part of FoKL._engine.Engine that `eng_predict` and `eng_derivative_draws`
(FoKL/FoKLRoutines.py) use: tensors live on the CPU, and `lib` holds Python functions with the C ABI's argument lists
(include/fokl_b200.h) that read / write those tensors through their raw addresses, computing with the kernels' own
formulas compiled for the host (tests/host_emu).  TEST INFRASTRUCTURE: it lets the CPU suite run the product's ctypes
plumbing -- argument order, layouts (column-major X with leading dimension ld, term rows as int16, derivative rows as
uint8, divisors [M][3]), offsets -- against the reference's golden outputs."""
import ctypes
import types

import numpy as np
import torch

import emu
from FoKL import _lib
from FoKL._engine import BERNOULLI, CUBIC, pack_phis


def _view(ptr, shape, dtype):
    n = int(np.prod(shape))
    if n == 0:
        return np.zeros(shape, dtype=dtype)
    buf = (ctypes.c_char * (n * np.dtype(dtype).itemsize)).from_address(int(ptr))
    return np.frombuffer(buf, dtype=dtype).reshape(shape)


class FakeDeviceEngine:
    def __init__(self):
        self.torch, self.device, self.ctx = torch, torch.device('cpu'), None
        self.calls = []
        self.lib = types.SimpleNamespace(fokl_fill_ones=self._fill_ones, fokl_basis_build=self._basis_build,
                                         fokl_basis_build_deriv=self._basis_build_deriv,
                                         fokl_predict_draws=self._predict_draws)

    # ---- Engine interface --------------------------------------------------------------------------------------------
    def set_phis(self, phis, kernel):
        self.tab = np.ascontiguousarray(pack_phis(phis, kernel))
        self.kernel_id = _lib.KERNEL_CUBIC if kernel == CUBIC else _lib.KERNEL_BERNOULLI
        assert kernel in (CUBIC, BERNOULLI)

    def upload(self, inputs, data):
        inputs = np.ascontiguousarray(inputs, dtype=np.float64)
        n, m = inputs.shape
        ldx = ((max(n, 1) + 15) // 16) * 16
        x = torch.zeros((m, ldx), dtype=torch.float64)
        x[:, :n] = torch.from_numpy(inputs).t()
        return types.SimpleNamespace(x=x, y=torch.zeros(ldx, dtype=torch.float64), n=n, m=m, ldx=ldx)

    def _ck(self, rc):
        assert rc == 0, rc

    def synchronize(self):
        pass

    # ---- C ABI (include/fokl_b200.h) on host addresses ---------------------------------------------------------------
    def _fill_ones(self, ctx, dst, n):
        _view(dst, (n,), np.float64)[:] = 1.0
        return 0

    def _factors(self, x, col, orders, e, div, deriv_form):
        """deriv_form: the bss_derivatives kernel evaluates EVERY factor -- differentiated or not -- at the twice-normalised
        input of FR:584-586 (csrc/basis.cu), the plain kernel at the local coordinate xsm."""
        cubic = self.kernel_id == _lib.KERNEL_CUBIC
        if not deriv_form:
            out = (emu.basis_cubic if cubic else emu.basis_bernoulli)(x[:, col], orders, self.tab)
            return out[0] if isinstance(out, tuple) else out
        return emu.deriv_factors(x[:, col], orders, self.tab, e, div, cubic=cubic)

    def _columns(self, x_ptr, n, ldx, m, terms_ptr, deriv_ptr, div_ptr, c, out_ptr, ld):
        x = _view(x_ptr, (m, ldx), np.float64)[:, :n].T
        terms = _view(terms_ptr, (c, m), np.int16)
        deriv = _view(deriv_ptr, (c, m), np.uint8) if deriv_ptr else np.zeros((c, m), dtype=np.uint8)
        div = _view(div_ptr, (m, 3), np.float64) if div_ptr else np.ones((m, 3))
        out = _view(out_ptr, (c, ld), np.float64)
        for j in range(c):
            col = np.ones(n)
            for k in range(m):                                    # multiplied in increasing k, like the kernel
                if terms[j, k]:
                    e = int(deriv[j, k])
                    col = col * self._factors(x, k, [int(terms[j, k])], e, div[k, e], bool(deriv_ptr))[:, 0]
            out[j, :n] = col
        return 0

    def _basis_build(self, ctx, kernel, x, n, ldx, m, terms, c, xnew, ld):
        assert kernel == self.kernel_id
        self.calls.append(('basis_build', c))
        return self._columns(x, n, ldx, m, terms, None, None, c, xnew, ld)

    def _basis_build_deriv(self, ctx, kernel, x, n, ldx, m, terms, deriv, divisors, c, xnew, ld):
        assert kernel == self.kernel_id
        self.calls.append(('basis_build_deriv', c))
        return self._columns(x, n, ldx, m, terms, deriv, divisors, c, xnew, ld)

    def _predict_draws(self, ctx, X, ld, n, p, betas, draws, out):
        self.calls.append(('predict_draws', p, draws))
        Xv = _view(X, (p, ld), np.float64)[:, :n]
        b = _view(betas, (draws, p), np.float64)
        _view(out, (n, draws), np.float64)[:] = Xv.T @ b.T
        return 0
