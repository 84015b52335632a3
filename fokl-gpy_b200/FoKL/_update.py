"""`fitupdate` (update=True): host side of the reference's update fits, FR:1850-2583.

Reference flow (src/FoKL/FoKLRoutines.py): `fit` hands over to `fitupdate` when `self.update` is set (FR:1365-1367).
`modelBuilder` (FR:1939-1948) turns the previous fit's draws into a Gaussian prior (mu_old, Sigma_old) once the model
is `built`; the term loop (FR:2481-2575) grows the interaction matrix two-way only and WITHOUT a kill loop, and calls
the three-case sampler `gibbs_Xin_update` (FR:1958-2430) once per stage whose matrix has at least as many terms as the
prior; the evidence of a stage is the maximum over the draws of the log-likelihood (FR:2143, 2257, 2419).

Here: K1 + K2 (Engine.append_terms) build the new columns and their Gram block exactly as for `fit`; every candidate
model of a stage is the leading principal sub-matrix of that Gram.  Per stage the spectral preparation runs on the
device -- eigendecompositions with the library's own solver (fokl_candidates_eval, no chain), projections as plain FP64
GEMMs -- and `fokl_update_chain` (csrc/update.cu) runs the draw loop:
  case 1  the eigenbasis chain of `fit` + the per-draw likelihood;
  case 2  ONE generalised eigendecomposition of (X'X, Sigma_old^-1) instead of the reference's eigh + inv per draw
          (FR:2197-2203): same conditional law at every draw, O(p) per draw instead of O(p^3);
  case 3  block Gibbs in the coordinates of the two fixed eigendecompositions (FR:2296, 2312).
The prior itself (mean / covariance of the previous draws, FR:1942-1943, and `np.linalg.inv(Sigma_old)`, FR:2171 / 2291)
is p x p host bookkeeping on host-resident draws and is computed with the same numpy calls as the reference.

Parity mode (`B200_CONFIG['rng'] = 'numpy'`): the variates of the global legacy numpy RNG are injected in the
reference's order and the eigenvector signs are aligned with scipy.linalg.eigh of the same matrix bits (SURVEY 0.7;
only +-1 factors come from the host).  Cases 1 and 3 then reproduce the reference's draws; case 2 reproduces its law
(see DESIGN.md section 3d).
"""
import ctypes
import math

import numpy as np

from . import _lib
from ._selection import PhiloxVariates, distinct_permutations


def model_prior(betas, burn):
    """FR:1939-1948 for a built model: (mu_old 1 x p matrix, sigma_old p x p)."""
    mu_old = np.asmatrix(np.mean(betas[burn:-1], axis=0))
    sigma_old = np.cov(betas[burn:-1].transpose())
    return mu_old, sigma_old


def _i_list(ind):
    """FR:2484-2489."""
    if ind == 1:
        return [0]
    return list(np.arange(0, math.floor(ind / 2) + 0.1, 1)[::-1])


def _numpy_variates(draws, po, pn, astar, atau_star):
    """The global legacy numpy RNG, consumed as one gibbs_Xin_update call does (FR:2106 / 2213 / 2355, 2367, then the
    two gammas; np.random.gamma(k, s) == s * standard_gamma(k) bitwise)."""
    w = po + pn
    out = np.empty((draws, w + 2))
    normal, sgamma = np.random.normal, np.random.standard_gamma
    for k in range(draws):
        if po:
            out[k, :po] = normal(loc=0, scale=1, size=(po, 1))[:, 0]
        if pn:
            out[k, po:w] = normal(loc=0, scale=1, size=(pn, 1))[:, 0]
        out[k, w] = sgamma(astar)
        out[k, w + 1] = sgamma(atau_star)
    return out


class _Sampler:
    """One gibbs_Xin_update call on the engine's current Gram (first p columns)."""

    def __init__(self, engine, hy, prior, rng, seed):
        self.e, self.hy, self.rng, self.seed = engine, hy, rng, seed
        self.torch = engine.torch
        self.calls = 0
        self.prior = None
        if prior is not None:
            mu_old, sigma_old = prior
            mu = np.asarray(mu_old, dtype=np.float64).reshape(-1)
            sinv = np.linalg.inv(np.asarray(sigma_old, dtype=np.float64))           # FR:2171, 2291
            if not np.all(np.isfinite(sinv)):
                raise ValueError("array must not contain infs or NaNs")             # what upstream's eigh says (FR:2296)
            self.prior = dict(p=len(mu), mu=self._dev(mu), sinv=self._dev(sinv), sinv_host=sinv)

    def _dev(self, a):
        return self.torch.from_numpy(np.ascontiguousarray(a, dtype=np.float64)).to(self.e.device)

    def _eigh(self, A):
        """(lam ascending, Q with eigenvectors as columns) of a symmetric positive definite device matrix; in parity
        mode the signs of the columns are those of scipy.linalg.eigh on the same bits, else every eigenvector is
        oriented by its largest component (a convention that does not depend on the solver)."""
        A = self.torch.tril(A) + self.torch.tril(A, -1).t()       # scipy's eigh reads the lower triangle (FR:2296)
        lam, Q = self.e.sym_eigh(A)
        if self.rng == 'numpy':
            from scipy.linalg import eigh as _eigh
            with np.errstate(all='ignore'):
                _, q_ref = _eigh(A.cpu().numpy())
            sg = np.sign(np.sum(Q.cpu().numpy() * q_ref, axis=0))
            sg[sg == 0] = 1.0
            Q = Q * self._dev(sg)[None, :]
        else:
            idx = Q.abs().argmax(dim=0)
            sg = self.torch.sign(Q.gather(0, idx[None, :]))[0]
            sg = self.torch.where(sg == 0, self.torch.ones_like(sg), sg)
            Q = Q * sg[None, :]
        return lam, Q

    def run(self, p):
        """-> (betas numpy draws x p, ev float).  p = mmtx + 1 columns of the current X."""
        e, hy, torch = self.e, self.hy, self.torch
        self.calls += 1
        D = int(hy['total_draws'])
        n = e.n_global
        G = e.G[:p, :p]
        Xty = e.Xty[:p]
        spec = dict(draws=D, b=hy['b'], btau=hy['btau'], sigsqd0=hy['sigsqd0'], yty=e.yty, squerr=0.0, n=n)
        if self.prior is None:
            case, po, pn = 1, 0, p
        elif self.prior['p'] == p:
            case, po, pn = 2, p, 0
        elif self.prior['p'] < p:
            case, po, pn = 3, self.prior['p'], p - self.prior['p']
        else:
            raise RuntimeError('Error: No appropriate cases for evaluation found.')      # FR:2428-2430
        mmtx = p - 1
        arrays = {}
        if case == 1:
            lam, Q = self._eigh(G)
            ct = Q.t().mv(Xty)
            betahat = Q.mv(ct / lam)                                                     # FR:2080
            spec.update(mode=1, po=0, pn=p, a_star=hy['a'] + 1 + n / 2 + (mmtx + 1) / 2, atau_star=hy['atau'] + mmtx / 2,
                        squerr=e.residual_sse(p, betahat))                               # FR:2083-2087
            arrays = dict(lam_n=lam, c_n=ct)
            back = lambda go, gn: gn.mm(Q.t())                                           # noqa: E731
        elif case == 2:
            # Sigma_old^-1 = L L',  L^-1 G L^-T = V D V',  T = L^-T V:  T'G T = D,  T' Sigma_old^-1 T = I
            sinv = self.prior['sinv']
            try:
                L = torch.linalg.cholesky(0.5 * (sinv + sinv.t()))
            except torch.linalg.LinAlgError as exc:
                # a prior covariance estimated from fewer draws than coefficients is singular: upstream's inv(Sigma_old)
                # (FR:2171) is then inf / nan / indefinite and its `eigh` refuses it with this ValueError
                raise ValueError("array must not contain infs or NaNs (the covariance of the previous draws is "
                                 "singular: keep more draws than the model has coefficients)") from exc
            Li = torch.linalg.solve_triangular(L, torch.eye(p, dtype=torch.float64, device=e.device), upper=False)
            C = Li.mm(G).mm(Li.t())
            C = 0.5 * (C + C.t())
            dg, V = self._eigh(C)
            T = Li.t().mm(V)
            spec.update(mode=2, po=p, pn=0, a_star=hy['a'] + n / 2 + (mmtx + 1) / 2,
                        atau_star=hy['atau'] + (mmtx + 1) / 2)                           # FR:2176-2177
            arrays = dict(lam_o=dg, c_o=T.t().mv(Xty), m_o=V.t().mv(L.t().mv(self.prior['mu'])))
            back = lambda go, gn: go.mm(T.t())                                           # noqa: E731
        else:
            sinv, mu = self.prior['sinv'], self.prior['mu']
            Goo, Gon, Gnn = G[:po, :po], G[:po, po:], G[po:, po:]
            lam_o, Qo = self._eigh(Goo + sinv)                                           # FR:2296
            lam_n, Qn = self._eigh(Gnn.contiguous())                                     # FR:2312
            h = sinv.mv(mu)
            M = Qo.t().mm(Gon).mm(Qn)
            spec.update(mode=3, po=po, pn=pn, a_star=hy['a'] + n / 2 + (mmtx + 1) / 2, atau_star=hy['atau'] + pn / 2)
            arrays = dict(lam_o=lam_o, c_o=Qo.t().mv(Xty[:po] + h), t_o=Qo.t().mv(Xty[:po]), m_o=Qo.t().mv(mu),
                          lam_n=lam_n, c_n=Qn.t().mv(Xty[po:]), M=M.contiguous(), Mt=M.t().contiguous(),
                          K=Qo.t().mm(Goo).mm(Qo).contiguous(), W=Qo.t().mm(sinv).mm(Qo).contiguous())
            back = lambda go, gn: torch.cat([go.mm(Qo.t()), gn.mm(Qn.t())], dim=1)       # noqa: E731
        if self.rng == 'numpy':
            out = e.update_chain(spec, arrays, _lib.RNG_INJECTED,
                                 variates=_numpy_variates(D, po, pn, spec['a_star'], spec['atau_star']))
        else:
            out = e.update_chain(spec, arrays, _lib.RNG_PHILOX, seed=self.seed, stream_id=self.calls)
        betas = back(out['gam_o'], out['gam_n']).cpu().numpy()
        lik = out['lik'].cpu().numpy()
        with np.errstate(all='ignore'):
            ev = (mmtx + 1) * np.log(n) - 2 * np.max(lik)
        self.last = dict(case=case, lik=lik, sigs=out['sigs'], taus=out['taus'], bad=out['bad'])
        return case, betas, ev


def update_select(engine, hy, m, n_phis, prior=None, console=False, rng='philox', on_call=None):
    """The term loop of FR:2481-2575 on an Engine with a dataset bound (engine.begin_fit).

    hy: dict with a, b, atau, btau, tolerance, total_draws, gimmie, aic, sigsqd0.  prior: None (model not built) or
    (mu_old, sigma_old) from model_prior.  Returns dict(betas, mtx, evs, built, n_gibbs) with the reference's types:
    case 1 -> betas ndarray, evs 1-D; cases 2 / 3 -> betas np.matrix, evs (k, 1) (a list of one array if k == 1)."""
    if m < 2:
        # FR:2529 unpacks np.shape(damtx) of a scalar for a single input
        raise ValueError("not enough values to unpack (expected 2, got 0)")
    n = engine.n_global
    seed = 0 if rng == 'numpy' else PhiloxVariates(engine).seed
    smp = _Sampler(engine, hy, prior, rng, seed)
    num_old_terms = 0 if prior is None else int(np.shape(prior[0])[1])
    damtx = np.zeros((0, m))
    evs = []
    ind = 1
    greater = 0
    finished = 0
    built = False
    betas = betas_best = mtx = None
    while True:
        for i in _i_list(ind):
            vecs = np.zeros(m)
            vecs[0] = ind - i
            vecs[1] = i
            vecs = distinct_permutations(vecs).astype(np.float64)                        # FR:2496-2500
            damtx = np.concatenate((damtx, vecs), axis=0)
            length = damtx.shape[0]
            if num_old_terms - 1 <= length:                                              # FR:2531
                have = engine.P - 1
                if length > have:
                    engine.append_terms(damtx[have:length].astype(np.int64))             # K1 + K2 (FR:2008-2055)
                case, betas, ev = smp.run(length + 1)
                if case != 1:
                    betas = np.asmatrix(betas)
                    ev = np.array([ev])
                if hy['aic']:
                    ev = ev + (2 - np.log(n)) * length                                   # FR:2538-2544
                if console:
                    print(ind, ev)
                if on_call is not None:
                    on_call(dict(call=smp.calls, discmtx=damtx.copy(), betas=betas, ev=ev, **smp.last))
                if np.size(evs) == 0:
                    evs = [ev]
                else:
                    evs = np.concatenate((evs, [ev]))
                if ev == np.min(evs):                                                    # FR:2556-2566
                    betas_best = betas
                    mtx = damtx
                    greater = 1
                elif greater <= hy['tolerance']:
                    greater = greater + 1
                else:
                    finished = 1
                    built = True
                    break
        if finished != 0:
            break
        ind = ind + 1
        if ind > n_phis:
            break
    if hy['gimmie']:
        betas_best = betas
        mtx = damtx
    return dict(betas=betas_best, mtx=mtx, evs=evs, built=built, n_gibbs=smp.calls)


def engine_update_chain(engine, spec, arrays, rng_mode, seed=0, stream_id=0, variates=None):
    """Engine.update_chain: bind the tensors and call fokl_update_chain."""
    torch = engine.torch
    D, po, pn = int(spec['draws']), int(spec['po']), int(spec['pn'])
    mdl = _lib.UpdateModel(int(spec['mode']), po, pn, D, float(spec['a_star']), float(spec['atau_star']), float(spec['b']),
                           float(spec['btau']), float(spec['sigsqd0']), float(spec['yty']), float(spec['squerr']),
                           int(spec['n']))
    f64 = dict(dtype=torch.float64, device=engine.device)
    keep = {k: v.contiguous() for k, v in arrays.items()}

    def ptr(k):
        t = keep.get(k)
        return None if t is None else t.data_ptr()
    gam_o = torch.zeros((D, po), **f64)
    gam_n = torch.zeros((D, pn), **f64)
    sigs, taus, lik = torch.empty(D, **f64), torch.empty(D, **f64), torch.empty(D, **f64)
    info = torch.zeros(1, dtype=torch.int32, device=engine.device)
    var_t = None
    if rng_mode == _lib.RNG_INJECTED:
        var_t = torch.from_numpy(np.ascontiguousarray(variates, dtype=np.float64)).to(engine.device)
    t0 = engine._tic()
    engine._ck(engine.lib.fokl_update_chain(
        engine.ctx, ctypes.byref(mdl), ptr('lam_o'), ptr('c_o'), ptr('t_o'), ptr('m_o'), ptr('lam_n'), ptr('c_n'),
        ptr('M'), ptr('Mt'), ptr('K'), ptr('W'), rng_mode, ctypes.c_uint64(int(seed)), ctypes.c_uint64(int(stream_id)),
        None if var_t is None else var_t.data_ptr(), gam_o.data_ptr() if po else None,
        gam_n.data_ptr() if pn else None, sigs.data_ptr(), taus.data_ptr(), lik.data_ptr(), info.data_ptr()))
    engine._toc(t0, 'update_chain', cands=1, pmax=po + pn)
    engine.work['chains_run'] += 1
    return dict(gam_o=gam_o, gam_n=gam_n, sigs=sigs, taus=taus, lik=lik, bad=int(info.item()))
