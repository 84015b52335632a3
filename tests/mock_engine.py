"""A numpy / host-emulation stand-in for FoKL._engine.Engine, so that the host side of the selection loop
(FoKL/_selection.py: batching, speculation, roll-back, bookkeeping) can be tested without a GPU.

TEST INFRASTRUCTURE.  Design columns come from the oracle's basis loop, Gram products from numpy, and the candidate
math (eigensolver, BIC, Philox Gibbs chain, kill loop) from tests/host_emu -- the kernels' own math compiled for the
host.  Only the interface forward_select() uses is provided."""
import numpy as np
import torch

import emu
import fokl_oracle as fo
from FoKL import _lib
from FoKL._engine import CandidateResult, Engine


class MockEngine:
    def __init__(self, x, y, phis, kernel, dist=None):
        """x, y: this rank's row shard.  dist: an initialised torch.distributed (gloo) module for the multi-rank form --
        partial Gram blocks and data moments are summed over the ranks like Engine._append_built / begin_fit do."""
        self.torch = torch
        self.device = torch.device('cpu')
        self.dist, self.group, self.world, self.rank = None, None, 1, 0
        if dist is not None and dist.get_world_size() > 1:
            self.dist, self.world, self.rank = dist, dist.get_world_size(), dist.get_rank()
        self.profile = None
        self.x = np.ascontiguousarray(x, dtype=np.float64)
        self.y = np.ascontiguousarray(y, dtype=np.float64).reshape(-1)
        self.phis, self.kernel = phis, kernel
        self.calls = []              # (kind, detail) log of the device work the loop asked for
        n = len(self.y)
        mom = self._sum(np.array([float(n), float(self.y.sum()), float(self.y @ self.y)]))
        self.n_global, self.sum_y, self.yty = int(round(mom[0])), float(mom[1]), float(mom[2])
        self.X = np.ones((n, 1))
        self._gram()

    def _sum(self, a):
        if self.dist is None:
            return a
        t = torch.from_numpy(np.ascontiguousarray(a, dtype=np.float64).copy())
        self.dist.all_reduce(t, op=self.dist.ReduceOp.SUM)
        return t.numpy()

    # forward_select's parity mode indexes engine.G / engine.Xty as torch tensors (Engine keeps them in HBM)
    @property
    def G(self):
        return torch.from_numpy(self._G)

    @property
    def Xty(self):
        return torch.from_numpy(self._Xty)

    def _gram(self):
        self._G = np.ascontiguousarray(self._sum(self.X.T @ self.X))
        self._Xty = np.ascontiguousarray(self._sum(self.X.T @ self.y))
        self.P = self.X.shape[1]

    # ---- K1 / K2 / compaction ----------------------------------------------------------------------------------
    def append_terms(self, terms):
        terms = np.asarray(terms, dtype=np.int64)
        self.calls.append(('append', len(terms)))
        if len(terms) == 0:
            return
        cols = fo.basis_columns(self.x, terms, self.phis, self.kernel)
        self.X = np.hstack([self.X, cols])
        self._gram()

    def compact(self, keep):
        keep = np.asarray(keep, dtype=np.int64)
        self.calls.append(('compact', len(keep)))
        if len(keep) == self.P:
            return
        self.X = self.X[:, keep]
        self._G = self._G[np.ix_(keep, keep)].copy()
        self._Xty = self._Xty[keep].copy()
        self.P = len(keep)

    def truncate(self, p):
        """Drop the columns from p on (roll-back of a speculative append)."""
        self.calls.append(('truncate', p))
        self.X = self.X[:, :p]
        self._G = self._G[:p, :p].copy()
        self._Xty = self._Xty[:p].copy()
        self.P = p

    # ---- K3 / K4 ----------------------------------------------------------------------------------------------
    def make_hypers(self, a, b, atau, btau, sigsqd0, tausqd0, draws):
        return dict(a=float(a), b=float(b), atau=float(atau), btau=float(btau), sigsqd0=float(sigsqd0),
                    tausqd0=float(tausqd0), yty=self.yty, sum_y=self.sum_y, n=self.n_global, draws=int(draws))

    def evaluate(self, col_sets, hyp, rng_mode=_lib.RNG_NONE, run_chain=None, seed=0, stream_ids=None, variates=None,
                 sign_fix=None, want_betas=False, want_eig=False, refine_tol=1e-7, gram=None):
        G, Xty = gram if gram is not None else (self._G, self._Xty)
        self.calls.append(('evaluate', [len(s) for s in col_sets]))
        if variates is not None and torch.is_tensor(variates):
            variates = variates.numpy()
        if sign_fix is not None and torch.is_tensor(sign_fix):
            sign_fix = sign_fix.numpy()
        n_cand = len(col_sets)
        p = np.array([len(s) for s in col_sets], dtype=np.int64)
        D = int(hyp['draws'])
        h0, h1 = int(np.ceil(D / 2)), int(np.ceil(D / 2 + 1))
        res = CandidateResult()
        res.p, res.draws = p, D
        res.vec_off = np.concatenate([[0], np.cumsum(p)[:-1]]).astype(np.int64)
        res.mat_off = np.concatenate([[0], np.cumsum(p * p)[:-1]]).astype(np.int64)
        ev = np.zeros(n_cand)
        stats, betas, betahat, qs, lambs = [], [], [], [], []
        for c, cols in enumerate(col_sets):
            chain = rng_mode != _lib.RNG_NONE and (run_chain is None or bool(run_chain[c]))
            pc = len(cols)
            var_c = sf_c = None
            if chain and rng_mode == _lib.RNG_INJECTED:
                # packed like the C ABI: candidate c at D * (vec_off[c] + 2 c), D rows of [z_0 .. z_{p-1}, g1, g2]
                o = D * (int(res.vec_off[c]) + 2 * c)
                var_c = np.asarray(variates).reshape(-1)[o:o + D * (pc + 2)].reshape(D, pc + 2)
                if sign_fix is not None:
                    sf_c = np.asarray(sign_fix).reshape(-1)[int(res.vec_off[c]):int(res.vec_off[c]) + pc]
            r = emu.candidate(G, Xty, np.asarray(cols, dtype=np.int32), hyp, rng_mode=rng_mode if chain else 0,
                              seed=int(seed), stream=int(stream_ids[c]) if stream_ids is not None else 0,
                              variates=var_c, sign_fix=sf_c)
            qs.append(r['Q'].T.reshape(-1))             # row r of the p x p block = eigenvector r, like the device
            lambs.append(r['lamb'])
            ev[c] = r['ev']
            betahat.append(r['betahat'])
            b = r['betas'] if chain else np.zeros((D, len(cols)))
            betas.append(b.reshape(-1))
            stats.append(np.stack([b[h1:].mean(axis=0), b[h1:].std(axis=0), b[h0:].mean(axis=0)]).reshape(-1)
                         if chain else np.zeros(3 * len(cols)))
        res.ev = ev
        res.info = np.zeros(n_cand, dtype=np.int32)
        res.stats = torch.from_numpy(np.concatenate(stats))
        res.betas = torch.from_numpy(np.concatenate(betas))
        res.betahat = torch.from_numpy(np.concatenate(betahat))
        res.sigs = res.taus = None
        res.lamb = torch.from_numpy(np.concatenate(lambs)) if want_eig else None
        res.Q = torch.from_numpy(np.concatenate(qs)) if want_eig else None
        self._refine(res, col_sets, refine_tol, gram)
        return res

    def _refine(self, res, col_sets, refine_tol, gram):
        res.refined = np.zeros(len(res.ev), dtype=bool)
        if refine_tol is not None and refine_tol > 0:
            bad = self.refine_mask(res.ev, res.p, refine_tol)
            res.refined = bad
            assert gram is None or not np.any(bad), 'refinement needs the columns in X'
            for c in np.nonzero(bad)[0]:
                self.calls.append(('refine', len(col_sets[c])))
                res.ev[c] = self.residual_bic(col_sets[c], res.betahat[res.vec_off[c]:res.vec_off[c] + res.p[c]])

    def residual_bic(self, cols, betahat):
        r = self.y - self.X[:, np.asarray(cols, dtype=np.int64)] @ betahat.numpy()
        n = float(self.n_global)
        sr, srr = self._sum(np.array([r.sum(), r @ r]))
        siglik = srr / n - (sr / n) ** 2
        with np.errstate(all='ignore'):
            lik = -(n / 2) * np.log(siglik) - (n - 1) / 2
        return len(cols) * np.log(n) - 2 * lik

    # ---- update fits (FoKL/_update.py) -------------------------------------------------------------------------------
    def sym_eigh(self, A):
        A = np.ascontiguousarray(A.numpy(), dtype=np.float64)
        p = A.shape[0]
        r = emu.candidate(A, np.zeros(p), np.arange(p, dtype=np.int32),
                          dict(a=1.0, b=1.0, atau=1.0, btau=1.0, sigsqd0=1.0, tausqd0=1.0, yty=1.0, sum_y=0.0, n=10, draws=1))
        self.calls.append(('sym_eigh', p))
        return torch.from_numpy(r['lamb'].copy()), torch.from_numpy(np.ascontiguousarray(r['Q']))

    def residual_sse(self, p, betahat):
        r = self.y - self.X[:, :p] @ betahat.numpy()
        return float(self._sum(np.array([r @ r]))[0])

    def update_chain(self, spec, arrays, rng_mode, seed=0, stream_id=0, variates=None):
        self.calls.append(('update_chain', (spec['mode'], spec['po'], spec['pn'])))
        arr = {k: v.numpy() for k, v in arrays.items()}
        r = emu.update_chain(spec['mode'], spec['po'], spec['pn'], spec['draws'], spec['a_star'], spec['atau_star'],
                             spec['b'], spec['btau'], spec['sigsqd0'], spec['yty'], spec['squerr'], spec['n'], arr,
                             variates=variates if rng_mode == _lib.RNG_INJECTED else None, seed=seed, stream=stream_id)
        return dict(gam_o=torch.from_numpy(r['gam_o'].copy()), gam_n=torch.from_numpy(r['gam_n'].copy()),
                    sigs=torch.from_numpy(r['sigs']), taus=torch.from_numpy(r['taus']), lik=torch.from_numpy(r['lik']),
                    bad=r['bad'])

    def mark(self):
        return None

    def evaluate_launch(self, col_sets, hyp, gram=None, side=False, after=None, **kw):
        self.calls.append(('launch', 'side' if side else 'main'))
        res = self.evaluate(col_sets, hyp, gram=gram, refine_tol=None, **kw)
        eng = self

        class Handle:
            def finish(self, refine_tol=None):
                eng._refine(res, col_sets, refine_tol, gram)
                return res
        return Handle()

    def gram_state(self):
        return (self._G, self._Xty)

    refine_mask = Engine.refine_mask

    def kill_loop(self, cols, cand_pos, bv0, bv1, hyp, threshav, threshstda, threshstdb, icpt, evmin, aic_adj, start):
        self.calls.append(('kill_loop', len(cand_pos)))
        return emu.kill_loop(self._G, self._Xty, cols, cand_pos, bv0, bv1, hyp, threshav=threshav, threshstda=threshstda,
                             threshstdb=threshstdb, icpt=icpt, evmin=evmin, aic_adj=aic_adj, start=start)

    def kill_loop_launch(self, *args, **kw):
        r = self.kill_loop(*args)
        self.calls.append(('kill_launch', 0))

        class Handle:
            def finish(self):
                return r
        return Handle()

    def _allreduce(self, t):
        if self.dist is not None:
            self.dist.all_reduce(t, op=self.dist.ReduceOp.SUM)
