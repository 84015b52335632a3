"""Host side of the forward-selection loop of `FoKL.fit` (reference: src/FoKL/FoKLRoutines.py:1561-1760).

The loop is inherently sequential in its stages; what it asks of the device per substage s is
  A(s)  append the new terms' columns to X and extend the Gram (K1 + K2),
  B(s)  evaluate the full model (eigensolver + BIC + Gibbs chain + column statistics, FR:1650),
  C(s)  the kill loop FR:1666-1690 -- literally (one `gibbs`-equivalent evaluation per proposal: parity mode, eager
        mode, degenerate Grams) or on the fast path: one device launch that walks all proposals on the sweep-operator
        tableau of the model's Gram, followed by one batch of chains for the accepted models whose draws can influence a
        later decision, each speculated outcome re-checked against its true chain -- and
  D(s)  drop the accepted kills (column / Gram compaction) and the bookkeeping of FR:1701-1721.
The fast path is exact: a proposal's outcome depends only on the kill set accepted so far and, through one threshold,
on the draws of the last accepted model (FR:1670-1690).  `forward_select` drives these phases as a software pipeline
(the batch of C(s) next to B(s + 1)); see the comment above its driver loop.
"""
import os

import numpy as np

from . import _lib


def distinct_permutations(v):
    """All distinct orderings of the multiset v as integer rows in ascending lexicographic order --
    what np.unique(perms(v), axis=0) returns at FR:1616, without enumerating len(v)! permutations."""
    a = sorted(int(t) for t in v)
    n = len(a)
    rows = [tuple(a)]
    while True:
        i = n - 2
        while i >= 0 and a[i] >= a[i + 1]:
            i -= 1
        if i < 0:
            break
        j = n - 1
        while a[j] <= a[i]:
            j -= 1
        a[i], a[j] = a[j], a[i]
        a[i + 1:] = reversed(a[i + 1:])
        rows.append(tuple(a))
    return np.array(rows, dtype=np.int64).reshape(len(rows), n)


def first_partition(ind, m, sett):
    """FR:1605-1613."""
    v = [0] * m
    left = ind
    while left:
        for j in range(sett):
            v[j] += 1
            left -= 1
            if left == 0:
                break
    return v


def next_partition(v, m, way3):
    """FR:1722-1740; mutates v, returns False when this `ind` is exhausted."""
    if m == 1:
        return False
    if way3:
        if v[1] > v[2]:
            v[0] += 1
            v[1] -= 1
            return True
        if v[2]:
            v[1] += 1
            v[2] -= 1
            if v[1] > v[0]:
                v[0] += 1
                v[1] -= 1
            return True
        return False
    if v[1]:
        v[0] += 1
        v[1] -= 1
        return True
    return False


class NumpyVariates:
    """Parity RNG: consumes the global legacy numpy stream exactly like one reference `gibbs` call
    (FR:1527, 1541, 1547): per draw normal(size=(p, 1)), gamma(astar, .), gamma(atau_star, .).
    np.random.gamma(k, s) == s * standard_gamma(k) bitwise, so the stream does not depend on values."""
    mode = _lib.RNG_INJECTED

    def __init__(self, a, atau, n, draws):
        self.a, self.atau, self.n, self.draws = a, atau, n, draws

    def draw(self, p):
        astar = self.a + 1 + self.n / 2 + p / 2
        atau_star = self.atau + (p - 1) / 2
        out = np.empty((self.draws, p + 2))
        normal, sgamma = np.random.normal, np.random.standard_gamma
        for k in range(self.draws):
            out[k, :p] = normal(loc=0, scale=1, size=(p, 1))[:, 0]
            out[k, p] = sgamma(astar)
            out[k, p + 1] = sgamma(atau_star)
        return out


class PhiloxVariates:
    """Free-running RNG on the device; one seed per fit drawn from the global numpy RNG (so np.random.seed
    still makes a fit reproducible), one stream per `gibbs` call index."""
    mode = _lib.RNG_PHILOX

    def __init__(self, engine=None):
        self.seed = int(np.random.randint(0, 2 ** 31 - 1)) * (2 ** 31) + int(np.random.randint(0, 2 ** 31 - 1))
        if engine is not None and engine.dist is not None:      # every rank must walk the same chain
            t = engine.torch.tensor([self.seed], dtype=engine.torch.int64, device=engine.device)
            engine.dist.broadcast(t, src=0, group=engine.group)
            self.seed = int(t.item())


def forward_select(engine, hy, m, n_phis, console=True, rng='philox', eager=False, on_substage=None,
                   recorder=None, pipeline=None, prefetch=False):
    """Run the selection loop on an Engine that has a dataset bound (engine.begin_fit).

    hy: dict with a, b, atau, btau, tolerance, total_draws, gimmie, way3, threshav, threshstda, threshstdb, aic.
    pipeline (default: on for the free-running fast path): evaluate the chains that verify substage s together with
    the full model of substage s + 1 (see the module docstring of the driver loop below); the result is bit-identical
    to pipeline=False.
    prefetch (opt-in): build the columns and most of the Gram block of substage s + 1 on a second stream while the
    candidate stage of s runs (Engine.prefetch_terms).  Every substage's block is then formed by the same two-part
    scheme whether or not it was started ahead of time, so the fit does not depend on the timing (it differs from the
    prefetch=False fit in the last bits of the Gram: another summation order).
    Returns dict(betas=(D x P) numpy of the chosen model, mtx, evs, n_gibbs, n_batches)."""
    torch = engine.torch
    n = engine.n_global
    D = int(hy['total_draws'])
    a, b, atau, btau = hy['a'], hy['b'], hy['atau'], hy['btau']
    hyp = engine.make_hypers(a, b, atau, btau, b / (1 + a), btau / (1 + atau), D)
    aic_adj = (2 - np.log(n)) if hy['aic'] else 0.0
    if rng == 'numpy':
        src = NumpyVariates(a, atau, n, D)
        seed = 0
    else:
        src = PhiloxVariates(engine)
        seed = src.seed
    mode = src.mode
    sett = 1 if m == 1 else (3 if hy['way3'] else 2)
    tolerance = hy['tolerance']
    literal = mode == _lib.RNG_INJECTED or eager
    if pipeline is None:
        pipeline = True
    pipeline = False if (literal or not pipeline) else pipeline      # True, or 'always' (tests: speculate even when
    #                                                                   the stopping rule is predicted to fire)
    world = engine.world if engine.dist is not None else 1
    # the kill loop of a substage starts from the full model's eigendecomposition (fokl_kill_params.lamb / Qt) unless
    # switched off (tests, tools); the tests' CPU stand-in engine has no such path
    kill_from_eig = bool(hy.get('kill_from_eig', True)) and not literal and getattr(engine, 'kill_from_eig', False)

    # selection state (FR:1590-1604)
    terms = np.zeros((0, m), dtype=np.int64)     # damtx: row j <-> column j + 1 of X
    evs = []
    best = None            # (betas tensor, mtx)
    last = None
    greater = 0
    cnt = dict(calls=0, gibbs=0, batches=0)      # `gibbs` call index (Philox stream id), gibbs-equivalents, launches

    def run(col_sets, chains, ids, refine=True):
        """One synchronous batch of spectral evaluations on the current Gram."""
        cnt['batches'] += 1
        sid = np.asarray(ids, dtype=np.uint64)
        if mode == _lib.RNG_INJECTED:
            # Parity mode.  Eigenvector signs are arbitrary and LAPACK's choice is what pairs each injected
            # normal with a direction (S = Q diag(...), FR:1525-1528), so the signs of the device eigenvectors are
            # aligned with scipy.linalg.eigh of the *same Gram bits* before the chain runs.  Only +-1 factors come
            # from the host; every number in the result is computed on the device.
            from scipy.linalg import eigh as _eigh
            pre = engine.evaluate(col_sets, hyp, rng_mode=_lib.RNG_NONE, want_eig=True)
            signs = []
            for c, cols in enumerate(col_sets):
                idx = torch.as_tensor(np.asarray(cols, dtype=np.int64), device=engine.device)
                g_sub = engine.G.index_select(0, idx).index_select(1, idx).cpu().numpy()
                if recorder is not None:
                    xty_sub = engine.Xty.index_select(0, idx).cpu().numpy()
                    recorder(tuple(map(tuple, terms[np.asarray(cols[1:], dtype=np.int64) - 1])), g_sub, xty_sub)
                pc = len(cols)
                q_dev = pre.Q[pre.mat_off[c]:pre.mat_off[c] + pc * pc].view(pc, pc).cpu().numpy().T
                with np.errstate(all='ignore'):
                    _, q_ref = _eigh(g_sub)
                sg = np.sign(np.sum(q_dev * q_ref, axis=0))
                sg[sg == 0] = 1.0
                signs.append(sg)
            variates = np.concatenate([
                (v if v is not None else np.zeros((D, len(s_) + 2))).reshape(-1) for v, s_ in zip(chains, col_sets)])
            flags = np.array([v is not None for v in chains], dtype=np.uint8)
            out = engine.evaluate(col_sets, hyp, rng_mode=mode, run_chain=flags, seed=seed, stream_ids=sid,
                                  variates=variates, sign_fix=np.concatenate(signs), want_betas=True)
            if recorder is not None and hasattr(recorder, 'on_result'):
                for c, cols in enumerate(col_sets):
                    recorder.on_result(tuple(map(tuple, terms[np.asarray(cols[1:], dtype=np.int64) - 1])),
                                       float(out.ev[c]))
            return out
        flags = np.array([bool(v) for v in chains], dtype=np.uint8)
        any_chain = bool(flags.any())
        return engine.evaluate(col_sets, hyp, rng_mode=mode if any_chain else _lib.RNG_NONE, run_chain=flags,
                               seed=seed, stream_ids=sid, want_betas=any_chain, refine_tol=1e-7 if refine else None)

    # ---- the partition walk (FR:1605-1616, 1722-1740) as a pure function of (ind, part) --------------------------
    def walk_next(ind, part):
        part = list(part)
        if next_partition(part, m, hy['way3']):
            return ind, part
        ind += 1
        if ind > n_phis:
            return None
        return ind, first_partition(ind, m, sett)

    # ---- phase A: append the new terms' columns (K1 + K2) -----------------------------------------------------------
    can_split = bool(prefetch) and hasattr(engine, 'prefetch_terms')      # (the tests' CPU stand-in engine has none)

    def open_substage(ind, part, terms_in, p_stable=None):
        """p_stable: the model width when the PREVIOUS substage began -- its first p_stable columns were final from
        then on, so the build of this substage may have been started ahead of time (see start_next)."""
        vecs = distinct_permutations(part)
        p_old = engine.P                     # columns of the model accepted so far (incl. intercept)
        if can_split:
            engine.append_terms(vecs, p_stable=p_stable)
        else:
            engine.append_terms(vecs)
        return dict(ind=ind, part=list(part), vecs=vecs, vm=vecs.shape[0], p_old=p_old, p_stable=p_stable,
                    terms=np.concatenate([terms_in, vecs], axis=0), full=list(range(engine.P)))

    def start_next(S, after):
        """Build ahead: K1 + the stable part of K2 of the substage that follows S, next to S's candidate stage."""
        if not can_split:
            return
        nxt = walk_next(S['ind'], S['part'])
        if nxt is None:
            return
        engine.prefetch_terms(distinct_permutations(nxt[1]), S['p_old'], after=after, wait_eig=True)

    # ---- phase B: the full model (FR:1650), evaluated as (launch, finish) so that other work can ride along ------------
    def full_launch(S):
        cnt['calls'] += 1
        cnt['gibbs'] += 1
        S['full_id'] = cnt['calls']
        if literal:
            return None
        cnt['batches'] += 1
        # a wide full model keeps its eigendecomposition: the nested chains of its accepted sub-models start from it
        S['keep_eig'] = bool(use_nested and len(S['full']) >= nested_min_p)
        # ... and every full model hands it to its kill loop, whose tableau is then formed from it (Engine.kill_loop_launch)
        return engine.evaluate_launch([S['full']], hyp, rng_mode=mode, run_chain=np.ones(1, dtype=np.uint8), seed=seed,
                                      stream_ids=np.asarray([S['full_id']], dtype=np.uint64), want_betas=True,
                                      want_eig=bool(S['keep_eig'] or kill_from_eig))

    def full_finish(S, handle):
        nonlocal terms
        full = S['full']
        if handle is None:
            terms = S['terms']               # the parity recorder keys candidates by their term rows
            chain_arg = src.draw(len(full)) if mode == _lib.RNG_INJECTED else True
            res = run([full], [chain_arg], [S['full_id']])
        else:
            res = handle.finish(1e-7)
            if S.get('keep_eig') and not (int(res.info[0]) & 6):
                S['head'] = (np.asarray(full, dtype=np.int32), res.lamb, res.Q)
            if kill_from_eig and res.lamb is not None and not (int(res.info[0]) & 6):
                S['eig'] = (res.lamb, res.Q)
        S['ev'] = float(res.ev[0]) + aic_adj * (S['terms'].shape[0] + 1)
        # a (nearly) interpolating full model: its BIC came from the residual pass because the Gram-only form has lost
        # its digits to cancellation (Engine.refine_mask) -- the device kill loop scores with the same Gram-only form,
        # so this substage takes the literal path
        S['refined'] = bool(getattr(res, 'refined', np.zeros(1, dtype=bool))[0])
        st = getattr(res, 'stats_host', None)          # (read back with the BIC in one copy)
        st = res.stats_of(0).cpu().numpy() if st is None else st[:3 * len(full)].reshape(3, len(full))
        S['betas'] = res.betas_of(0)
        new_cols = np.arange(S['p_old'], S['p_old'] + S['vm'])
        with np.errstate(all='ignore'):
            bv0 = np.abs(st[0, new_cols])                          # |mean| over rows h1.. (FR:1656)
            bv1 = st[1, new_cols] / np.abs(st[2, new_cols])        # std / |mean over rows h0..| (FR:1657-58)
        order = np.argsort(bv0, kind='quicksort')
        S['bv0'], S['bv1'], S['cand_cols'] = bv0[order], bv1[order], new_cols[order]
        S['icpt'] = abs(float(st[2, 0]))                           # |mean(beters[h0:, 0])| (FR:1671)
        with np.errstate(invalid='ignore'):
            S['always'] = S['bv1'] > hy['threshstdb']              # proposed whatever the intercept is
            S['maybe'] = S['bv1'] > hy['threshstda']

    def prop_mask(S, icpt_now):
        with np.errstate(invalid='ignore'):
            return S['always'] | (S['maybe'] & (S['bv0'] < hy['threshav'] * icpt_now))

    def proposals(S, start, icpt_now):
        return (np.nonzero(prop_mask(S, icpt_now)[start:])[0] + start).tolist()

    # ---- phase C, literal order (FR:1666-1692): one `gibbs`-equivalent evaluation (eig + chain) per proposal ------------
    def kill_literal(S, parity):
        """In parity mode the numpy stream is consumed test by test; otherwise all proposals of a round run side by
        side (also the fall-back of the fast path when the Gram is not numerically positive definite)."""
        full, cand_cols, vm = S['full'], S['cand_cols'], S['vm']
        killed, evmin, cur, icpt, cur_betas = [], S['ev'], 0, S['icpt'], S['betas']
        while cur < vm:
            props = proposals(S, cur, icpt)
            if not props:
                break
            sets = []
            for i in props:
                drop = set(killed) | {int(cand_cols[i])}
                sets.append([c for c in full if c not in drop])
            ids = [cnt['calls'] + r + 1 for r in range(len(props))]
            accepted = None
            if parity:
                for r, i in enumerate(props):
                    cnt['calls'] += 1
                    cnt['gibbs'] += 1
                    rr = run([sets[r]], [src.draw(len(sets[r]))], [ids[r]])
                    evt = float(rr.ev[0]) + aic_adj * len(sets[r])
                    if evt < evmin:
                        accepted = (i, evt, rr, 0)
                        break
            else:
                rr = run(sets, [True] * len(sets), ids)
                for r, i in enumerate(props):
                    evt = float(rr.ev[r]) + aic_adj * len(sets[r])
                    if evt < evmin:
                        accepted = (i, evt, rr, r)
                        break
                tested = (props.index(accepted[0]) + 1) if accepted else len(props)
                cnt['calls'] += tested
                cnt['gibbs'] += tested
            if accepted is None:
                break
            i, evt, rr, slot = accepted
            killed.append(int(cand_cols[i]))
            evmin = evt
            cur_betas = rr.betas_of(slot).clone()
            icpt = abs(float(rr.stats_of(slot)[2, 0].item()))
            cur = i + 1
        return dict(killed=killed, evmin=evmin, betas=cur_betas, calls=cnt['calls'], gibbs=cnt['gibbs'])

    # ---- chains of accepted models: launch / collect ----------------------------------------------------------------
    # The models of a batch are independent: with several ranks each evaluates every world-th one (all ranks hold the
    # full Gram) and the two scalars per model that drive the loop are summed into place by one allreduce, so every
    # rank takes the same decisions.
    # Nested path (csrc/nested.cu): the accepted models of a substage are nested (each is its predecessor minus one
    # column), and all the loop reads from their chains is |mean intercept|.  A batch of wide models is therefore
    # evaluated by ONE eigensolver run + one secular-equation step per deleted column instead of one eigensolver run per
    # model; their BIC is the kill loop's own (the same Gram-only quantity).  The last accepted model -- whose draws the
    # substage returns (FR:1690) -- always takes the ordinary path.
    use_nested = (mode == _lib.RNG_PHILOX and hasattr(engine, 'nested_chains_launch') and
                  bool(hy.get('nested_chains', True)))
    # below ~250 columns a batch of cold models (cluster eigensolver, ~0.1 ms per model in a batch of 160) beats the
    # sequential chain of updates (~0.1 ms per step + the head's solve): measured on cfg4 (widths <= 227: 151 ms per
    # fit without, 188 ms with) and cfg5 (0.90 s per fit with 256 or 160, 0.95 s with 384)
    nested_min_models, nested_min_p = int(hy.get('nested_min_models', 4)), int(hy.get('nested_min_p', 256))

    nested_head_max_steps = int(hy.get('nested_head_max_steps', 96))

    def nested_ok(todo, idx):
        if not use_nested or len(idx) < nested_min_models or len(todo[idx[0]]['cols']) < nested_min_p:
            return False
        if any('ev_dev' not in todo[i] for i in idx):
            return False
        ev = np.array([todo[i]['ev_dev'] - aic_adj * len(todo[i]['cols']) for i in idx])
        pw = np.array([len(todo[i]['cols']) for i in idx])
        return not bool(np.any(engine.refine_mask(ev, pw)))       # Gram-only BICs must be trustworthy (no residual pass here)

    def chains_launch(todo, which, gram=None, side=False, after=None):
        """Returns a list of parts (kind, handle, indices) or None."""
        if not which:
            return None
        parts = []
        cold = list(which)
        last = len(todo) - 1
        nest = [i for i in which if i != last]
        if nested_ok(todo, nest):
            cnt['batches'] += 1
            head = todo[nest[0]].get('head')
            if head is not None and len(head[0]) - len(todo[nest[0]]['cols']) > nested_head_max_steps:
                head = None          # too many steps from the full model to this run's first model: cold solve instead
            h = engine.nested_chains_launch([todo[i]['cols'] for i in nest], hyp, seed,
                                            np.asarray([todo[i]['stream'] for i in nest], dtype=np.uint64), gram=gram,
                                            side=side, after=after, head=head)
            parts.append(('nested', h, nest, gram))
            cold = [i for i in which if i == last]
        if cold:
            cnt['batches'] += 1
            h = engine.evaluate_launch([todo[i]['cols'] for i in cold], hyp, rng_mode=mode,
                                       run_chain=np.ones(len(cold), dtype=np.uint8), seed=seed,
                                       stream_ids=np.asarray([todo[i]['stream'] for i in cold], dtype=np.uint64),
                                       want_betas=True, gram=gram, side=side, after=after)
            parts.append(('cold', h, cold, gram))
        return parts

    def chains_collect(todo, which, parts, vals, refine):
        """Fill vals[i] = (|mean intercept|, BIC) and todo[i]['rr' / 'slot'] for the models `which`; returns True if a
        model's Gram-only BIC is not trustworthy and refine is off (see Engine.refine_mask)."""
        if parts is None:
            return False
        need = False
        for kind, handle, idx, gram in parts:
            if kind == 'nested':
                r = handle.finish()
                if r['ok']:
                    for slot, i in enumerate(idx):
                        vals[i, 0] = abs(float(r['mean0'][slot]))
                        vals[i, 1] = float(todo[i]['ev_dev'])
                        todo[i]['rr'], todo[i]['slot'] = None, None
                    continue
                # equal eigenvalues met on the way (or a failed chain): the ordinary path, now
                if os.environ.get('FOKL_B200_DEBUG'):
                    print('nested run not ok: status', r['status'], 'models', len(idx), 'widths', len(todo[idx[0]]['cols']),
                          '..', len(todo[idx[-1]]['cols']), 'finite', bool(np.all(np.isfinite(r['mean0']))), flush=True)
                cnt['batches'] += 1
                handle = engine.evaluate_launch([todo[i]['cols'] for i in idx], hyp, rng_mode=mode,
                                                run_chain=np.ones(len(idx), dtype=np.uint8), seed=seed,
                                                stream_ids=np.asarray([todo[i]['stream'] for i in idx], dtype=np.uint64),
                                                want_betas=True, gram=gram)
            rr = handle.finish(1e-7 if refine else None)
            stats_h = getattr(rr, 'stats_host', None)      # one read-back for the whole batch
            if stats_h is None:
                stats_h = rr.stats.cpu().numpy()
            for slot, i in enumerate(idx):
                # mean(beters[h0:, 0]) of this model: row 2, column 0 of its 3 x p statistics block
                vals[i, 0] = abs(float(stats_h[3 * rr.vec_off[slot] + 2 * rr.p[slot]]))
                vals[i, 1] = float(rr.ev[slot]) + aic_adj * len(todo[i]['cols'])
                todo[i]['rr'], todo[i]['slot'] = rr, slot
            need = need or ((not refine) and bool(np.any(engine.refine_mask(rr.ev, rr.p))))
        return need

    def share_of(todo):
        """(indices this rank evaluates, owner rank of every index or None).  Wide nested batches are cut into one
        contiguous run per rank (a run costs one eigensolver + its steps); otherwise models are dealt round-robin."""
        n_t = len(todo)
        if world == 1:
            return list(range(n_t)), None
        if nested_ok(todo, list(range(n_t - 1))) and n_t - 1 >= world * nested_min_models:
            w = np.array([float(len(rd['cols'])) ** 2 for rd in todo[:-1]])
            cum = np.cumsum(w) / w.sum()
            owners = np.minimum((cum * world - 1e-9).astype(np.int64), world - 1)
            owners = np.concatenate([owners, [world - 1]])
        else:
            owners = np.arange(n_t) % world
        return [i for i in range(n_t) if owners[i] == engine.rank], owners

    def my_share(todo):
        return share_of(todo)[0]

    def chains_reduce(todo, vals, need_refine):
        """All ranks: sum the per-model scalars into place; returns True if any rank flagged a model for refinement."""
        if world > 1:
            vals[len(todo), 0] = float(need_refine)
            if hasattr(engine, 'ctl_allreduce'):
                vals[:] = engine.ctl_allreduce(vals)
            else:                                   # (the tests' CPU stand-in engine)
                t = torch.from_numpy(vals).to(engine.device)
                engine._allreduce(t)
                vals[:] = t.cpu().numpy()
            owners = share_of(todo)[1]
            for i, rd in enumerate(todo):
                rd['owner'] = int(owners[i])
            need_refine = vals[len(todo), 0] > 0
        return bool(need_refine)

    def chains_now(todo, launched=None):
        """Synchronous form: evaluate `todo` on the current Gram (X still holds every column: refinement possible).
        launched: (indices, parts) of a chains_launch(todo, indices) made earlier for exactly this list."""
        vals = np.zeros((len(todo) + 1, 2))
        mine = my_share(todo) if launched is None else launched[0]
        first = chains_launch(todo, mine) if launched is None else launched[1]
        if world > 1:
            need = chains_collect(todo, mine, first, vals, refine=False)
            if chains_reduce(todo, vals, need):
                # some model needs the N-length residual pass (a collective): evaluate the batch replicated
                vals = np.zeros((len(todo) + 1, 2))
                everything = list(range(len(todo)))
                chains_collect(todo, everything, chains_launch(todo, everything), vals, refine=True)
                for rd in todo:
                    rd['owner'] = -1
        else:
            chains_collect(todo, mine, first, vals, refine=True)
        for i, rd in enumerate(todo):
            rd['icpt_true'], rd['ev_true'] = float(vals[i, 0]), float(vals[i, 1])

    # ---- phase C, fast path: a generator that yields when it needs chains --------------------------------------------
    def kill_fast(S):
        """The whole kill loop runs on the device in one launch (fokl_kill_loop: sweep-operator inverse of the model's
        Gram, one O(p^2) reverse sweep per accepted kill).  The chains of the accepted models -- which only feed later
        rounds' threshold through |mean intercept| (FR:1671) and the final draws (FR:1690) -- are run afterwards as one
        batch, only for the rounds whose remaining proposals can depend on that threshold, and every such round is
        re-checked against its true chain (the loop is re-run from the first round that differs), so the outcome is
        exactly that of the sequential loop.  Round k's chain matters only through the threshold it sets for the
        candidates examined before the next acceptance, i.e. indices (i_k, i_{k+1}] of the sorted list (to the end for
        the last round): if none of them is threshold-dependent (threshstda < bv1 <= threshstdb) the chain of round k
        is never looked at.

        Protocol: `yield ('kill', model, pos, icpt, evmin, start)` asks the driver for one fokl_kill_loop launch and is
        answered with its result (so the driver can do host work while the kernel runs); `yield ('chains', todo,
        outlook)` asks for the chains of `todo` (the driver fills icpt_true / ev_true / rr / slot / owner of every
        entry) and is answered with 'done' -- or with 'speculated' if the driver has, on the strength of `outlook` (the
        outcome if every check passes; None after the first request), already compacted the engine and moved on; then a
        failed check makes the generator `yield ('rollback',)` (the driver restores the substage's columns and Gram)
        before it continues.  Returns the substage's outcome."""
        full, cand_cols, vm, bv0, bv1 = S['full'], S['cand_cols'], S['vm'], S['bv0'], S['bv1']
        with np.errstate(invalid='ignore'):
            bv1_sens = (bv1 > hy['threshstda']) & ~(bv1 > hy['threshstdb'])
        sens_cum = np.concatenate([[0], np.cumsum(bv1_sens)])      # threshold-dependent candidates in [0, j)
        bv0_finite = bv0[:int(np.count_nonzero(~np.isnan(bv0)))]   # argsort puts NaN last

        def span(k_):
            return rounds[k_]['i'] + 1, (vm if k_ == len(rounds) - 1 else rounds[k_ + 1]['i'] + 1)

        state = dict(killed=[], evmin=S['ev'], cur=0, icpt=S['icpt'], calls=cnt['calls'], gibbs=cnt['gibbs'])
        rounds = []        # accepted kills: dict(i, cols, stream, icpt_used, gibbs_after [, icpt_true, ev_true])
        first_request = True
        while True:
            if state['cur'] < vm and proposals(S, state['cur'], state['icpt']):
                # column bookkeeping on numpy masks (the lists are up to 160 models x 220 columns per substage)
                alive = np.ones(len(full), dtype=bool)
                alive[np.asarray(state['killed'], dtype=np.int64)] = False
                model = np.nonzero(alive)[0].astype(np.int32)
                where = np.ones(len(full), dtype=np.int32)            # killed candidates: any valid position
                where[model] = np.arange(len(model), dtype=np.int32)
                pos = where[cand_cols]
                r = yield ('kill', model, pos, state['icpt'], state['evmin'], state['cur'])
                cnt['batches'] += 1
                if r['bad']:
                    # Gram not numerically positive definite (p >= N regimes): literal loop, one spectral evaluation
                    # (eig + chain) per proposal, proposals of a round side by side.  (Only ever seen on the first
                    # launch of a substage: a later one works on a principal sub-matrix of a factorisable Gram.)
                    return kill_literal(S, parity=False)
                n_acc = int(r['n_acc'])
                if n_acc:
                    # column lists of the n_acc nested models in one shot: model k = alive minus the first k + 1 kills
                    # (this runs between the kill kernel and the compaction, i.e. on the device's critical path: one
                    # comparison matrix, no per-model numpy calls)
                    kcols = cand_cols[np.asarray(r['acc'][:n_acc], dtype=np.int64)]
                    when = np.full(len(full), n_acc, dtype=np.int32)       # kill number of every column (never: n_acc)
                    when[kcols] = np.arange(n_acc, dtype=np.int32)
                    member = (when[None, :] > np.arange(n_acc, dtype=np.int32)[:, None]) & alive[None, :]
                    flat_cols = np.nonzero(member)[1].astype(np.int32)
                    ends = np.cumsum(len(model) - 1 - np.arange(n_acc)).tolist()
                    acc_l = np.asarray(r['acc'][:n_acc]).tolist()
                    calls_l = np.asarray(r['calls'][:n_acc]).tolist()
                    ev_l = np.asarray(r['ev'][:n_acc], dtype=np.float64).tolist()
                    lo_ = 0
                    for k_ in range(n_acc):
                        rounds.append(dict(i=int(acc_l[k_]), cols=flat_cols[lo_:ends[k_]],
                                           stream=state['calls'] + int(calls_l[k_]), head=S.get('head'),
                                           gibbs_after=state['gibbs'] + int(calls_l[k_]),
                                           ev_dev=float(ev_l[k_]), icpt_used=state['icpt']))
                        lo_ = ends[k_]
                state['calls'] += r['tested']
                state['gibbs'] += r['tested']
                state['killed'] = state['killed'] + [int(cand_cols[int(i)]) for i in r['acc']]
                if r['n_acc']:
                    state['evmin'] = float(r['ev'][-1])
                    state['cur'] = int(r['acc'][-1]) + 1
            if not rounds:
                return dict(killed=[], evmin=S['ev'], betas=S['betas'], calls=state['calls'], gibbs=state['gibbs'])
            # chains of the accepted models that matter: rounds with threshold-dependent proposals left, + last
            todo = []
            for k_, rd in enumerate(rounds):
                lo_, hi_ = span(k_)
                if 'icpt_true' not in rd and (k_ == len(rounds) - 1 or sens_cum[hi_] > sens_cum[lo_]):
                    todo.append(rd)
            speculated = False
            if todo:
                outlook = dict(killed=list(state['killed']), ev=state['evmin'], calls=state['calls'],
                               gibbs=state['gibbs']) if first_request else None
                answer = yield ('chains', todo, outlook)
                first_request = False
                speculated = answer == 'speculated'
            def first_changed_round():
                """First round whose true chain proposes other candidates in its span than the speculated threshold
                did.  Only the threshold-dependent candidates can differ, and only where `bv0 < threshav * icpt`
                flips; bv0 is sorted ascending, so those are a contiguous index range (NaN sorts last and compares
                false either way)."""
                ks = [k_ for k_, rd in enumerate(rounds) if 'icpt_true' in rd and rd['icpt_true'] != rd['icpt_used']]
                if not ks:
                    return None
                t0 = hy['threshav'] * np.array([rounds[k_]['icpt_used'] for k_ in ks])
                t1 = hy['threshav'] * np.array([rounds[k_]['icpt_true'] for k_ in ks])
                if np.isnan(t0).any() or np.isnan(t1).any():
                    for k_ in ks:
                        lo_, hi_ = span(k_)
                        if not np.array_equal(prop_mask(S, rounds[k_]['icpt_true'])[lo_:hi_],
                                              prop_mask(S, rounds[k_]['icpt_used'])[lo_:hi_]):
                            return k_
                    return None
                lo_ = np.array([rounds[k_]['i'] + 1 for k_ in ks])
                hi_ = np.array([vm if k_ == len(rounds) - 1 else rounds[k_ + 1]['i'] + 1 for k_ in ks])
                a_ = np.maximum(lo_, np.searchsorted(bv0_finite, np.minimum(t0, t1), side='left'))
                b_ = np.maximum(a_, np.minimum(hi_, np.searchsorted(bv0_finite, np.maximum(t0, t1), side='left')))
                changed = np.nonzero(sens_cum[b_] > sens_cum[a_])[0]
                return ks[int(changed[0])] if len(changed) else None

            redo = first_changed_round()
            unrefined = any(rd.get('unrefined') for rd in todo)
            if speculated and (redo is not None or unrefined):
                yield ('rollback',)
                if unrefined:
                    # a Gram-only BIC of the batch was not trustworthy: now that X holds the columns again, redo the
                    # batch with the residual pass and check again
                    for rd in todo:
                        for key in ('icpt_true', 'ev_true', 'rr', 'slot', 'owner', 'unrefined'):
                            rd.pop(key, None)
                    yield ('chains', todo, None)
                    redo = first_changed_round()
            if redo is None:
                last_rd = rounds[-1]
                owner = last_rd.get('owner', -1)          # -1: every rank ran this model's chain itself
                cur_betas = None
                if owner < 0 or owner == engine.rank:
                    cur_betas = last_rd['rr'].betas_of(last_rd['slot']).clone()
                if owner >= 0:
                    # the accepted model's draws live on the rank that ran its chain
                    if owner != engine.rank:
                        cur_betas = torch.empty((D, len(last_rd['cols'])), dtype=torch.float64, device=engine.device)
                    g = engine.group
                    src_rank = engine.dist.get_global_rank(g, owner) if g is not None else owner
                    engine.dist.broadcast(cur_betas, src=src_rank, group=g)
                for rd in rounds:
                    rd.pop('rr', None)
                # report the spectral-path BIC of the accepted model
                return dict(killed=list(state['killed']), evmin=last_rd['ev_true'], betas=cur_betas,
                            calls=state['calls'], gibbs=state['gibbs'])
            # roll back to just after round `redo`, now with its true threshold, and continue from there
            rd = rounds[redo]
            rounds = rounds[:redo + 1]
            rd['icpt_used'] = rd['icpt_true']
            state = dict(killed=sorted(set(full) - set(int(c_) for c_ in rd['cols'])), evmin=rd['ev_true'],
                         cur=rd['i'] + 1, icpt=rd['icpt_true'], calls=rd['stream'], gibbs=rd['gibbs_after'])

    def kill_launch(S, request):
        _, model, pos, icpt_now, evmin_now, start = request
        if S.get('eig') is not None and len(model) == len(S['full']):
            # the loop starts from the full model, whose eigendecomposition is at hand
            return engine.kill_loop_launch(model, pos, S['bv0'], S['bv1'], hyp, hy['threshav'], hy['threshstda'],
                                           hy['threshstdb'], icpt_now, evmin_now, aic_adj, start, eig=S['eig'])
        return engine.kill_loop_launch(model, pos, S['bv0'], S['bv1'], hyp, hy['threshav'], hy['threshstda'],
                                       hy['threshstdb'], icpt_now, evmin_now, aic_adj, start)

    def step_gen(gen, answer=None, first=False):
        """Advance a kill_fast generator: (request, None) or (None, outcome)."""
        try:
            return (next(gen) if first else gen.send(answer)), None
        except StopIteration as stop:
            return None, stop.value

    def drive(gen, request, S, pre=None):
        """Answer a kill_fast generator synchronously until it returns its outcome.  pre: (todo, indices, parts) of a
        chains_launch already made for the pending request."""
        while True:
            if request[0] == 'kill':
                answer = kill_launch(S, request).finish()
            else:
                assert request[0] == 'chains', request[0]
                if pre is not None and request[1] is pre[0]:
                    chains_now(request[1], launched=pre[1:])
                else:
                    chains_now(request[1])
                pre = None
                answer = 'done'
            request, outcome = step_gen(gen, answer)
            if request is None:
                return outcome

    # ---- phase D: drop the accepted kills (FR:1691-1695) and the bookkeeping of FR:1701-1721 ----------------------------
    def same_model_bic(ev, n_killed, S, prev_ev):
        """When EVERY new term of a substage is killed (FR:1673-1692) the surviving model is the previous substage's
        model, and upstream its BIC is the previous value bit for bit (same X, same arithmetic: the reference's traces
        hold exact ties only, never near-ties) -- a non-improvement under the strict `ev < min(evs)` of FR:1702.  Here
        the same model's BIC may come from another route (tableau score, another eigensolver batch, a secular update)
        and differ in the last bits, which would turn the tie into a coin toss: carry the previous value over."""
        if prev_ev is not None and S['vm'] > 0 and n_killed == S['vm']:
            return prev_ev
        return ev

    def close_substage(S, outcome, compacted):
        """Returns True when the fit is finished.  compacted: the engine and `terms` already reflect outcome['killed']
        (the driver speculated on it)."""
        nonlocal terms, last, best, greater
        killed = outcome['killed']
        if compacted:
            terms_s = S['terms_final']
        else:
            terms_s = S['terms']
            if killed:
                kset = set(killed)
                keep = [c for c in S['full'] if c not in kset]
                engine.compact(keep)
                terms_s = np.delete(terms_s, [c - 1 for c in killed], axis=0)
            terms = terms_s
        cnt['calls'], cnt['gibbs'] = outcome['calls'], outcome['gibbs']
        ev = same_model_bic(outcome['evmin'], len(killed), S, evs[-1] if evs else None)
        last = (outcome['betas'], terms_s.copy())
        if console:
            print([S['ind'], float(ev)])
        if on_substage is not None:
            on_substage(S['ind'], ev, terms_s)
        if evs:
            if ev < np.min(evs):
                best = last
                greater = 1
                evs.append(ev)
            elif greater < tolerance:
                greater += 1
                evs.append(ev)
            else:
                evs.append(ev)
                return True
        else:
            greater += 1
            best = last
            evs.append(ev)
        return False

    def will_finish(ev_outlook, pending=None):
        """The stopping rule of FR:1701-1721 applied to a BIC that is not final yet (a prediction only).  pending: the
        predicted BIC of the previous substage when that one has not been closed yet (it is applied first)."""
        evs_s, g = list(evs), greater
        if pending is not None:
            if evs_s:
                if pending < np.min(evs_s):
                    g = 1
                elif g < tolerance:
                    g += 1
                else:
                    return True              # the fit is predicted to end with the previous substage
            else:
                g += 1
            evs_s.append(pending)
        return bool(evs_s) and not (ev_outlook < np.min(evs_s)) and g >= tolerance

    # ---- driver ------------------------------------------------------------------------------------------------------------
    # Sequential form, per substage s:  A(s) append -> B(s) full model -> C(s) kill loop [-> chains of the accepted
    # models -> checks] -> D(s) compact + bookkeeping.  B and the chains of C are latency-bound (one eigensolver + one
    # 2000-draw chain each) and leave the device almost idle, so the pipelined form runs the chains of C(s) *together
    # with* B(s + 1): after the kill loop it assumes the checks will pass (they almost always do), compacts, appends the
    # terms of s + 1, and launches the full model of s + 1 on the main context and the chains of s -- which index the
    # Gram as it was before the compaction -- on the side context.  Every decision is then taken from the true chains
    # exactly as in the sequential form; if a check fails (or the fit turns out to be finished) the speculative work
    # is discarded and the substage's columns are rebuilt.  Philox streams are keyed by the `gibbs` call index, so
    # both forms draw the same variates and give bit-identical fits.
    step = (1, first_partition(1, m, sett))
    S = open_substage(step[0], step[1], terms)
    carry = None            # dict(S=previous substage, gen=its generator, todo, gram=its Gram before compaction, cnt0, ev)
    while True:
        # ---- B(s) [+ chains of s - 1] ----
        mark = engine.mark() if (carry is not None or can_split) else None
        handle = full_launch(S)
        if carry is not None:
            # the side batch's stream waits for what the main stream held at `mark` (K1 + K2 of s, which it must not
            # share the SMs with), not for the full model just enqueued -- whose single cluster is launched first so that
            # the batch's many clusters fill the device around it
            todo_prev, vals, mine = carry['todo'], np.zeros((len(carry['todo']) + 1, 2)), my_share(carry['todo'])
            side = chains_launch(todo_prev, mine, gram=carry['gram'], side=True, after=mark)
        start_next(S, mark)          # after both launches: the build waits for their eigensolver kernels
        full_finish(S, handle)
        # ---- C(s): the kill loop ----
        gen_s = request = outcome_s = None
        if not (literal or S['refined']):
            gen_s = kill_fast(S)
            request, outcome_s = step_gen(gen_s, first=True)
            while request is not None and request[0] == 'kill':
                request, outcome_s = step_gen(gen_s, kill_launch(S, request).finish())
        step = walk_next(S['ind'], S['part'])
        # ---- D(s) + A(s + 1), speculatively, BEFORE the host turns to the chains of s - 1: the device goes straight
        # from the kill loop to the compaction and the next substage's columns, and the collection / checks /
        # bookkeeping of s - 1 (read-backs, an allreduce with several ranks) run while those kernels do.  The stopping
        # rule is predicted with the previous substage's predicted BIC, which is almost always its true one.
        spec = None
        if gen_s is not None and request is not None:
            _, todo, outlook = request
            if outlook is not None:
                prev_ev = carry['ev'] if carry is not None else (evs[-1] if evs else None)
                outlook = dict(outlook, ev=same_model_bic(outlook['ev'], len(outlook['killed']), S, prev_ev))
            if pipeline and step is not None and outlook is not None and \
                    (pipeline == 'always' or not will_finish(outlook['ev'], None if carry is None else carry['ev'])):
                # (the chains of s - 1 may still be reading the Gram as it was two compactions ago: Engine.compact
                # rotates three buffers, so this compaction does not write to it)
                gram = engine.gram_state()
                cnt0 = dict(cnt)
                killed = outlook['killed']
                kset = set(killed)
                keep = [c for c in S['full'] if c not in kset]
                engine.compact(keep)
                S['terms_final'] = np.delete(S['terms'], [c - 1 for c in killed], axis=0)
                terms_before = terms
                terms = S['terms_final']
                cnt_spec = dict(calls=outlook['calls'], gibbs=outlook['gibbs'])
                spec = dict(carry=dict(S=S, gen=gen_s, todo=todo, gram=gram, cnt0=cnt0, ev=outlook['ev']),
                            S_next=open_substage(step[0], step[1], terms, p_stable=S['p_old']), cnt=cnt_spec)
                terms = terms_before             # (restored: the bookkeeping of s - 1 below runs in its own state)
        # no speculation (the fit is expected to end here, or this is the last substage of the walk): the chains that
        # verify s are at least ENQUEUED before the host waits for those of s - 1
        pre = None
        if spec is None and carry is not None and gen_s is not None and request is not None and request[0] == 'chains':
            idx_s = my_share(request[1])
            pre = (request[1], idx_s, chains_launch(request[1], idx_s))
        # ---- the chains of s - 1: collect, check, close ----
        if carry is not None:
            need = chains_collect(todo_prev, mine, side, vals, refine=False)
            need = chains_reduce(todo_prev, vals, need)
            for i, rd in enumerate(todo_prev):
                rd['icpt_true'], rd['ev_true'] = float(vals[i, 0]), float(vals[i, 1])
                if need:
                    rd['unrefined'] = True
            prev, gen = carry['S'], carry['gen']
            req_prev, outcome = step_gen(gen, 'speculated')
            if outcome is None:
                # a check failed: rebuild substage s - 1 as it was before the speculation and finish it synchronously
                # (what was started for substage s -- its full model, kill loop, compaction, the columns of s + 1 -- is
                # dropped)
                assert req_prev[0] == 'rollback'
                engine.truncate(prev['p_old'])
                if can_split:
                    engine.append_terms(prev['vecs'], p_stable=prev['p_stable'])
                else:
                    engine.append_terms(prev['vecs'])
                cnt['calls'], cnt['gibbs'] = carry['cnt0']['calls'], carry['cnt0']['gibbs']
                terms = prev['terms']
                req_prev, outcome = step_gen(gen, 'done')
                if outcome is None:
                    outcome = drive(gen, req_prev, prev)
                carry = None
                S = prev
                finished = close_substage(S, outcome, compacted=False)
                step = walk_next(S['ind'], S['part'])
                if finished or step is None:
                    break
                S = open_substage(step[0], step[1], terms, p_stable=S['p_old'])
                continue
            carry = None
            saved = dict(cnt)
            finished = close_substage(prev, outcome, compacted=True)
            cnt.update(calls=saved['calls'], gibbs=saved['gibbs'])      # the speculative substage's counters stand
            if finished:
                engine.truncate(S['p_old'])                              # drop the speculative columns of s (and s + 1)
                cnt['calls'], cnt['gibbs'] = outcome['calls'], outcome['gibbs']
                break
        if spec is not None:
            carry, S, terms = spec['carry'], spec['S_next'], spec['carry']['S']['terms_final']
            cnt['calls'], cnt['gibbs'] = spec['cnt']['calls'], spec['cnt']['gibbs']
            continue
        # ---- C(s) / D(s), synchronous forms ----
        if gen_s is None:
            outcome = kill_literal(S, parity=(mode == _lib.RNG_INJECTED))
        else:
            outcome = outcome_s
            if request is not None:
                outcome = drive(gen_s, request, S, pre=pre)
        finished = close_substage(S, outcome, compacted=False)
        if finished or step is None:
            break
        S = open_substage(step[0], step[1], terms, p_stable=S['p_old'])

    if can_split:
        engine.drop_prefetch()          # a build started for a substage that never opens
    chosen = last if hy['gimmie'] else best
    betas = chosen[0].cpu().numpy()
    return dict(betas=betas, mtx=chosen[1].astype(np.float64), evs=np.array(evs, dtype=np.float64),
                n_gibbs=cnt['gibbs'], n_batches=cnt['batches'])
