"""Host side of the forward-selection loop of `FoKL.fit` (reference: src/FoKL/FoKLRoutines.py:1561-1760).

The loop is inherently sequential in its stages; what it asks of the device per substage is
  (1) append the new terms' columns to X and extend the Gram (K1 + K2),
  (2) evaluate the full model (eig + BIC + Gibbs chain + column statistics),
  (3) evaluate kill proposals -- here as *batches* of independent candidate models (one CTA each) instead
      of one `gibbs` call at a time -- and
  (4) drop the accepted kills (column / Gram compaction).
The batching is exact: a proposal's outcome depends only on the kill set accepted so far and on the
draws of the last accepted model (FR:1670-1690), so all proposals that pass the threshold test under the
current state are evaluated together, the first (in the reference's order) that lowers the BIC is
accepted, and the remainder is re-batched under the new state.
"""
import math

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
                   recorder=None):
    """Run the selection loop on an Engine that has a dataset bound (engine.begin_fit).

    hy: dict with a, b, atau, btau, tolerance, total_draws, gimmie, way3, threshav, threshstda, threshstdb, aic.
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

    terms = np.zeros((0, m), dtype=np.int64)     # damtx: row j <-> column j + 1 of X
    evs = []
    best = None            # (betas tensor, mtx)
    last = None
    greater = 0
    call_id = [0]
    n_gibbs = 0
    n_batches = 0

    def run(col_sets, chains, ids, refine=True):
        nonlocal n_batches
        n_batches += 1
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

    ind = 1
    finished = False
    while True:
        part = first_partition(ind, m, sett)
        while True:
            vecs = distinct_permutations(part)
            vm = vecs.shape[0]
            p_old = engine.P                     # columns of the model accepted so far (incl. intercept)
            engine.append_terms(vecs)
            terms = np.concatenate([terms, vecs], axis=0)
            dam = terms.shape[0]
            full = list(range(engine.P))

            # ---- full model (FR:1650) -----------------------------------------------------------------------
            call_id[0] += 1
            n_gibbs += 1
            chain_arg = src.draw(len(full)) if mode == _lib.RNG_INJECTED else True
            res = run([full], [chain_arg], [call_id[0]])
            ev = float(res.ev[0]) + aic_adj * (dam + 1)
            st = res.stats_of(0).cpu().numpy()
            cur_betas = res.betas_of(0)
            new_cols = np.arange(p_old, p_old + vm)
            with np.errstate(all='ignore'):
                bv0 = np.abs(st[0, new_cols])                          # |mean| over rows h1.. (FR:1656)
                bv1 = st[1, new_cols] / np.abs(st[2, new_cols])        # std / |mean over rows h0..| (FR:1657-58)
            order = np.argsort(bv0, kind='quicksort')
            bv0, bv1, cand_cols = bv0[order], bv1[order], new_cols[order]
            icpt = abs(float(st[2, 0]))                                # |mean(beters[h0:, 0])| (FR:1671)

            # ---- kill proposals (FR:1666-1692) -------------------------------------------------------------------
            killed = []            # column indices (into current X) accepted for removal
            evmin = ev
            cur = 0

            with np.errstate(invalid='ignore'):
                always = bv1 > hy['threshstdb']                       # proposed whatever the intercept is
                maybe = bv1 > hy['threshstda']

            def proposals(start, icpt_now):
                with np.errstate(invalid='ignore'):
                    mask = always | (maybe & (bv0 < hy['threshav'] * icpt_now))
                return (np.nonzero(mask[start:])[0] + start).tolist()

            if mode == _lib.RNG_INJECTED or eager:
                # literal order: one `gibbs`-equivalent evaluation (eig + chain) per proposal.  In parity mode the
                # numpy stream is consumed test by test; in eager mode all proposals of a batch run side by side.
                while cur < vm:
                    props = proposals(cur, icpt)
                    if not props:
                        break
                    sets = []
                    for i in props:
                        drop = set(killed) | {int(cand_cols[i])}
                        sets.append([c for c in full if c not in drop])
                    ids = [call_id[0] + r + 1 for r in range(len(props))]
                    accepted = None
                    if mode == _lib.RNG_INJECTED:
                        for r, i in enumerate(props):
                            call_id[0] += 1
                            n_gibbs += 1
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
                        call_id[0] += tested
                        n_gibbs += tested
                    if accepted is None:
                        break
                    i, evt, rr, slot = accepted
                    killed.append(int(cand_cols[i]))
                    evmin = evt
                    cur_betas = rr.betas_of(slot).clone()
                    icpt = abs(float(rr.stats_of(slot)[2, 0].item()))
                    cur = i + 1
            else:
                # fast path: the whole kill loop runs on the device in one launch (fokl_kill_loop: sweep-operator
                # inverse of the model's Gram, one O(p^2) reverse sweep per accepted kill).  The chains of the accepted
                # models -- which only feed later rounds' threshold through |mean intercept| (FR:1671) and the final
                # draws (FR:1690) -- are run afterwards as one batch, only for the rounds whose remaining proposals can
                # depend on that threshold, and every such round is re-checked against its true chain (the loop is
                # re-run from the first round that differs), so the outcome is exactly that of the sequential loop.
                # Round k's chain matters only through the threshold it sets for the candidates examined before the next
                # acceptance, i.e. indices (i_k, i_{k+1}] of the sorted list (to the end for the last round): if none of
                # them is threshold-dependent (threshstda < bv1 <= threshstdb) the chain of round k is never looked at.
                with np.errstate(invalid='ignore'):
                    bv1_sens = (bv1 > hy['threshstda']) & ~(bv1 > hy['threshstdb'])
                sens_cum = np.concatenate([[0], np.cumsum(bv1_sens)])      # threshold-dependent candidates in [0, j)

                def span(k_):
                    return rounds[k_]['i'] + 1, (vm if k_ == len(rounds) - 1 else rounds[k_ + 1]['i'] + 1)

                def prop_mask(icpt_now):
                    with np.errstate(invalid='ignore'):
                        return always | (maybe & (bv0 < hy['threshav'] * icpt_now))
                state = dict(killed=[], evmin=evmin, cur=0, icpt=icpt, calls=call_id[0], gibbs=n_gibbs)
                rounds = []        # accepted kills: dict(i, cols, stream, icpt_used, gibbs_after [, icpt_true, ev_true])
                fallback = False
                while True:
                    if state['cur'] < vm and proposals(state['cur'], state['icpt']):
                        # column bookkeeping on numpy masks (the lists are up to 160 models x 220 columns per substage)
                        alive = np.ones(len(full), dtype=bool)
                        alive[np.asarray(state['killed'], dtype=np.int64)] = False
                        model = np.nonzero(alive)[0].astype(np.int32)
                        where = np.ones(len(full), dtype=np.int32)            # killed candidates: any valid position
                        where[model] = np.arange(len(model), dtype=np.int32)
                        pos = where[cand_cols]
                        r = engine.kill_loop(model, pos, bv0, bv1, hyp, hy['threshav'], hy['threshstda'],
                                             hy['threshstdb'], state['icpt'], state['evmin'], aic_adj, state['cur'])
                        n_batches += 1
                        if r['bad']:
                            fallback = True
                            break
                        for k_ in range(r['n_acc']):
                            i = int(r['acc'][k_])
                            alive[int(cand_cols[i])] = False
                            cols_k = np.nonzero(alive)[0].astype(np.int32)
                            rounds.append(dict(i=i, cols=cols_k, stream=state['calls'] + int(r['calls'][k_]),
                                               gibbs_after=state['gibbs'] + int(r['calls'][k_]),
                                               ev_dev=float(r['ev'][k_]), icpt_used=state['icpt']))
                        state['calls'] += r['tested']
                        state['gibbs'] += r['tested']
                        state['killed'] = state['killed'] + [int(cand_cols[int(i)]) for i in r['acc']]
                        if r['n_acc']:
                            state['evmin'] = float(r['ev'][-1])
                            state['cur'] = int(r['acc'][-1]) + 1
                    if not rounds:
                        break
                    # chains of the accepted models that matter: rounds with threshold-dependent proposals left, + last
                    todo = []
                    for k_, rd in enumerate(rounds):
                        lo_, hi_ = span(k_)
                        if 'icpt_true' not in rd and (k_ == len(rounds) - 1 or sens_cum[hi_] > sens_cum[lo_]):
                            todo.append(rd)
                    if todo:
                        # The models of a batch are independent: with several ranks each evaluates every world-th one
                        # (all ranks hold the full Gram) and the two scalars per model that drive the loop are summed
                        # into place by one allreduce, so every rank takes the same decisions.
                        world = engine.world if engine.dist is not None else 1
                        mine = list(range(engine.rank, len(todo), world)) if world > 1 else list(range(len(todo)))
                        vals = np.zeros((len(todo) + 1, 2))

                        def chains_of(which, refine):
                            rr = run([todo[i]['cols'] for i in which], [True] * len(which),
                                     [todo[i]['stream'] for i in which], refine=refine)
                            stats_h = rr.stats.cpu().numpy()      # one read-back for the whole batch
                            for slot, i in enumerate(which):
                                # mean(beters[h0:, 0]) of this model: row 2, column 0 of its 3 x p statistics block
                                vals[i, 0] = abs(float(stats_h[3 * rr.vec_off[slot] + 2 * rr.p[slot]]))
                                vals[i, 1] = float(rr.ev[slot]) + aic_adj * len(todo[i]['cols'])
                                todo[i]['rr'], todo[i]['slot'] = rr, slot
                            return rr

                        if world > 1:
                            if mine:
                                rr_ = chains_of(mine, refine=False)
                                vals[len(todo), 0] = float(np.any(engine.refine_mask(rr_.ev, rr_.p)))
                            t = torch.from_numpy(vals).to(engine.device)
                            engine._allreduce(t)
                            vals = t.cpu().numpy()
                            for i, rd in enumerate(todo):
                                rd['owner'] = i % world
                            if vals[len(todo), 0] > 0:
                                # some model needs the N-length residual pass (a collective): evaluate the batch replicated
                                vals = np.zeros((len(todo) + 1, 2))
                                chains_of(list(range(len(todo))), refine=True)
                                for rd in todo:
                                    rd['owner'] = -1
                        else:
                            chains_of(mine, refine=True)
                        for i, rd in enumerate(todo):
                            rd['icpt_true'], rd['ev_true'] = float(vals[i, 0]), float(vals[i, 1])
                    redo = None
                    for k_, rd in enumerate(rounds):
                        if 'icpt_true' not in rd or rd['icpt_true'] == rd['icpt_used']:
                            continue
                        lo_, hi_ = span(k_)
                        if not np.array_equal(prop_mask(rd['icpt_true'])[lo_:hi_], prop_mask(rd['icpt_used'])[lo_:hi_]):
                            redo = k_
                            break
                    if redo is None:
                        last_rd = rounds[-1]
                        owner = last_rd.get('owner', -1)          # -1: every rank ran this model's chain itself
                        if owner < 0 or owner == engine.rank:
                            cur_betas = last_rd['rr'].betas_of(last_rd['slot']).clone()
                        if owner >= 0:
                            # the accepted model's draws live on the rank that ran its chain
                            if owner != engine.rank:
                                cur_betas = torch.empty((D, len(last_rd['cols'])), dtype=torch.float64,
                                                        device=engine.device)
                            g = engine.group
                            src = engine.dist.get_global_rank(g, owner) if g is not None else owner
                            engine.dist.broadcast(cur_betas, src=src, group=g)
                        icpt = last_rd['icpt_true']
                        evmin = last_rd['ev_true']        # report the spectral-path BIC of the accepted model
                        killed = list(state['killed'])
                        for rd in rounds:
                            rd.pop('rr', None)
                        break
                    # roll back to just after round `redo`, now with its true threshold, and continue from there
                    rd = rounds[redo]
                    rounds = rounds[:redo + 1]
                    rd['icpt_used'] = rd['icpt_true']
                    state = dict(killed=sorted(set(full) - set(int(c_) for c_ in rd['cols'])), evmin=rd['ev_true'],
                                 cur=rd['i'] + 1, icpt=rd['icpt_true'], calls=rd['stream'], gibbs=rd['gibbs_after'])
                if fallback:
                    # Gram not numerically positive definite (p >= N regimes): literal loop, one spectral evaluation
                    # (eig + chain) per proposal, proposals of a round side by side
                    killed, evmin, cur = [], ev, 0
                    while cur < vm:
                        props = proposals(cur, icpt)
                        if not props:
                            break
                        sets = []
                        for i in props:
                            drop = set(killed) | {int(cand_cols[i])}
                            sets.append([c for c in full if c not in drop])
                        ids = [call_id[0] + r_ + 1 for r_ in range(len(props))]
                        rr = run(sets, [True] * len(sets), ids)
                        accepted = None
                        for r_, i in enumerate(props):
                            evt = float(rr.ev[r_]) + aic_adj * len(sets[r_])
                            if evt < evmin:
                                accepted = (i, evt, rr, r_)
                                break
                        tested = (props.index(accepted[0]) + 1) if accepted else len(props)
                        call_id[0] += tested
                        n_gibbs += tested
                        if accepted is None:
                            break
                        i, evt, rr, slot = accepted
                        killed.append(int(cand_cols[i]))
                        evmin = evt
                        cur_betas = rr.betas_of(slot).clone()
                        icpt = abs(float(rr.stats_of(slot)[2, 0].item()))
                        cur = i + 1
                else:
                    call_id[0] = state['calls']
                    n_gibbs = state['gibbs']
                    if not rounds:
                        killed = []

            # ---- drop accepted kills (FR:1691-1695) ---------------------------------------------------------------
            if killed:
                keep = [c for c in full if c not in set(killed)]
                engine.compact(keep)
                terms = np.delete(terms, [c - 1 for c in killed], axis=0)
            ev = evmin
            last = (cur_betas, terms.copy())
            if console:
                print([ind, float(ev)])
            if on_substage is not None:
                on_substage(ind, ev, terms)

            # ---- bookkeeping (FR:1701-1721) -----------------------------------------------------------------------
            if evs:
                if ev < np.min(evs):
                    best = last
                    greater = 1
                    evs.append(ev)
                elif greater < tolerance:
                    greater += 1
                    evs.append(ev)
                else:
                    finished = True
                    evs.append(ev)
                    break
            else:
                greater += 1
                best = last
                evs.append(ev)
            if not next_partition(part, m, hy['way3']):
                break
        if finished:
            break
        ind += 1
        if ind > n_phis:
            break

    chosen = last if hy['gimmie'] else best
    betas = chosen[0].cpu().numpy()
    return dict(betas=betas, mtx=chosen[1].astype(np.float64), evs=np.array(evs, dtype=np.float64),
                n_gibbs=n_gibbs, n_batches=n_batches)
