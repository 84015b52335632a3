"""CPU oracle for `FoKL.fitupdate` (update=True)  --  TEST INFRASTRUCTURE, NOT PRODUCT CODE.

Only `tests/` may import this module (see the header of fokl_oracle.py for the rule).

Restates `/root/reference/src/FoKL/FoKLRoutines.py` (FR):

    model_prior         FR:1939-1948    mu_old / Sigma_old of the previous fit's draws (`modelBuilder`)
    gibbs_update        FR:2057-2430    the three-case sampler `gibbs_Xin_update` given the finished design matrix
                                        (case 1 FR:2062-2147, case 2 FR:2150-2264, case 3 FR:2267-2426)
    fitupdate           FR:2432-2583    term generation (two-way only, no kill loop), stopping rule, outputs

`form='literal'` follows the reference statement by statement (same numpy / LAPACK calls, draws from the global legacy
numpy RNG): this is the form PINNED against runs of the unmodified reference (oracle/gen_golden.py `update_*`,
tests/test_oracle_golden.py).

`form='eig'` is the algebraically identical arrangement the device kernels use (csrc/update.cu), with the variates of
the literal form injected in the same order:
  * case 1: eigenbasis chain of `fit` + the per-draw log-likelihood of FR:2111-2115 as
    lik = -(n/2) log sig2 - (squerr + sum lam_j (chat_j - gam_j)^2) / (2 sig2);
  * case 3: coordinates gam_o = Q_o' beta_o, gam_n = Q_n' beta_n of the two fixed eigendecompositions
    (FR:2296, 2312); every per-draw `inv` (FR:2365) becomes a diagonal scaling, the cross terms one (po x pn) matrix;
    per draw identical values (1e-9: tests);
  * case 2: the reference factorises `XotXo + Sigma_old^-1 / tausqd` INSIDE the draw loop (FR:2200-2203: 2000 `eigh`
    + 2000 `inv` per call).  One generalised eigendecomposition  Sigma_old^-1 = L L',  L^-1 XotXo L^-T = V D V',
    T = L^-T V  diagonalises every one of those matrices at once (T'(XotXo + c Sigma^-1)T = D + c I), so the draw is
    O(p).  The conditional distributions are IDENTICAL (mean and covariance: tested at every visited state), but the
    square root that maps the p normals to the draw is T (D + c)^-1/2 instead of LAPACK's Q_k Lam_k^-1/2 of that draw,
    so for the same variates the individual draws differ (same law).  Documented departure.
"""
import math

import numpy as np
from numpy import linalg as LA
from scipy.linalg import eigh

import fokl_oracle as fo


def model_prior(betas, burn):
    """FR:1939-1948 (`modelBuilder` for a built model)."""
    mu_old = np.asmatrix(np.mean(betas[burn:-1], axis=0))
    sigma_old = np.cov(betas[burn:-1].transpose())
    return mu_old, sigma_old


def draw_update_variates(case, draws, p_old, p_new, astar, atau_star):
    """Consume the global legacy numpy RNG as one `gibbs_Xin_update` call does: per draw normal(p_old) [cases 2, 3:
    FR:2213, 2355], normal(p_new) [case 3: FR:2367; case 1: its p normals FR:2106], then standard_gamma(astar),
    standard_gamma(atau_star)."""
    zo = np.empty((draws, p_old))
    zn = np.empty((draws, p_new))
    g1 = np.empty(draws)
    g2 = np.empty(draws)
    for k in range(draws):
        if p_old:
            zo[k] = np.random.normal(loc=0, scale=1, size=(p_old, 1))[:, 0]
        if p_new:
            zn[k] = np.random.normal(loc=0, scale=1, size=(p_new, 1))[:, 0]
        g1[k] = np.random.standard_gamma(astar)
        g2[k] = np.random.standard_gamma(atau_star)
    return zo, zn, g1, g2


def update_case(mu_old, mmtx):
    if np.size(mu_old) == 0:
        return 1
    if np.shape(mu_old)[1] == mmtx + 1:
        return 2
    if np.shape(mu_old)[1] < mmtx + 1:
        return 3
    return 4


def gibbs_update(X, data, a, b, atau, btau, draws, sigsqd0, mu_old, Sigma_old, form='literal', variates=None,
                 gram=None, record=None):
    """FR:2057-2430 given the finished design matrix X (n x (mmtx + 1), column 0 = ones).

    gram=(XtX, Xty): parity-harness hook (the device's Gram bits), as in fokl_oracle.gibbs_from_X.
    record (dict, optional): receives per-draw state (sigs, taus, lik, and for case 2 the visited tausqd values).
    Returns (betas, sigs, taus, ev)."""
    mmtx = X.shape[1] - 1
    n = len(data)
    case = update_case(mu_old, mmtx)
    if gram is None:
        XtX_full = np.transpose(X).dot(X)
        Xty_full = np.transpose(X).dot(data)
    else:
        XtX_full = np.array(gram[0], dtype=np.float64)
        Xty_full = np.array(gram[1], dtype=np.float64).reshape(-1, 1)

    # ------------------------------------------------------------------------------------------ case 1 (FR:2062-2147)
    if case == 1:
        tausqd = 1 / sigsqd0
        XtX = XtX_full
        Xty = Xty_full
        Lamb, Q = eigh(XtX)
        Lamb_inv = np.diag(1 / Lamb)
        betahat = Q.dot(Lamb_inv).dot(np.transpose(Q)).dot(Xty)
        squerr = LA.norm(data - X.dot(betahat)) ** 2
        astar = a + 1 + len(data) / 2 + (mmtx + 1) / 2
        atau_star = atau + mmtx / 2
        dtd = np.transpose(data).dot(data)
        betas = np.zeros((draws, mmtx + 1))
        sigs = np.zeros((draws, 1))
        taus = np.zeros((draws, 1))
        sigsqd = sigsqd0
        lik = np.zeros((draws, 1))
        if form == 'literal':
            for k in range(draws):
                Lamb_tausqd = np.diag(Lamb) + (1 / tausqd) * np.identity(mmtx + 1)
                Lamb_tausqd_inv = np.diag(1 / np.diag(Lamb_tausqd))
                mun = Q.dot(Lamb_tausqd_inv).dot(np.transpose(Q)).dot(Xty)
                S = Q.dot(np.diag(np.diag(Lamb_tausqd_inv) ** (1 / 2)))
                vec = np.random.normal(loc=0, scale=1, size=(mmtx + 1, 1))
                betas[k][:] = np.transpose(mun + sigsqd ** (1 / 2) * (S).dot(vec))
                comp1 = -(n / 2) * np.log(sigsqd)
                comp2 = np.transpose(betahat) - betas[k][:]
                comp3 = betahat - np.reshape(betas[k][:], (len(betas[k][:]), 1))
                lik[k] = (comp1 - (squerr + comp2.dot(XtX).dot(comp3)) / (2 * sigsqd)).item()
                vecc = mun - np.reshape(betas[k][:], (len(betas[k][:]), 1))
                comp1 = 0.5 * np.transpose(vecc)
                comp2 = (XtX + (1 / tausqd) * np.identity(mmtx + 1)).dot(vecc)
                comp3 = 0.5 * np.transpose(mun).dot(Xty)
                bstar = b + comp1.dot(comp2) + 0.5 * dtd - comp3
                if bstar < 0:
                    sigsqd = math.nan
                else:
                    sigsqd = (1 / np.random.gamma(astar, 1 / bstar)).item()
                sigs[k] = sigsqd
                btau_star = ((1 / (2 * sigsqd)) * (
                    betas[k][:].dot(np.reshape(betas[k][:], (len(betas[k][:]), 1)))) + btau).item()
                tausqd = 1 / np.random.gamma(atau_star, 1 / btau_star)
                taus[k] = tausqd
        else:
            p = mmtx + 1
            if variates is None:
                variates = draw_update_variates(1, draws, 0, p, astar, atau_star)
            ct = np.transpose(Q).dot(Xty)[:, 0]
            r = spectral_chain(dict(mode=1, po=0, pn=p, draws=draws, b=b, btau=btau, sigsqd0=sigsqd0,
                                    yty=float(np.asarray(dtd).reshape(-1)[0]), squerr=float(squerr), n=n),
                               dict(lam_n=Lamb, c_n=ct), variates)
            sigs[:, 0], taus[:, 0], lik[:, 0] = r['sigs'], r['taus'], r['lik']
            betas = r['gam_n'].dot(np.transpose(Q))
        ev = (mmtx + 1) * np.log(n) - 2 * np.max(lik)
        if record is not None:
            record.update(case=1, sigs=sigs.copy(), taus=taus.copy(), lik=lik.copy())
        return betas, sigs, taus, ev

    mu_old = np.asmatrix(mu_old)
    num_old_terms = np.shape(mu_old)[1]

    # ------------------------------------------------------------------------------------------ case 2 (FR:2150-2264)
    if case == 2:
        length_old = num_old_terms
        tausqd = 1 / sigsqd0
        XotXo = np.asmatrix(XtX_full)
        Xoty = Xty_full
        Sigma_old_inverse = np.linalg.inv(Sigma_old)
        astar = a + len(data) / 2 + (mmtx + 1) / 2
        atau_star = atau + (mmtx + 1) / 2
        yty = np.transpose(data).dot(data)
        ytXo = np.transpose(Xoty) if gram is not None else np.transpose(data).dot(X)
        betas_old = np.asmatrix(np.zeros((draws, num_old_terms)))
        sigs = np.zeros((draws, 1))
        taus = np.zeros((draws, 1))
        sigsqd = sigsqd0
        lik = np.zeros((draws, 1))
        mu_old = mu_old.transpose()
        visited = []
        if form == 'literal':
            for k in range(draws):
                visited.append((float(tausqd), float(sigsqd)))
                Sigma_old_inverse_post = XotXo + (1 / tausqd) * Sigma_old_inverse
                Sigma_old_post = np.linalg.inv(Sigma_old_inverse_post)
                Lamb_old, Q_old = eigh(XotXo + (1 / tausqd) * Sigma_old_inverse)
                Lamb_tausqd_inv_old = 1 / Lamb_old
                mu_old_first_part = (Xoty + (1 / tausqd * Sigma_old_inverse).dot(mu_old))
                mu_old_post = Sigma_old_post.dot(mu_old_first_part)
                S_old = Q_old.dot((np.diag(Lamb_tausqd_inv_old) ** (1 / 2)))
                vec_old = np.random.normal(loc=0, scale=1, size=(length_old, 1))
                betas_old[k][:] = np.transpose(mu_old_post + sigsqd ** (1 / 2) * (S_old).dot(vec_old))
                bo = betas_old[k][:]
                comp1 = 0.5 * (yty - ytXo.dot(bo.transpose()))
                comp2 = 0.5 * (-(bo).dot(Xoty) + (bo).dot(XotXo).dot(bo.transpose()))
                comp3 = 0.5 * (1 / tausqd) * ((bo).dot(Sigma_old_inverse).dot(bo.transpose()) - (bo).dot(
                    Sigma_old_inverse).dot(mu_old))
                comp4 = 0.5 * (1 / tausqd) * (-np.transpose(mu_old).dot(Sigma_old_inverse).dot(
                    bo.transpose()) + np.transpose(mu_old).dot(Sigma_old_inverse).dot(mu_old))
                bstar = comp1 + comp2 + comp3 + comp4 + b
                if bstar < 0:
                    sigsqd = math.nan
                else:
                    sigsqd = (1 / np.random.gamma(astar, 1 / bstar)).item()
                sigs[k] = sigsqd
                comp1 = 0.5 * (1 / sigsqd) * ((bo).dot(Sigma_old_inverse).dot(bo.transpose()) - (bo).dot(
                    Sigma_old_inverse).dot(mu_old))
                comp2 = 0.5 * (1 / sigsqd) * (-np.transpose(mu_old).dot(Sigma_old_inverse).dot(
                    bo.transpose()) + np.transpose(mu_old).dot(Sigma_old_inverse).dot(mu_old))
                btau_star = (comp1 + comp2 + btau).item()
                tausqd = 1 / np.random.gamma(atau_star, 1 / btau_star)
                taus[k] = tausqd
                comp1 = -(n / 2) * np.log(sigsqd)
                comp2 = yty - ytXo.dot(bo.transpose())
                comp3 = -(bo).dot(Xoty) + (bo).dot(XotXo).dot(bo.transpose())
                lik[k] = (comp1 - 0.5 / sigsqd * (comp2 + comp3)).item()
            betas = betas_old
        else:
            p = num_old_terms
            if variates is None:
                variates = draw_update_variates(2, draws, p, 0, astar, atau_star)
            geig = generalised_eig(np.asarray(XotXo), Sigma_old_inverse)
            T, Dg = geig['T'], geig['D']
            r = spectral_chain(dict(mode=2, po=p, pn=0, draws=draws, b=b, btau=btau, sigsqd0=sigsqd0,
                                    yty=float(np.asarray(yty).reshape(-1)[0]), squerr=0.0, n=n),
                               dict(lam_o=Dg, c_o=T.T.dot(np.asarray(Xoty))[:, 0],
                                    m_o=geig['Tinv'].dot(np.asarray(mu_old))[:, 0]), variates)
            sigs[:, 0], taus[:, 0], lik[:, 0] = r['sigs'], r['taus'], r['lik']
            visited = r['visited']
            betas = np.asmatrix(r['gam_o'].dot(T.T))
        ev = (mmtx + 1) * np.log(n) - 2 * max(lik)
        if record is not None:
            record.update(case=2, sigs=sigs.copy(), taus=taus.copy(), lik=lik.copy(), visited=visited)
        return betas, sigs, taus, ev

    # ------------------------------------------------------------------------------------------ case 3 (FR:2267-2426)
    if case == 3:
        length_old = num_old_terms
        length_new = mmtx - num_old_terms + 1
        tausqd = 1 / sigsqd0
        if gram is None:
            # the reference forms every block by its own product (FR:2285-2318): same values as slices of X'X up to the
            # last bit, and the last bit matters to this chain (see tests/test_oracle_golden.py)
            X_old = X[:, 0:length_old]
            X_new = X[:, length_old: length_old + length_new]
            XotXo = np.asmatrix(np.transpose(X_old).dot(X_old))
            Xoty = np.transpose(X_old).dot(data)
        else:
            XotXo = np.asmatrix(XtX_full[0:length_old, 0:length_old])
            Xoty = Xty_full[0:length_old]
        Sigma_old_inverse = np.linalg.inv(Sigma_old)
        Sigma_old_inverse_post = XotXo + Sigma_old_inverse
        Sigma_old_post = np.linalg.inv(Sigma_old_inverse_post)
        Lamb_old, Q_old = eigh(XotXo + Sigma_old_inverse)
        if gram is None:
            XntXn = np.transpose(X_new).dot(X_new)
            Xnty = np.transpose(X_new).dot(data)
        else:
            XntXn = XtX_full[length_old:length_old + length_new, length_old:length_old + length_new]
            Xnty = Xty_full[length_old:length_old + length_new]
        Lamb_new, Q_new = eigh(XntXn)
        if gram is None:
            XotXn = np.transpose(X_old).dot(X_new)
            XntXo = np.transpose(X_new).dot(X_old)
        else:
            XotXn = XtX_full[0:length_old, length_old:length_old + length_new]
            XntXo = np.transpose(XotXn)
        Lamb_tausqd_inv_old = np.diag(np.linalg.inv((np.diag(Lamb_old))))
        astar = a + len(data) / 2 + (mmtx + 1) / 2
        atau_star = atau + (length_new) / 2
        yty = np.transpose(data).dot(data)
        ytXo = np.transpose(Xoty)
        ytXn = np.transpose(Xnty)
        betas_old = np.asmatrix(np.zeros((draws, num_old_terms)))
        betas_new = np.asmatrix(np.zeros((draws, mmtx - num_old_terms + 1)))
        sigs = np.zeros((draws, 1))
        taus = np.zeros((draws, 1))
        sigsqd = sigsqd0
        lik = np.zeros((draws, 1))
        mu_old = mu_old.transpose()
        if form == 'literal':
            for k in range(draws):
                mu_old_first_part = Xoty - XotXn.dot(betas_new[k - 1].transpose()) + Sigma_old_inverse.dot(mu_old)
                mu_old_post = Sigma_old_post.dot(mu_old_first_part)
                S_old = Q_old.dot((np.diag(Lamb_tausqd_inv_old) ** (1 / 2)))
                vec_old = np.random.normal(loc=0, scale=1, size=(length_old, 1))
                betas_old[k][:] = np.transpose(mu_old_post + sigsqd ** (1 / 2) * (S_old).dot(vec_old))
                Lamb_tausqd_new = np.diag(Lamb_new) + (1 / tausqd) * np.identity(length_new)
                Lamb_tausqd_inv_new = np.diag(np.linalg.inv(Lamb_tausqd_new))
                mu_new_first_part = Xnty - XntXo.dot(betas_old[k].transpose())
                Sigma_new_inverse_post = XntXn + (1 / tausqd) * np.identity(length_new)
                mu_new_post = (np.linalg.inv(Sigma_new_inverse_post)).dot(mu_new_first_part)
                S_new = Q_new.dot((np.diag(Lamb_tausqd_inv_new) ** (1 / 2)))
                vec_new = np.random.normal(loc=0, scale=1, size=(length_new, 1))
                betas_new[k][:] = np.transpose(mu_new_post + sigsqd ** (1 / 2) * (S_new).dot(vec_new))
                bo = betas_old[k][:]
                bn = betas_new[k][:]
                comp1 = 0.5 * (yty - ytXo.dot(bo.transpose()) - ytXn.dot(bn.transpose()))
                comp2 = 0.5 * (-(bo).dot(Xoty) + (bo).dot(XotXo).dot(bo.transpose()) + (bo).dot(XotXn).dot(
                    bn.transpose()))
                comp3 = 0.5 * (-(bn).dot(Xnty) + (bn).dot(XntXo).dot(bo.transpose()) + (bn).dot(XntXn).dot(
                    bn.transpose()))
                comp4 = 0.5 / tausqd * ((bn).dot(bn.transpose()))
                comp5 = 0.5 * ((bo).dot(Sigma_old_inverse).dot(bo.transpose()) - (bo).dot(Sigma_old_inverse).dot(
                    mu_old))
                comp6 = 0.5 * (-np.transpose(mu_old).dot(Sigma_old_inverse).dot(bo.transpose()) + np.transpose(
                    mu_old).dot(Sigma_old_inverse).dot(mu_old))
                bstar = comp1 + comp2 + comp3 + comp4 + comp5 + comp6 + b
                if bstar < 0:
                    sigsqd = math.nan
                else:
                    sigsqd = (1 / np.random.gamma(astar, 1 / bstar)).item()
                sigs[k] = sigsqd
                btau_star = ((1 / (2 * sigsqd)) * (bn.dot(bn.transpose())) + btau).item()
                tausqd = 1 / np.random.gamma(atau_star, 1 / btau_star)
                taus[k] = tausqd
                comp1 = -(n / 2) * np.log(sigsqd)
                comp2 = yty - ytXo.dot(bo.transpose()) - ytXn.dot(bn.transpose())
                comp3 = -(bo).dot(Xoty) + (bo).dot(XotXo).dot(bo.transpose()) + (bo).dot(XotXn).dot(bn.transpose())
                comp4 = -(bn).dot(Xnty) + (bn).dot(XntXo).dot(bo.transpose()) + (bn).dot(XntXn).dot(bn.transpose())
                lik[k] = (comp1 - 0.5 / sigsqd * (comp2 + comp3 + comp4)).item()
            betas = np.concatenate((betas_old, betas_new), axis=1)
        else:
            po, pn = length_old, length_new
            if variates is None:
                variates = draw_update_variates(3, draws, po, pn, astar, atau_star)
            pre = case3_precompute(np.asarray(XotXo), np.asarray(XotXn), np.asarray(XntXn), np.asarray(Xoty)[:, 0],
                                   np.asarray(Xnty)[:, 0], Sigma_old_inverse, np.asarray(mu_old)[:, 0],
                                   Lamb_old, Q_old, Lamb_new, Q_new)
            r = spectral_chain(dict(mode=3, po=po, pn=pn, draws=draws, b=b, btau=btau, sigsqd0=sigsqd0,
                                    yty=float(np.asarray(yty).reshape(-1)[0]), squerr=0.0, n=n),
                               dict(lam_o=Lamb_old, c_o=pre['co'], t_o=pre['to'], m_o=pre['mo'], lam_n=Lamb_new,
                                    c_n=pre['cn'], M=pre['M'], K=pre['K'], W=pre['W']), variates)
            sigs[:, 0], taus[:, 0], lik[:, 0] = r['sigs'], r['taus'], r['lik']
            betas = np.asmatrix(np.concatenate((r['gam_o'].dot(Q_old.T), r['gam_n'].dot(Q_new.T)), axis=1))
        ev = (mmtx + 1) * np.log(n) - 2 * max(lik)
        if record is not None:
            record.update(case=3, sigs=sigs.copy(), taus=taus.copy(), lik=lik.copy())
        return betas, sigs, taus, ev

    print('Error: No appropriate cases for evaluation found.')
    return None


def spectral_chain(spec, arrays, variates):
    """numpy statement of the draw loop the device runs (csrc/update_math.cuh `update_chain`), for the kernel tests.
    spec: mode, po, pn, draws, b, btau, sigsqd0, yty, squerr, n.  arrays: the keys of fokl_update_chain (M is po x pn).
    variates: (z_o draws x po, z_n draws x pn, g1, g2).  Returns dict(gam_o, gam_n, sigs, taus, lik, visited)."""
    mode, po, pn, D = spec['mode'], spec['po'], spec['pn'], spec['draws']
    zo, zn, g1, g2 = variates
    b, btau, n, yty = spec['b'], spec['btau'], spec['n'], spec['yty']
    go = np.zeros((D, po))
    gn = np.zeros((D, pn))
    sigs, taus, lik = np.zeros(D), np.zeros(D), np.zeros(D)
    sig = float(spec['sigsqd0'])
    itau = float(spec['sigsqd0'])           # 1 / tausqd0, tausqd0 = 1 / sigsqd0
    visited = []
    gn_prev = np.zeros(pn)
    A = arrays
    for k in range(D):
        visited.append((1 / itau, sig))
        if mode == 1:
            lam, ct = A['lam_n'], A['c_n']
            d = 1 / (lam + itau)
            g = d * ct + sig ** (1 / 2) * (d ** (1 / 2)) * zn[k]
            gn[k] = g
            lik[k] = -(n / 2) * np.log(sig) - (spec['squerr'] + np.sum(lam * (ct / lam - g) ** 2)) / (2 * sig)
            s3 = np.sum(g * g)
            bstar = b + 0.5 * (np.sum(lam * g * g) - 2 * np.sum(g * ct) + yty + s3 * itau)
            sig = math.nan if bstar < 0 else 1 / ((1 / bstar) * g1[k])
            btau_star = (1 / (2 * sig)) * s3 + btau
        elif mode == 2:
            Dg, c1, mc = A['lam_o'], A['c_o'], A['m_o']
            d = 1 / (Dg + itau)
            g = d * (c1 + itau * mc) + sig ** (1 / 2) * (d ** (1 / 2)) * zo[k]
            go[k] = g
            sse = yty - 2 * np.sum(g * c1) + np.sum(Dg * g * g)
            dev = np.sum((g - mc) ** 2)
            bstar = 0.5 * sse + 0.5 * itau * dev + b
            sig = math.nan if bstar < 0 else 1 / ((1 / bstar) * g1[k])
            btau_star = 0.5 * (1 / sig) * dev + btau
            lik[k] = -(n / 2) * np.log(sig) - 0.5 / sig * sse
        else:
            lo, ln_ = A['lam_o'], A['lam_n']
            g_o = (A['c_o'] - A['M'].dot(gn_prev)) / lo + sig ** (1 / 2) * (lo ** (-1 / 2)) * zo[k]
            dn = 1 / (ln_ + itau)
            g_n = dn * (A['c_n'] - A['M'].T.dot(g_o)) + sig ** (1 / 2) * (dn ** (1 / 2)) * zn[k]
            go[k] = g_o
            gn[k] = g_n
            gn_prev = g_n
            sse = yty - 2 * (np.sum(g_o * A['t_o']) + np.sum(g_n * A['c_n'])) + g_o.dot(A['K'].dot(g_o)) + \
                2 * g_o.dot(A['M'].dot(g_n)) + np.sum(ln_ * g_n * g_n)
            dv = g_o - A['m_o']
            nn = np.sum(g_n * g_n)
            bstar = 0.5 * sse + 0.5 * itau * nn + 0.5 * dv.dot(A['W'].dot(dv)) + b
            sig = math.nan if bstar < 0 else 1 / ((1 / bstar) * g1[k])
            btau_star = (1 / (2 * sig)) * nn + btau
            lik[k] = -(n / 2) * np.log(sig) - 0.5 / sig * sse
        sigs[k] = sig
        tau = 1 / ((1 / btau_star) * g2[k])
        itau = 1 / tau
        taus[k] = tau
    return dict(gam_o=go, gam_n=gn, sigs=sigs, taus=taus, lik=lik, visited=visited)


def generalised_eig(G, Sinv):
    """Sinv = L L' (Cholesky), L^-1 G L^-T = V D V'  ->  T = L^-T V with  T' G T = D,  T' Sinv T = I."""
    L = np.linalg.cholesky(Sinv)
    Li = np.linalg.inv(L)
    C = Li.dot(G).dot(Li.T)
    C = 0.5 * (C + C.T)
    D, V = eigh(C)
    T = Li.T.dot(V)
    Tinv = V.T.dot(L.T)
    return dict(L=L, C=C, D=D, V=V, T=T, Tinv=Tinv)


def case3_precompute(Goo, Gon, Gnn, Xoty, Xnty, Sinv, mu, Lamb_old, Q_old, Lamb_new, Q_new):
    """The fixed quantities of the case-3 chain in the coordinates gam_o = Q_o' beta_o, gam_n = Q_n' beta_n."""
    h = Sinv.dot(mu)
    return dict(M=Q_old.T.dot(Gon).dot(Q_new), K=Q_old.T.dot(Goo).dot(Q_old), W=Q_old.T.dot(Sinv).dot(Q_old),
                co=Q_old.T.dot(Xoty + h), to=Q_old.T.dot(Xoty), cn=Q_new.T.dot(Xnty), mo=Q_old.T.dot(mu))


def update_vecs(ind, i, m):
    """FR:2496-2500: np.unique(perms([ind - i, i, 0, ...]), axis=0)."""
    v = np.zeros(m)
    v[0] = ind - i
    v[1] = i
    return fo.distinct_perms(v)


def update_i_list(ind):
    """FR:2484-2489."""
    if ind == 1:
        return [0]
    i_list = np.arange(0, math.floor(ind / 2) + 0.1, 1)
    return i_list[::-1]


def fitupdate(inputs, data, phis, kernel=fo.CUBIC, a=4, b=None, atau=4, btau=None, tolerance=3, draws=2000,
              gimmie=False, aic=False, sigsqd0=0.5, prior=None, form='literal', on_gibbs=None, gram_hook=None,
              threads=1):
    """FR:2432-2583 on normalised `inputs` (n x m, m >= 2) and `data` (n x 1); relats_in = [] only.
    `draws` = burnin + draws (FR:1932).  prior = None (model not built) or (mu_old, sigma_old) from model_prior.
    Returns dict(betas, mtx, evs, built, n_gibbs): `built` is True only when the tolerance rule ended the loop
    (FR:2563-2566)."""
    inputs = np.asarray(inputs, dtype=np.float64)
    data = np.asarray(data, dtype=np.float64).reshape(-1, 1)
    n, m = inputs.shape
    if m < 2:
        raise ValueError("not enough values to unpack (expected 2, got 0)")       # FR:2529 for a single input
    if prior is None:
        mu_old, sigma_old = [], []
        num_old_terms = 0
    else:
        mu_old, sigma_old = prior
        num_old_terms = np.shape(mu_old)[1]
    table = fo.phis_to_table(phis, kernel)
    damtx = np.zeros((0, m))
    evs = []
    X = np.ones((n, 1))
    ind = 1
    greater = 0
    finished = 0
    built = False
    calls = 0
    betas = betas_best = mtx = None
    while True:
        for i in update_i_list(ind):
            vecs = update_vecs(ind, i, m)
            damtx = np.concatenate((damtx, vecs), axis=0)
            length = damtx.shape[0]
            if num_old_terms - 1 <= length:
                # FR:2008-2055: columns nxin .. mmtx of X are built (X carries over between calls)
                have = X.shape[1] - 1
                if length > have:
                    X = np.append(X, fo.basis_columns(inputs, damtx[have:length], phis, kernel, table=table,
                                                      threads=threads), axis=1)
                gram = gram_hook(damtx) if gram_hook is not None else None
                rec = {} if on_gibbs is not None else None
                betas, sigs, taus, ev = gibbs_update(X, data, a, b, atau, btau, draws, sigsqd0, mu_old, sigma_old,
                                                     form=form, gram=gram, record=rec)
                calls += 1
                if aic:
                    ev = ev + (2 - np.log(n)) * length
                if on_gibbs is not None:
                    on_gibbs(dict(call=calls, discmtx=damtx.copy(), betas=betas, ev=ev, **rec))
                if np.size(evs) == 0:
                    evs = [ev]
                else:
                    evs = np.concatenate((evs, [ev]))
                if ev == np.min(evs):
                    betas_best = betas
                    mtx = damtx
                    greater = 1
                elif greater <= tolerance:
                    greater = greater + 1
                else:
                    finished = 1
                    built = True
                    break
        if finished != 0:
            break
        ind = ind + 1
        if ind > len(phis):
            break
    if gimmie:
        betas_best = betas
        mtx = damtx
    return dict(betas=betas_best, mtx=mtx, evs=evs, built=built, n_gibbs=calls)
