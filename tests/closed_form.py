"""Closed-form likelihoods that nobody in this repository computed: textbook transition
probabilities of JC69 (Jukes & Cantor 1969), F81 (Felsenstein 1981) and HKY85 (Hasegawa, Kishino
& Yano 1985) written out as formulas, Yang's (1994) mean discrete-Gamma rates from scipy's
incomplete Gamma function, and star-tree site likelihoods as explicit sums over the centre state.
They pin oracle/lik_oracle.c (tests/test_oracle_cpu.py) and the CUDA engine
(tests/test_closed_form_gpu.py) to numbers independent of either's eigendecomposition, pruning
order and scaling.  States are ordered A, C, G, T; exchangeabilities AC, AG, AT, CG, CT, GT."""
import numpy as np


def gamma_mean_rates(alpha, ncat):
    """Mean of each of `ncat` equal-probability slices of Gamma(alpha, rate alpha) (Yang 1994, eq. 10)."""
    from scipy.special import gammainc
    from scipy.stats import gamma
    if ncat == 1:
        return np.ones(1)
    cuts = gamma.ppf(np.arange(1, ncat) / ncat, alpha, scale=1.0 / alpha)
    upper = np.concatenate([[0.0], gammainc(alpha + 1.0, cuts * alpha), [1.0]])
    return ncat * np.diff(upper)


def p_jc69(t):
    e = np.exp(-4.0 * t / 3.0)
    return np.full((4, 4), 0.25 - 0.25 * e) + np.eye(4) * e


def p_f81(freqs, t):
    pi = np.asarray(freqs, dtype=float)
    beta = 1.0 / (1.0 - np.sum(pi * pi))
    e = np.exp(-beta * t)
    return e * np.eye(4) + (1.0 - e) * np.tile(pi, (4, 1))


def hky_exch(kappa):
    return np.array([1.0, kappa, 1.0, 1.0, kappa, 1.0])


def p_hky85(freqs, kappa, t):
    """HKY85 transition probabilities (Hasegawa et al. 1985), mean rate normalised to 1."""
    pi = np.asarray(freqs, dtype=float)
    A, C, G, T = pi
    R, Y = A + G, C + T
    beta = 1.0 / (2.0 * R * Y + 2.0 * kappa * (A * G + C * T))
    group = [R, Y, R, Y]  # purines A, G; pyrimidines C, T
    P = np.zeros((4, 4))
    for i in range(4):
        for j in range(4):
            Pj = group[j]
            e1 = np.exp(-beta * t)
            e2 = np.exp(-beta * t * (1.0 + Pj * (kappa - 1.0)))
            if i == j:
                P[i, j] = pi[j] + pi[j] * (1.0 / Pj - 1.0) * e1 + (Pj - pi[j]) / Pj * e2
            elif group[i] == group[j] and (i % 2) == (j % 2):  # transition: same group
                P[i, j] = pi[j] + pi[j] * (1.0 / Pj - 1.0) * e1 - pi[j] / Pj * e2
            else:
                P[i, j] = pi[j] * (1.0 - e1)
    return P


def star_lnl(pfun, freqs, tip_states, lengths, weights=None, rates=(1.0,)):
    """lnL of a star tree: tips with observed states tip_states[tip][site] hang off one centre node by
    branches `lengths`; site likelihood = mean over rate categories of sum_k pi_k prod_tips P_k,x(r t)."""
    pi = np.asarray(freqs, dtype=float)
    tip_states = np.asarray(tip_states)
    n_sites = tip_states.shape[1]
    w = np.ones(n_sites) if weights is None else np.asarray(weights, dtype=float)
    site = np.zeros(n_sites)
    for r in rates:
        Ps = [pfun(r * t) for t in lengths]
        for s in range(n_sites):
            col = pi.copy()
            for P, x in zip(Ps, tip_states[:, s]):
                col = col * P[:, x]
            site[s] += col.sum() / len(rates)
    return float(np.sum(w * np.log(site)))


def tree_lnl_sum_over_states(pfun, freqs, desc, brlen, tip_states, weights=None, rates=(1.0,)):
    """lnL of a small rooted binary tree straight from the definition: the sum over ALL assignments of
    states to the inner nodes of pi(root) * prod over branches of P(parent -> child) (no pruning, no
    scaling).  desc/brlen as in include/pylogeny_b200.h; tip_states[tip][site] in 0..3."""
    import itertools
    pi = np.asarray(freqs, dtype=float)
    desc = np.asarray(desc)
    brlen = np.asarray(brlen, dtype=float)
    tip_states = np.asarray(tip_states)
    n = desc.shape[0] + 1
    ni = desc.shape[0]
    n_sites = tip_states.shape[1]
    w = np.ones(n_sites) if weights is None else np.asarray(weights, dtype=float)
    site = np.zeros(n_sites)
    for r in rates:
        P = [[pfun(r * brlen[i][k]) for k in range(2)] for i in range(ni)]
        for s in range(n_sites):
            tot = 0.0
            for inner in itertools.product(range(4), repeat=ni):
                state = lambda node: tip_states[node, s] if node < n else inner[node - n]
                p = pi[inner[ni - 1]]
                for i in range(ni):
                    for k in range(2):
                        p *= P[i][k][inner[i], state(int(desc[i][k]))]
                tot += p
            site[s] += tot / len(rates)
    return float(np.sum(w * np.log(site)))
