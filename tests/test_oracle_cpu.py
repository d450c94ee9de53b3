"""Pin the CPU oracles (test infrastructure) before anything is compared with them."""
import itertools
import random

import numpy as np
import pytest

import oracle


def test_fitch_oracle_named_golden(al_profiles, al_fitch):
    x = oracle.FitchInputs(al_profiles["profiles"], al_profiles["weights"], al_profiles["taxa"])
    got = {c["name"]: oracle.fitch_cost(c["newick"], inputs=x) for c in al_fitch["named"]}
    want = {c["name"]: c["cost"] for c in al_fitch["named"]}
    assert got == want
    # BASELINE.md section 5 known answers
    assert want["caterpillar_file_order"] == 1167 and want["balanced"] == 1296
    assert want["caterpillar_sorted"] == 1396 and want["unrooted_trifurcation"] == 1103


def test_fitch_oracle_spr_golden(al_profiles, al_fitch):
    x = oracle.FitchInputs(al_profiles["profiles"], al_profiles["weights"], al_profiles["taxa"])
    costs = [oracle.fitch_cost(m["newick"], inputs=x) for m in al_fitch["spr_moves"]]
    assert costs == [m["cost"] for m in al_fitch["spr_moves"]]
    assert len(costs) == 120 and min(costs) == 1149 and max(costs) == 1416


def test_fitch_oracle_greedy_calls(al_profiles, al_greedy):
    x = oracle.FitchInputs(al_profiles["profiles"], al_profiles["weights"], al_profiles["taxa"])
    assert al_greedy["path_scores"] == [1167, 1149, 1138, 1135] and al_greedy["n_trees"] == 306
    for nw, c in al_greedy["calls"]:
        assert oracle.fitch_cost(nw, inputs=x) == c


def test_fitch_oracle_random_golden(random_fitch):
    for c in random_fitch:
        assert oracle.fitch_cost(c["newick"], c["profiles"], c["weights"], c["taxa"]) == c["cost"]


@pytest.mark.skipif(not oracle.ref_available(), reason="oracle/_ref not built (needs /root/reference)")
def test_fitch_oracle_vs_compiled_reference_live():
    rng = random.Random(7)
    from golden.make_golden import random_newick
    for it in range(200):
        n = rng.randint(2, 20)
        labels = ["T%d" % i for i in range(n)]
        taxa = {l: i for i, l in enumerate(labels)}
        P = rng.randint(1, 30)
        profs = ["".join(str(rng.randrange(rng.choice([2, 5, 10]))) for _ in range(n)) for _ in range(P)]
        w = [rng.randint(1, 9) for _ in range(P)]
        nw = random_newick(rng, labels, multif=rng.choice([0.0, 0.5]), lengths=rng.random() < 0.5)
        assert oracle.fitch_cost(nw, profs, w, taxa) == oracle.ref_fitch_cost(nw, profs, w, taxa), nw


def test_gamma_rates_vs_scipy():
    from scipy import special, stats
    for alpha in (0.05, 0.3, 1.0, 2.5, 20.0):
        for ncat in (4, 8):
            r = oracle.gamma_rates(alpha, ncat)
            cuts = stats.gamma.ppf(np.arange(1, ncat) / ncat, alpha, scale=1.0 / alpha)
            cdf = np.concatenate([[0.0], special.gammainc(alpha + 1.0, cuts * alpha), [1.0]])
            want = np.diff(cdf) * ncat
            np.testing.assert_allclose(r, want, rtol=1e-9, atol=1e-300)
            assert abs(r.mean() - 1.0) < 1e-12


def test_pmatrix_vs_expm():
    from scipy.linalg import expm
    rng = np.random.default_rng(3)
    for K in (4, 20):
        exch = rng.uniform(0.1, 5.0, K * (K - 1) // 2)
        f = rng.dirichlet(np.ones(K) * 5)
        Q = np.zeros((K, K))
        iu = np.triu_indices(K, 1)
        Q[iu] = exch
        Q = Q + Q.T
        Q = Q * f[None, :]
        np.fill_diagonal(Q, 0)
        np.fill_diagonal(Q, -Q.sum(1))
        Q /= -(f * np.diag(Q)).sum()
        for t in (0.0, 1e-6, 0.1, 1.0, 7.5):
            np.testing.assert_allclose(oracle.pmatrix(exch, f, t), expm(Q * t), rtol=1e-9, atol=1e-12)


def _rand_tree_desc(rng, n):
    nodes = list(range(n))
    desc, br = [], []
    nxt = n
    while len(nodes) > 1:
        a = nodes.pop(rng.integers(len(nodes)))
        b = nodes.pop(rng.integers(len(nodes)))
        desc.append((a, b))
        br.append(rng.exponential(0.2, 2))
        nodes.append(nxt)
        nxt += 1
    return np.array(desc, np.int32), np.array(br)


def test_lnl_oracle_vs_bruteforce():
    rng = np.random.default_rng(11)
    for K, n in ((4, 5), (4, 6), (20, 4)):
        desc, br = _rand_tree_desc(rng, n)
        ncodes = 16 if K == 4 else 23
        codes = rng.integers(1 if K == 4 else 0, ncodes, (n, 7)).astype(np.uint8)
        w = rng.integers(1, 5, 7).astype(float)
        exch = rng.uniform(0.2, 3.0, K * (K - 1) // 2)
        f = rng.dirichlet(np.ones(K) * 4)
        a = oracle.lnl(codes, w, exch, f, 0.7, desc, br)
        b = oracle.lnl(codes, w, exch, f, 0.7, desc, br, brute=True)
        assert abs(a - b) <= 1e-11 * abs(b)


def test_lnl_oracle_normalises_frequencies():
    """State frequencies are a distribution: a table that sums to 1 only within its printed digits
    (PAML's wag.dat sums to 1 + 3e-6, which would shift lnL by P * log(sum)) scores like its normalised
    form -- the convention the engine follows too (csrc/model.cpp), pinned on the GPU by
    tests/test_configs_gpu.py::test_c5 on the shipped WAG table."""
    rng = np.random.default_rng(12)
    desc, br = _rand_tree_desc(rng, 5)
    codes = rng.integers(0, 20, (5, 9)).astype(np.uint8)
    exch, f = rng.uniform(0.2, 3.0, 190), rng.dirichlet(np.ones(20) * 4)
    a = oracle.lnl(codes, np.ones(9), exch, f, 0.7, desc, br)
    for scale in (1.0 + 3.2e-6, 0.5, 7.0):
        assert abs(oracle.lnl(codes, np.ones(9), exch, f * scale, 0.7, desc, br) - a) <= 1e-13 * abs(a)
        assert abs(oracle.lnl(codes, np.ones(9), exch, f * scale, 0.7, desc, br, brute=True) - a) <= 1e-11 * abs(a)


def test_lnl_oracle_scaling_long_tree():
    """Deep caterpillar with long branches forces the 2^-256 rescaling path; compare with
    a log-space computation in extended precision via per-site renormalisation."""
    rng = np.random.default_rng(5)
    n = 600
    desc = [(0, 1)] + [(n + i - 1, i + 1) for i in range(1, n - 1)]
    desc = np.array(desc, np.int32)
    br = rng.exponential(0.5, (n - 1, 2))
    codes = (1 << rng.integers(0, 4, (n, 3))).astype(np.uint8)
    w = np.ones(3)
    exch = np.ones(6)
    f = np.array([0.1, 0.2, 0.3, 0.4])
    got = oracle.lnl(codes, w, exch, f, 1.0, desc, br)
    # independent numpy pruning with per-node max renormalisation
    ev, U, Ui = oracle.eigen(exch, f)
    rates = oracle.gamma_rates(1.0, 4)
    tot = 0.0
    for p in range(3):
        site = []
        for r in rates:
            clv = {}
            lg = {}
            for i, (a, b) in enumerate(desc):
                v = np.ones(4)
                l = 0.0
                for c, t in ((a, br[i, 0]), (b, br[i, 1])):
                    Pm = (U * np.exp(ev * t * r)) @ Ui
                    if c < n:
                        x = np.array([(codes[c, p] >> k) & 1 for k in range(4)], float)
                    else:
                        x = clv[c]
                        l += lg[c]
                    v = v * (Pm @ x)
                m = v.max()
                clv[n + i] = v / m
                lg[n + i] = l + np.log(m)
            root = n + len(desc) - 1
            site.append(np.log((f * clv[root]).sum()) + lg[root])
        site = np.array(site)
        mx = site.max()
        tot += mx + np.log(np.exp(site - mx).mean())
    assert got < -700  # really in the underflow regime
    assert abs(got - tot) <= 1e-10 * abs(tot)


def test_closed_form_formulas_vs_expm():
    """The textbook P(t) formulas of tests/closed_form.py are themselves transcribed correctly:
    they equal expm(Qt) of the rate matrix they claim to solve."""
    from scipy.linalg import expm
    import closed_form as cf
    import closed_form_cases as cc

    def rate_matrix(exch, pi):
        R = np.zeros((4, 4))
        R[np.triu_indices(4, 1)] = exch
        R = R + R.T
        Q = R * pi[None, :]
        np.fill_diagonal(Q, -Q.sum(1))
        return Q / -np.sum(pi * np.diag(Q))

    for t in (0.0, 0.01, 0.3, 2.0, 9.0):
        np.testing.assert_allclose(cf.p_jc69(t), expm(rate_matrix(np.ones(6), np.full(4, 0.25)) * t), atol=1e-14)
        np.testing.assert_allclose(cf.p_f81(cc.F81_FREQS, t), expm(rate_matrix(np.ones(6), cc.F81_FREQS) * t), atol=1e-14)
        np.testing.assert_allclose(cf.p_hky85(cc.HKY_FREQS, 4.3, t),
                                   expm(rate_matrix(cf.hky_exch(4.3), cc.HKY_FREQS) * t), atol=1e-14)
    # Yang's mean rates: mean 1, increasing, and the alpha -> infinity limit is a single rate
    for alpha in (0.2, 1.0, 3.0):
        r = cf.gamma_mean_rates(alpha, 4)
        assert abs(r.mean() - 1.0) < 1e-12 and np.all(np.diff(r) > 0)
    np.testing.assert_allclose(cf.gamma_mean_rates(1e7, 4), np.ones(4), atol=2e-3)


def test_lnl_oracle_pinned_to_closed_forms():
    """oracle/lik_oracle.c against numbers nobody here computed: JC69 / F81 / HKY85 closed-form
    transition probabilities, Yang-1994 mean Gamma rates from scipy, two-, three- and four-taxon
    trees whose site likelihoods are explicit sums (star trees, the pulley principle, the
    zero-length internal branch, saturation), every site pattern, at 1e-12 relative."""
    import closed_form_cases as cc
    cases = cc.cases()
    assert len(cases) >= 70
    for name, exch, freqs, alpha, ncat, codes, w, desc, br, want in cases:
        got = oracle.lnl(codes, w, exch, freqs, alpha, desc, br, ncat=ncat)
        assert abs(got - want) <= 1e-12 * abs(want), (name, got, want)
