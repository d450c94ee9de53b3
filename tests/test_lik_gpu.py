"""Parity of the CUDA likelihood path (through the C-ABI) with the FP64 CPU oracle at identical
(topology, branch lengths, Q, pi, alpha).  Tolerance: 1e-8 relative in lnL (BASELINE.json north_star);
observed differences are ~1e-13 (FMA contraction / summation order)."""
import numpy as np
import pytest

import oracle
from util import random_desc

pytestmark = pytest.mark.gpu
RTOL = 1e-8


def _dna_case(rng, n, P, amb=0.02):
    codes = (1 << rng.integers(0, 4, (n, P))).astype(np.uint8)
    codes = np.where(rng.random((n, P)) < amb, rng.integers(1, 16, (n, P)), codes).astype(np.uint8)
    w = np.where(rng.random(P) < 0.2, rng.integers(1, 30, P), 1).astype(float)
    return codes, w


def test_model_host_pieces_vs_oracle():
    from pylogeny_b200 import likelihood
    rng = np.random.default_rng(2)
    for K in (4, 20):
        exch = rng.uniform(0.1, 4.0, K * (K - 1) // 2)
        f = rng.dirichlet(np.ones(K) * 6)
        for alpha in (0.2, 1.0, 3.3):
            m = likelihood.Model(exch, f, alpha, 4)
            np.testing.assert_allclose(m.gamma_rates(), oracle.gamma_rates(alpha, 4), rtol=1e-12)
            for t in (0.0, 0.013, 0.9):
                np.testing.assert_allclose(m.pmatrix(t), oracle.pmatrix(exch, f, t), rtol=1e-10, atol=1e-14)
            m.close()


@pytest.mark.parametrize("n,P", [(2, 1), (3, 5), (8, 300), (64, 2000)])
def test_dna_gtr_gamma_vs_oracle(n, P):
    from pylogeny_b200 import likelihood
    rng = np.random.default_rng(100 + n)
    codes, w = _dna_case(rng, n, P)
    exch = rng.uniform(0.3, 3.0, 6)
    f = rng.dirichlet(np.ones(4) * 8)
    alpha = 0.6
    a = likelihood.LikAlignment(codes, w)
    m = likelihood.Model(exch, f, alpha)
    descs = np.stack([random_desc(rng, n) for _ in range(6)])
    brs = rng.exponential(0.1, descs.shape)
    got = a.score_trees(m, descs, brs)
    want = [oracle.lnl(codes, w, exch, f, alpha, d, b) for d, b in zip(descs, brs)]
    np.testing.assert_allclose(got, want, rtol=RTOL, atol=0)
    assert np.max(np.abs(got - want) / np.abs(want)) < 1e-11
    a.close(); m.close()


def test_protein_gamma_vs_oracle():
    from pylogeny_b200 import _lib, likelihood
    rng = np.random.default_rng(55)
    n, P = 12, 400
    codes = rng.integers(0, 20, (n, P)).astype(np.uint8)
    codes = np.where(rng.random((n, P)) < 0.03, rng.integers(20, 23, (n, P)), codes).astype(np.uint8)
    w = rng.integers(1, 4, P).astype(float)
    exch = rng.uniform(0.05, 5.0, 190)
    f = rng.dirichlet(np.ones(20) * 10)
    a = likelihood.LikAlignment(codes, w, _lib.DATA_AA)
    m = likelihood.Model(exch, f, 0.9)
    descs = np.stack([random_desc(rng, n) for _ in range(4)])
    brs = rng.exponential(0.15, descs.shape)
    got = a.score_trees(m, descs, brs)
    want = [oracle.lnl(codes, w, exch, f, 0.9, d, b) for d, b in zip(descs, brs)]
    np.testing.assert_allclose(got, want, rtol=RTOL, atol=0)
    a.close(); m.close()


@pytest.mark.parametrize("n,P,scale", [(12, 400, 1.0), (70, 130, 30.0), (5, 64, 1.0)])
def test_protein_staged_update_is_bit_identical(n, P, scale, monkeypatch):
    """The 20-state update behind every protein level program runs with its operands staged by cp.async
    (clv20s_kernel, pattern-tile-major blocks); PLG_A20_STAGED=0 keeps clv20_kernel.  Same products, same
    scaling rule -- identical lnL, also where 2^-256 factors are taken (long branches, 70 taxa) -- and the
    oracle on one tree."""
    from pylogeny_b200 import _lib, likelihood
    rng = np.random.default_rng(3 * n + P)
    codes = rng.integers(0, 20, (n, P)).astype(np.uint8)
    codes = np.where(rng.random((n, P)) < 0.03, rng.integers(20, 23, (n, P)), codes).astype(np.uint8)
    w = rng.integers(0, 4, P).astype(float)
    exch, f = likelihood.protein_model("JTT")
    a = likelihood.LikAlignment(codes, w, _lib.DATA_AA)
    m = likelihood.Model(exch, f, 0.7)
    descs = np.stack([random_desc(rng, n) for _ in range(5)])
    brs = rng.exponential(0.1 * scale, descs.shape)
    got = a.score_trees(m, descs, brs)
    monkeypatch.setenv("PLG_A20_STAGED", "0")
    ref = a.score_trees(m, descs, brs)
    assert got.tolist() == ref.tolist()
    want = oracle.lnl(codes, w, exch, f, 0.7, descs[0], brs[0])
    assert abs(got[0] - want) <= RTOL * abs(want)
    a.close(); m.close()


def test_scaling_counters_deep_tree():
    """600-taxon caterpillar with long branches: per-site CLVs underflow 2^-256 several times."""
    from pylogeny_b200 import likelihood
    rng = np.random.default_rng(5)
    n, P = 600, 40
    desc = np.array([(0, 1)] + [(n + i - 1, i + 1) for i in range(1, n - 1)], np.int32)
    br = rng.exponential(0.5, (n - 1, 2))
    codes = (1 << rng.integers(0, 4, (n, P))).astype(np.uint8)
    w = np.ones(P)
    exch, f = np.ones(6), np.array([0.1, 0.2, 0.3, 0.4])
    a = likelihood.LikAlignment(codes, w)
    m = likelihood.Model(exch, f, 1.0)
    got = a.score_trees(m, desc, br)[0]
    want = oracle.lnl(codes, w, exch, f, 1.0, desc, br)
    assert want < -700 * P / 3
    assert abs(got - want) <= RTOL * abs(want)
    a.close(); m.close()


def test_rate_categories_one():
    from pylogeny_b200 import likelihood
    rng = np.random.default_rng(8)
    codes, w = _dna_case(rng, 10, 257)
    exch, f = rng.uniform(0.5, 2, 6), rng.dirichlet(np.ones(4) * 5)
    a = likelihood.LikAlignment(codes, w)
    m = likelihood.Model(exch, f, 1.0, 1)
    d = random_desc(rng, 10)
    b = rng.exponential(0.1, d.shape)
    got = a.score_trees(m, d, b)[0]
    want = oracle.lnl(codes, w, exch, f, 1.0, d, b, ncat=1)
    assert abs(got - want) <= RTOL * abs(want)
    a.close(); m.close()


def test_reroot_invariance_full_width():
    """BASELINE config 2 shape (64 taxa x 10k sites): pulley principle -- moving the root along
    an edge leaves lnL unchanged (size-independent property), plus one tree against the oracle."""
    from pylogeny_b200 import likelihood
    rng = np.random.default_rng(1002)
    n, P = 64, 10_000
    codes = (1 << rng.integers(0, 4, (n, P))).astype(np.uint8)
    w = np.ones(P)
    exch, f = np.array([1.0, 2.5, 0.7, 1.2, 3.1, 1.0]), np.array([0.3, 0.2, 0.25, 0.25])
    a = likelihood.LikAlignment(codes, w)
    m = likelihood.Model(exch, f, 0.5)
    d = random_desc(rng, n)
    b = rng.exponential(0.1, d.shape)
    b2 = b.copy()
    tot = b[-1].sum()
    b2[-1] = (0.25 * tot, 0.75 * tot)
    got = a.score_trees(m, np.stack([d, d]), np.stack([b, b2]))
    assert abs(got[0] - got[1]) <= 1e-12 * abs(got[0])
    want = oracle.lnl(codes, w, exch, f, 0.5, d, b)
    assert abs(got[0] - want) <= RTOL * abs(want)
    a.close(); m.close()
