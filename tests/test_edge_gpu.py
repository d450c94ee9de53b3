"""Edge cases of the CUDA path: minimum sizes, word-boundary pattern counts, wide state alphabets,
degenerate branch lengths, ambiguity-only columns -- each against the CPU oracle."""
import numpy as np
import pytest

import oracle
from util import desc_to_newick, random_desc, random_states, states_to_profiles

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("P", [1, 31, 32, 127, 128, 129, 255, 257])
def test_fitch_pattern_count_boundaries(P):
    from pylogeny_b200 import fitch
    rng = np.random.default_rng(P)
    n = 7
    st = random_states(rng, n, P, 4, 0.1)
    w = rng.integers(0, 4, P).astype(np.int64)            # includes zero weights
    a = fitch.FitchAlignment.from_dense(st, w)
    d = np.stack([random_desc(rng, n) for _ in range(5)])
    labels = ["T%d" % i for i in range(n)]
    x = oracle.FitchInputs(states_to_profiles(st), w.tolist(), {l: i for i, l in enumerate(labels)})
    assert a.score_trees(d).tolist() == [oracle.fitch_cost(desc_to_newick(t, labels), inputs=x) for t in d]
    a.close()


@pytest.mark.parametrize("ns", [2, 3, 6, 8, 10, 13, 20])
def test_fitch_wide_alphabets(ns):
    """Every plane-count template instance (2..20 planes); symbols are arbitrary bytes."""
    from pylogeny_b200 import fitch
    rng = np.random.default_rng(ns)
    n, P = 24, 300
    raw = rng.integers(0, ns, (P, n))
    raw[0, :ns] = np.arange(ns)                             # one column that really has ns states
    st = (raw + 65).astype(np.uint8)                        # 'A'.. : not digits
    w = rng.integers(1, 5, P).astype(np.int64)
    a = fitch.FitchAlignment.from_dense(st, w)
    assert a.info()["n_planes"] >= ns
    d = np.stack([random_desc(rng, n) for _ in range(3)])
    labels = ["T%d" % i for i in range(n)]
    x = oracle.FitchInputs(states_to_profiles(st), w.tolist(), {l: i for i, l in enumerate(labels)})
    assert a.score_trees(d).tolist() == [oracle.fitch_cost(desc_to_newick(t, labels), inputs=x) for t in d]
    a.close()


def test_fitch_too_many_states_is_an_error():
    from pylogeny_b200 import _lib, fitch
    st = np.arange(21, dtype=np.uint8)[None, :] + 40
    with pytest.raises(_lib.EngineError) as e:
        fitch.FitchAlignment.from_dense(st, np.ones(1, dtype=np.int64))
    assert e.value.code == _lib.ERR_STATES


def test_minimum_trees_and_neighbourhoods():
    from pylogeny_b200 import fitch, likelihood, neighbourhood as nbh
    rng = np.random.default_rng(1)
    st = random_states(rng, 4, 40, 4)
    a = fitch.FitchAlignment.from_dense(st, np.ones(40, dtype=np.int64))
    codes = (1 << rng.integers(0, 4, (4, 40))).astype(np.uint8)
    la = likelihood.LikAlignment(codes, np.ones(40))
    m = likelihood.Model(np.ones(6), np.full(4, 0.25), 1.0)
    # two tips: a single branch
    d2 = np.array([[0, 1]], np.int32)
    x = oracle.FitchInputs(states_to_profiles(st), [1] * 40, {"T0": 0, "T1": 1})
    assert a.score_trees(d2[None])[0] == oracle.fitch_cost("(T0,T1);", inputs=x)
    b2 = np.array([[0.3, 0.2]])
    assert abs(la.score_trees(m, d2, b2)[0] - oracle.lnl(codes, np.ones(40), np.ones(6), np.full(4, .25), 1.0, d2, b2)) < 1e-9
    # three tips: no SPR neighbours; four tips: 2 SPR = 2 NNI neighbours
    d3 = np.array([[0, 1], [3, 2]], np.int32)
    assert len(nbh.score_parsimony(a, d3)[1]) == 0
    d4 = np.array([[0, 1], [2, 3], [4, 5]], np.int32)
    br4 = rng.exponential(0.1, d4.shape)
    mv, c = nbh.score_parsimony(a, d4)
    descs, brs, _ = nbh.neighbour_trees(d4, br4)
    assert len(c) == 2 and c.tolist() == a.score_trees(descs).tolist()
    _, l = nbh.score_likelihood(la, m, d4, br4)
    np.testing.assert_allclose(l, la.score_trees(m, descs, brs), rtol=1e-12)
    a.close(); la.close(); m.close()


def test_lik_degenerate_branches_and_gap_columns():
    from pylogeny_b200 import likelihood
    rng = np.random.default_rng(2)
    n, P = 9, 200
    codes = (1 << rng.integers(0, 4, (n, P))).astype(np.uint8)
    codes[:, :10] = 15                                       # all-gap columns: site likelihood 1
    codes[3, :] = 15                                         # a taxon with no data at all
    w = np.ones(P)
    exch, f = rng.uniform(0.5, 2, 6), rng.dirichlet(np.ones(4) * 5)
    a = likelihood.LikAlignment(codes, w)
    m = likelihood.Model(exch, f, 0.5)
    d = random_desc(rng, n)
    for br in (np.zeros(d.shape), np.full(d.shape, 50.0), rng.exponential(0.1, d.shape) * (rng.random(d.shape) < 0.5)):
        got = a.score_trees(m, d, br)[0]
        want = oracle.lnl(codes, w, exch, f, 0.5, d, br)
        assert not np.isnan(got)
        if np.isinf(want):   # different tips joined by zero-length paths: the data are impossible, lnL = -inf
            assert got == want
        else:
            assert abs(got - want) <= 1e-8 * max(1.0, abs(want))
    a.close(); m.close()


def test_lik_bad_inputs():
    from pylogeny_b200 import _lib, likelihood
    with pytest.raises(_lib.EngineError):
        likelihood.LikAlignment(np.full((3, 5), 16, np.uint8), np.ones(5))          # DNA code out of range
    with pytest.raises(_lib.EngineError):
        likelihood.Model(np.ones(6), np.array([0.5, 0.5, 0.0, 0.0]), 1.0)           # zero frequency
    a = likelihood.LikAlignment(np.ones((3, 5), np.uint8), np.ones(5))
    m = likelihood.Model(np.ones(6), np.full(4, 0.25), 1.0)
    with pytest.raises(_lib.EngineError):
        a.score_trees(m, np.array([[0, 1], [3, 7]], np.int32), np.ones((2, 2)))      # child id not earlier
    with pytest.raises(_lib.EngineError):
        a.score_trees(m, np.array([[0, 1], [3, 2]], np.int32), np.ones((2, 2)), tip_rows=[0, 1, 9])
    a.close(); m.close()


def test_newick_forms_accepted_by_the_reference_parser():
    """Branch lengths in scientific notation, internal labels / support values, no trailing ';'."""
    from pylogeny_b200 import fitch
    profs, w, taxa = ["0011", "0101", "0012"], [2, 3, 1], {"A": 0, "B": 1, "C": 2, "D": 3}
    x = oracle.FitchInputs(profs, w, taxa)
    for nw in ("((A:1e-05,B:0.1):0.5,(C:1,D:2):0.25);", "((A,B)0.95:0.1,(C,D)n1:0.2)root;", "((A,B),(C,D))",
               "(A:0.1,B:0.2,(C:0.3,D:0.4):0.5);"):
        assert fitch.calculateCost(nw, profs, w, taxa) == oracle.fitch_cost(nw, inputs=x), nw
