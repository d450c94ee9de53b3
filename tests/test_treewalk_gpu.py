"""Tree walks (csrc/treewalk.cu, fitch_walk.cuh, lik_twalk.cuh): descriptor batches scored by the
on-chip depth-first kernels must equal the level-synchronous programs (PLG_TREE_MODE=levels) and
the oracle -- Fitch bit-exact, lnL within 1e-8 relative (observed ~1e-13)."""
import os

import numpy as np
import pytest

import oracle
from util import desc_to_newick, random_desc, random_states, states_to_profiles

pytestmark = pytest.mark.gpu


def caterpillar(n, order=None):
    order = list(range(n)) if order is None else order
    d, cur = [], order[0]
    for i in range(1, n):
        d.append((cur, order[i]) if i % 2 else (order[i], cur))
        cur = n + i - 1
    return np.array(d, np.int32)


def balanced(n):
    nodes, d, nxt = list(range(n)), [], n
    while len(nodes) > 1:
        nn = []
        for i in range(0, len(nodes) - 1, 2):
            d.append((nodes[i], nodes[i + 1])); nn.append(nxt); nxt += 1
        if len(nodes) % 2:
            nn.append(nodes[-1])
        nodes = nn
    return np.array(d, np.int32)


def shapes(rng, n):
    out = [caterpillar(n), caterpillar(n, list(rng.permutation(n))), balanced(n)]
    out += [random_desc(rng, n) for _ in range(5)]
    return np.stack(out)


@pytest.mark.parametrize("n,P,states,gaps", [(2, 3, 4, 0.0), (3, 130, 4, 0.1), (17, 1000, 4, 0.05), (64, 5000, 4, 0.0),
                                              (33, 700, 9, 0.0), (40, 300, 19, 0.0), (300, 257, 4, 0.02)])
def test_fitch_walk_equals_levels_and_oracle(n, P, states, gaps, monkeypatch):
    from pylogeny_b200 import fitch
    rng = np.random.default_rng(n * 7 + P)
    st = random_states(rng, n, P, states, gaps)
    w = np.where(rng.random(P) < 0.3, rng.integers(1, 2000, P), 1).astype(np.int64)
    a = fitch.FitchAlignment.from_dense(st, w)
    descs = shapes(rng, n)
    monkeypatch.setenv("PLG_TREE_MODE", "walk")
    got = a.score_trees(descs)
    monkeypatch.setenv("PLG_TREE_MODE", "levels")
    lev = a.score_trees(descs)
    assert got.tolist() == lev.tolist()
    labels = ["T%d" % i for i in range(n)]
    x = oracle.FitchInputs(states_to_profiles(st), w.tolist(), {l: i for i, l in enumerate(labels)})
    for i in (0, 2, 5):
        assert got[i] == oracle.fitch_cost(desc_to_newick(descs[i], labels), inputs=x)
    # tip_rows: the same trees over a permuted row order
    perm = rng.permutation(n).astype(np.int32)
    monkeypatch.setenv("PLG_TREE_MODE", "walk")
    got_p = a.score_trees(descs, perm)
    monkeypatch.setenv("PLG_TREE_MODE", "levels")
    assert got_p.tolist() == a.score_trees(descs, perm).tolist()
    a.close()


def test_fitch_walk_unit_weights_and_wrap(monkeypatch):
    """All-ones weights take the popcount-only path; huge weights wrap mod 2^32 like fitch.cpp:247,268."""
    from pylogeny_b200 import fitch
    rng = np.random.default_rng(3)
    n, P = 24, 4000
    st = random_states(rng, n, P, 4)
    descs = shapes(rng, n)
    for w in (np.ones(P, np.int64), rng.integers(2 ** 28, 2 ** 31, P).astype(np.int64)):
        a = fitch.FitchAlignment.from_dense(st, w)
        monkeypatch.setenv("PLG_TREE_MODE", "walk")
        got = a.score_trees(descs)
        monkeypatch.setenv("PLG_TREE_MODE", "levels")
        assert got.tolist() == a.score_trees(descs).tolist()
        a.close()


@pytest.mark.parametrize("n,P,scale", [(2, 1, 1.0), (3, 5, 1.0), (9, 300, 1.0), (64, 2000, 1.0), (130, 400, 1.0),
                                        (48, 500, 40.0)])
def test_lnl_walk_equals_levels_and_oracle(n, P, scale, monkeypatch):
    """scale = 40: branch lengths ~4 => the 2^-256 rescaling runs inside the walk."""
    from pylogeny_b200 import likelihood
    rng = np.random.default_rng(1000 + n)
    codes = (1 << rng.integers(0, 4, (n, P))).astype(np.uint8)
    codes = np.where(rng.random((n, P)) < 0.03, rng.integers(0, 16, (n, P)), codes).astype(np.uint8)
    w = np.where(rng.random(P) < 0.2, rng.integers(1, 30, P), 1).astype(float)
    exch, f, alpha = rng.uniform(0.3, 3.0, 6), rng.dirichlet(np.ones(4) * 8), 0.55
    a = likelihood.LikAlignment(codes, w)
    m = likelihood.Model(exch, f, alpha)
    descs = shapes(rng, n)
    brs = rng.exponential(0.1 * scale, descs.shape)
    brs[0, 0, 0] = 0.0  # a zero-length branch is the identity exactly
    monkeypatch.setenv("PLG_TREE_MODE", "walk")
    got = a.score_trees(m, descs, brs)
    monkeypatch.setenv("PLG_TREE_MODE", "levels")
    lev = a.score_trees(m, descs, brs)
    np.testing.assert_allclose(got, lev, rtol=1e-12, atol=0)
    for i in (0, 2, 6):
        want = oracle.lnl(codes, w, exch, f, alpha, descs[i], brs[i])
        assert abs(got[i] - want) <= 1e-8 * abs(want)
    perm = rng.permutation(n).astype(np.int32)
    monkeypatch.setenv("PLG_TREE_MODE", "walk")
    got_p = a.score_trees(m, descs, brs, perm)
    monkeypatch.setenv("PLG_TREE_MODE", "levels")
    np.testing.assert_allclose(got_p, a.score_trees(m, descs, brs, perm), rtol=1e-12, atol=0)
    a.close(); m.close()


def test_walk_all_unambiguous_columns(monkeypatch):
    """Pure ACGT data takes the column-of-P tip path in every warp."""
    from pylogeny_b200 import likelihood
    rng = np.random.default_rng(77)
    n, P = 20, 1500
    codes = (1 << rng.integers(0, 4, (n, P))).astype(np.uint8)
    a = likelihood.LikAlignment(codes, np.ones(P))
    m = likelihood.Model(np.ones(6), np.full(4, 0.25), 1.0)
    descs = shapes(rng, n)
    brs = rng.exponential(0.2, descs.shape)
    got = a.score_trees(m, descs, brs)
    for i in (1, 3):
        want = oracle.lnl(codes, np.ones(P), np.ones(6), np.full(4, 0.25), 1.0, descs[i], brs[i])
        assert abs(got[i] - want) <= 1e-8 * abs(want)
    a.close(); m.close()


def test_walk_batches_larger_than_a_chunk(monkeypatch):
    """Tree lists are processed in chunks that fit the planner / P-matrix buffers (config 4: 100k trees);
    PLG_WALK_CHUNK forces several chunks on a small case -- same scores, same order."""
    from pylogeny_b200 import fitch, likelihood
    rng = np.random.default_rng(5)
    n, P, T = 21, 900, 23
    st = random_states(rng, n, P, 4, 0.03)
    w = rng.integers(1, 40, P).astype(np.int64)
    descs = np.stack([random_desc(rng, n) for _ in range(T)])
    brs = rng.exponential(0.1, descs.shape)
    codes = (1 << rng.integers(0, 4, (n, P))).astype(np.uint8)
    fa = fitch.FitchAlignment.from_dense(st, w)
    la = likelihood.LikAlignment(codes, w.astype(float))
    m = likelihood.Model(rng.uniform(0.5, 2, 6), rng.dirichlet(np.ones(4) * 5), 0.8)
    want_f, want_l = fa.score_trees(descs), la.score_trees(m, descs, brs)
    for chunk in ("1", "5", "22"):
        monkeypatch.setenv("PLG_WALK_CHUNK", chunk)
        assert fa.score_trees(descs).tolist() == want_f.tolist()
        np.testing.assert_array_equal(la.score_trees(m, descs, brs), want_l)
    fa.close(); la.close(); m.close()


def test_walk_rejects_bad_descriptors():
    """Same argument errors as the level path (child id not earlier in post-order, row outside the alignment)."""
    from pylogeny_b200 import fitch, likelihood, _lib
    rng = np.random.default_rng(6)
    st = random_states(rng, 3, 50, 4)
    fa = fitch.FitchAlignment.from_dense(st, np.ones(50, dtype=np.int64))
    la = likelihood.LikAlignment((1 << rng.integers(0, 4, (3, 50))).astype(np.uint8), np.ones(50))
    m = likelihood.Model(np.ones(6), np.full(4, 0.25), 1.0)
    bad = np.array([[0, 1], [3, 7]], np.int32)
    with pytest.raises(Exception):
        fa.score_trees(bad)
    with pytest.raises(Exception):
        la.score_trees(m, bad, np.ones((2, 2)))
    with pytest.raises(Exception):
        fa.score_trees(np.array([[0, 1], [3, 2]], np.int32), tip_rows=[0, 1, 9])
    with pytest.raises(Exception):
        la.score_trees(m, np.array([[0, 1], [3, 2]], np.int32), np.ones((2, 2)), tip_rows=[0, 1, -1])
    # not a tree: a node used twice as a child (the device planner would emit more ops than the tree's
    # slot holds), on the walk path and on the level path
    st4 = random_states(rng, 4, 50, 4)
    fa4 = fitch.FitchAlignment.from_dense(st4, np.ones(50, dtype=np.int64))
    la4 = likelihood.LikAlignment((1 << rng.integers(0, 4, (4, 50))).astype(np.uint8), np.ones(50))
    dag = np.array([[0, 1], [4, 4], [5, 5]], np.int32)
    forest = np.array([[0, 1], [2, 3], [4, 4]], np.int32)
    for mode in ("walk", "levels"):
        os.environ["PLG_TREE_MODE"] = mode
        try:
            for d in (dag, forest):
                with pytest.raises(_lib.EngineError) as e1:
                    fa4.score_trees(d)
                assert e1.value.code == _lib.ERR_ARG
                with pytest.raises(_lib.EngineError) as e2:
                    la4.score_trees(m, d, np.ones((3, 2)))
                assert e2.value.code == _lib.ERR_ARG
            # still fine afterwards
            assert fa4.score_trees(np.array([[0, 1], [2, 3], [4, 5]], np.int32)).shape == (1,)
        finally:
            os.environ.pop("PLG_TREE_MODE")
    with pytest.raises(_lib.EngineError):  # negative rows on the level path are not inner-node references
        os.environ["PLG_TREE_MODE"] = "levels"
        try:
            fa4.score_trees(np.array([[0, 1], [2, 3], [4, 5]], np.int32), tip_rows=[0, 1, 2, -2])
        finally:
            os.environ.pop("PLG_TREE_MODE")
    fa.close(); la.close(); m.close(); fa4.close(); la4.close()
