"""Host-side neighbourhood enumeration (no GPU): counts, distinctness, hash consistency."""
import numpy as np
import pytest

from util import random_desc


@pytest.mark.parametrize("n", [4, 5, 8, 13, 40])
def test_counts_and_hash_consistency(n):
    from pylogeny_b200 import neighbourhood as nbh
    rng = np.random.default_rng(n)
    d = random_desc(rng, n)
    M = nbh.neighbourhood_size(d, nbh.MOVE_SPR)
    assert M == 2 * (n - 3) * (2 * n - 7)            # distinct SPR neighbours of an unrooted binary tree
    assert nbh.neighbourhood_size(d, nbh.MOVE_NNI) == 2 * (n - 3)
    br = rng.exponential(0.1, d.shape)
    descs, brs, hashes = nbh.neighbour_trees(d, br)
    assert len(set(hashes.tolist())) == M             # all distinct
    base = nbh.topology_hash(d)
    assert base not in set(hashes.tolist())
    # the incrementally maintained hash equals the hash of the materialised neighbour
    for i in range(0, M, max(1, M // 60)):
        assert nbh.topology_hash(descs[i]) == int(hashes[i])
        # total branch length is conserved by the reference's halving / summing rule
        assert abs(brs[i].sum() - br.sum()) < 1e-12


def test_small_trees_have_no_neighbours():
    from pylogeny_b200 import neighbourhood as nbh
    assert nbh.neighbourhood_size(np.array([[0, 1]]), nbh.MOVE_SPR) == 0
    assert nbh.neighbourhood_size(np.array([[0, 1], [3, 2]]), nbh.MOVE_SPR) == 0


def test_hash_is_rooting_invariant():
    from pylogeny_b200 import neighbourhood as nbh
    # ((0,1),(2,3)) rooted three different ways is one unrooted topology; ((0,2),(1,3)) is another
    a = np.array([[0, 1], [2, 3], [4, 5]])
    b = np.array([[2, 3], [1, 4], [0, 5]])
    c = np.array([[0, 2], [1, 3], [4, 5]])
    assert nbh.topology_hash(a) == nbh.topology_hash(b) != nbh.topology_hash(c)


def test_hash_identifies_tips_by_row_not_by_position():
    from pylogeny_b200 import neighbourhood as nbh
    taxa = {"a": 0, "b": 1, "c": 2, "d": 3, "e": 4}
    d1, _, r1, _ = nbh.newick_to_desc("((a,b),(c,(d,e)));", taxa)
    d2, _, r2, _ = nbh.newick_to_desc("(((e,d),c),(b,a));", taxa)
    d3, _, r3, _ = nbh.newick_to_desc("((a,c),(b,(d,e)));", taxa)
    assert nbh.topology_hash(d1, r1) == nbh.topology_hash(d2, r2) != nbh.topology_hash(d3, r3)
