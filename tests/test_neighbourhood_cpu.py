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


def _splits(desc, n):
    """Set of non-trivial splits (frozenset of the side without tip 0) of a rooted descriptor's unrooted tree."""
    below = {i: frozenset([i]) for i in range(n)}
    for i, (a, b) in enumerate(desc):
        below[n + i] = below[int(a)] | below[int(b)]
    full = frozenset(range(n))
    out = set()
    for k, s in below.items():
        s = s if 0 not in s else full - s
        if 1 < len(s) < n - 1:
            out.add(s)
    return frozenset(out)


def _brute_tbr(desc, n):
    """All TBR neighbours by brute force on split sets: cut every edge, re-root both parts on every
    edge, join.  Trees as nested tuples over tip ids."""
    import itertools
    nodes = {i: i for i in range(n)}
    adj = {}
    def link(x, y):
        adj.setdefault(x, []).append(y); adj.setdefault(y, []).append(x)
    root = n + len(desc) - 1
    for i, (a, b) in enumerate(desc[:-1]):
        link(n + i, int(a)); link(n + i, int(b))
    link(int(desc[-1][0]), int(desc[-1][1]))
    edges = [(x, y) for x in adj for y in adj[x] if x < y]

    def component(start, banned):
        seen, st = {start}, [start]
        while st:
            x = st.pop()
            for y in adj[x]:
                if (x, y) != banned and (y, x) != banned and y not in seen:
                    seen.add(y); st.append(y)
        return seen

    def rooted(x, frm, comp, banned):
        kids = [y for y in adj[x] if y != frm and y in comp and (x, y) != banned and (y, x) != banned]
        if not kids:
            return x
        subs = [rooted(y, x, comp, banned) for y in kids]
        return subs[0] if len(subs) == 1 else tuple(subs)

    def rootings(comp, banned):
        es = [(x, y) for (x, y) in edges if x in comp and y in comp and (x, y) != banned]
        if not es:
            return [next(iter(comp))] if len(comp) == 1 else []
        return [(rooted(x, y, comp, banned), rooted(y, x, comp, banned)) for x, y in es]

    def leafset(t):
        return frozenset([t]) if not isinstance(t, tuple) else frozenset().union(*[leafset(c) for c in t])

    def splits_of(t, acc):
        if isinstance(t, tuple):
            for c in t:
                splits_of(c, acc)
            acc.add(leafset(t))

    full = frozenset(range(n))
    out = set()
    for e in edges:
        c1, c2 = component(e[0], e), component(e[1], e)
        for r1 in rootings(c1, e):
            for r2 in rootings(c2, e):
                acc = set()
                splits_of((r1, r2), acc)
                canon = frozenset(s if 0 not in s else full - s for s in acc)
                out.add(frozenset(s for s in canon if 1 < len(s) < n - 1))
    return out


@pytest.mark.parametrize("n", [4, 5, 6, 7, 9])
def test_tbr_enumeration_vs_bruteforce(n):
    from pylogeny_b200 import neighbourhood as nbh
    rng = np.random.default_rng(100 + n)
    d = random_desc(rng, n)
    brute = _brute_tbr(d, n)
    brute.discard(_splits(d, n))
    M = nbh.neighbourhood_size(d, nbh.MOVE_TBR)
    br = rng.exponential(0.1, d.shape)
    descs, brs, hashes = nbh.neighbour_trees(d, br, None, nbh.MOVE_TBR)
    got = {_splits(x, n) for x in descs}
    assert len(got) == M == len(set(hashes.tolist()))
    assert got == brute
    for i in range(M):
        assert nbh.topology_hash(descs[i]) == int(hashes[i])
        assert abs(brs[i].sum() - br.sum()) < 1e-12
    # SPR neighbours are a subset of TBR neighbours
    sd, _, sh = nbh.neighbour_trees(d, None, None, nbh.MOVE_SPR)
    assert set(sh.tolist()) <= set(hashes.tolist())


@pytest.mark.parametrize("n,radius", [(6, 1), (7, 2), (9, 3), (12, 2), (12, 4), (12, 30)])
def test_radius_limited_spr_vs_split_distance(n, radius):
    """getSPRMovesInDistance's maxdist (libpllWrapper.c:241-262): the radius-r neighbourhood is exactly
    the SPR neighbours that differ from the base tree in at most r splits (an SPR across d internal
    nodes replaces d splits)."""
    from pylogeny_b200 import neighbourhood as nbh
    rng = np.random.default_rng(500 + n + radius)
    d = random_desc(rng, n)
    base = _splits(d, n)
    descs, _, hashes = nbh.neighbour_trees(d, None, None, nbh.MOVE_SPR)
    want = {int(h) for x, h in zip(descs, hashes) if len(_splits(x, n) - base) <= radius}
    mt = nbh.MOVE_SPR_WITHIN(radius)
    M = nbh.neighbourhood_size(d, mt)
    rd, _, rh = nbh.neighbour_trees(d, None, None, mt)
    assert M == len(rh) == len(set(rh.tolist()))
    assert set(int(h) for h in rh) == want
    for x, h in zip(rd, rh):
        assert nbh.topology_hash(x) == int(h)
    if radius == 1:
        assert M == nbh.neighbourhood_size(d, nbh.MOVE_NNI)
    if radius >= n:
        assert M == nbh.neighbourhood_size(d, nbh.MOVE_SPR)


@pytest.mark.parametrize("n", [4, 5, 9, 40, 130])
def test_single_moves_without_enumeration(n):
    """neighbour_trees(sel=few) expands only the searches that hold the requested ids (what a climb
    step uses to materialise its winner); it must agree with the full enumeration: descriptors,
    branch lengths and topology hashes."""
    from pylogeny_b200 import neighbourhood as nbh
    rng = np.random.default_rng(3 * n + 1)
    d = random_desc(rng, n)
    br = rng.exponential(0.1, d.shape)
    perm = rng.permutation(n).astype(np.int32)
    M = nbh.neighbourhood_size(d)
    sel_all = np.arange(M, dtype=np.int64)
    if M <= 64:  # force the enumeration path with a padded selection
        sel_all = np.concatenate([sel_all, np.zeros(65 - M, dtype=np.int64)])
    full_d, full_b, full_h = nbh.neighbour_trees(d, br, sel_all, tip_rows=perm)
    pick = np.unique(np.concatenate([[0, M - 1], rng.integers(0, M, 20)])).astype(np.int64)
    for chunk in (pick[:1], pick):
        few_d, few_b, few_h = nbh.neighbour_trees(d, br, chunk, tip_rows=perm)
        assert few_d.tolist() == full_d[chunk].tolist()
        assert few_b.tolist() == full_b[chunk].tolist()
        assert few_h.tolist() == full_h[chunk].tolist()
    with pytest.raises(Exception):
        nbh.neighbour_trees(d, br, np.array([M], dtype=np.int64))
