"""Batched exploreTree mixin on the reference's landscape/heuristic (py3 shim, build container
only).  Host logic under test: move -> Newick materialisation, canonical-hash dedupe, graph
wiring.  The device call (neighbourhood.score_parsimony) is answered by the CPU oracle here; its
GPU parity is tests/test_neighbourhood_gpu.py."""
import types

import numpy as np
import pytest

import oracle
import refshim
from conftest import load_golden

pytestmark = pytest.mark.skipif(not refshim.available(), reason="needs /root/reference (build container)")


@pytest.fixture
def ref_world(monkeypatch):
    """The reference's landscape / tree / heuristic modules (py3 shim) over the product's scoring,
    parsimony and alignment modules, the device call answered by the CPU oracle."""
    from pylogeny_b200 import alignment, explore, fitch, neighbourhood as nbh, parsimony, scoring
    g = load_golden("al_alignment.json")
    ali = alignment.alignment(names=g["taxa"], seqs=g["seqs"])

    class FakeAln(object):
        def __init__(self, profiles, weights):
            self.x = None
            self.profiles, self.weights = profiles, weights

        def score_newicks(self, newicks, taxa):
            x = oracle.FitchInputs(self.profiles, self.weights, taxa)
            return [oracle.fitch_cost(s, inputs=x) for s in newicks]

    def fake_score(aln, desc, rows, move_type=nbh.MOVE_SPR, shard=(0, 1)):
        descs, _, _ = nbh.neighbour_trees(desc, None, None, move_type)
        labels = ["L%d" % i for i in range(len(rows))]
        taxa = {labels[i]: int(rows[i]) for i in range(len(rows))}
        x = oracle.FitchInputs(aln.profiles, aln.weights, taxa)
        from util import desc_to_newick
        costs = np.array([oracle.fitch_cost(desc_to_newick(d, labels), inputs=x) for d in descs], dtype=np.int32)
        return np.zeros((len(costs), 4), np.int32), costs

    monkeypatch.setattr(fitch, "FitchAlignment", FakeAln)
    monkeypatch.setattr(nbh, "score_parsimony", fake_score)
    stubs = {"scoring": scoring, "parsimony": parsimony, "alignment": alignment,
             "executable": types.ModuleType("executable")}
    stubs["executable"].raxml = None
    mods = refshim.install(stubs, want=("base", "newick", "rearrangement", "tree", "landscape", "heuristic"))

    class batchedLandscape(explore.BatchedExploreMixin, mods["landscape"].landscape):
        pass

    names = g["taxa"]
    cat = names[-1]
    for n in reversed(names[:-1]):
        cat = "(%s,%s)" % (n, cat)
    w = types.SimpleNamespace(mods=mods, ali=ali, names=names, start=cat + ";", batched=batchedLandscape,
                              plain=mods["landscape"].landscape, taxa={n: i for i, n in enumerate(names)})
    w.make = lambda cls, newick=None: cls(ali, starting_tree=mods["tree"].tree(newick or w.start), root=True, operator="SPR")
    w.hash_of = lambda nw: nbh.topology_hash(*[nbh.newick_to_desc(nw, w.taxa)[k] for k in (0, 2)])
    try:
        yield w
    finally:
        refshim.uninstall()


def _first_occurrences(seq):
    out, seen = [], set()
    for x in seq:
        if x not in seen:
            seen.add(x)
            out.append(x)
    return out


@pytest.mark.parametrize("move", ["SPR", "NNI"])
def test_locks_and_reference_order(ref_world, move):
    """landscape.locks (rearrangement.py:228-240, landscape.py:714-718) through the batched explore: for
    every branch of the starting tree locked in turn, and for two locks at once, the mixin reaches
    exactly the topologies the UNCHANGED reference reaches, with the same scores, in the order of
    their first occurrence in the reference's own enumeration (reference_order is implied by locks)."""
    w = ref_world
    typ = getattr(w.mods["rearrangement"], "TYPE_" + move)
    n_br = len(w.make(w.plain).getTree(0).toTopology().getBranches())
    cases = [()] + [(k,) for k in range(n_br)] + [(2, 9), (4, 12), (3, 5, 11)]
    sizes = set()
    for case in cases:
        runs = []
        for cls in (w.plain, w.batched):
            ls = w.make(cls)
            ls.reference_order = True   # (read by the mixin only)
            for k in case:
                br = ls.getTree(ls.root).toTopology().getBranches()[k]
                assert ls.lockBranchFoundInTree(ls.root, br) is not None
            nb = ls.exploreTree(ls.root, typ)
            runs.append([(w.hash_of(ls.getTree(j).getNewick()), ls.getTree(j).score[1], ls.getTree(j).origin) for j in nb])
        ref, got = runs
        assert len(set(h for h, _, _ in got)) == len(got)              # one node per topology
        assert _first_occurrences(ref) == got, case                    # same topologies, scores, origin, ORDER
        sizes.add(len(got))
    if move == "SPR":
        assert len(sizes) > 3                                          # the locks did restrict the neighbourhoods
    else:
        # the reference's NNI producer lists no move at all (rearrangement.py:609-612: every candidate
        # destination is in the branch's own forbidden list; tests/golden n_nni_reference = 0), so in
        # reference_order / lock mode the mixin inserts none either; without them it scores the real
        # 2(n-3) NNI neighbours (tests/test_dropin_gpu.py)
        assert sizes == {0}


def test_violating_tree_steps_only_to_non_violating_neighbours(ref_world):
    """landscape.py:743-746: a tree that does not hold a locked bipartition has no branch to lock; its
    neighbours that violate the lock as well are skipped."""
    w = ref_world
    runs = []
    for cls in (w.plain, w.batched):
        ls = w.make(cls)
        ls.reference_order = True
        other = w.mods["tree"].tree("((%s,%s),(%s,%s),(%s,(%s,(%s,%s))));" % tuple(w.names[k] for k in (0, 5, 2, 7, 1, 3, 4, 6)))
        topl = other.toTopology()
        mine = ls.getTree(ls.root).toTopology()
        lock = [b for b in (w.mods["tree"].bipartition(topl, br) for br in topl.getBranches())
                if not mine.getBranchFromBipartition(b)][0]       # a split of `other` the start tree lacks
        ls.toggleLock(lock)
        assert ls._isViolating(ls.getTree(ls.root).toTopology())
        nb = ls.exploreTree(ls.root)
        runs.append([(w.hash_of(ls.getTree(j).getNewick()), ls.getTree(j).score[1]) for j in nb])
    assert _first_occurrences(runs[0]) == runs[1] and 0 < len(runs[1]) < 90


def test_reference_order_greedy_path_is_the_reference_path(ref_world, al_greedy):
    """With reference_order the greedy climb breaks ties as the reference does (heuristic.py:92-95
    sorted(...)[0] over nodes in insertion order): the same TREE at every step of the path, not only
    the same score, and every exploration inserts the distinct topologies in the reference's order."""
    w = ref_world
    paths = []
    for cls in (w.plain, w.batched):
        ls = w.make(cls)
        ls.reference_order = True
        h = w.mods["heuristic"].parsimonyGreedy(ls, ls.getNode(ls.root))
        h.explore()
        # (heuristic.py keeps the path in a list shared between instances: take this run's tail)
        mine = [(w.hash_of(nd["tree"].getNewick()), nd["tree"].score[1]) for nd in h.getPath()]
        paths.append(mine[-len(al_greedy["path_scores"]):])
        order = [w.hash_of(ls.getTree(j).getNewick()) for j in sorted(ls.graph.nodes)]
        paths.append(_first_occurrences(order))
    assert paths[0] == paths[2] and [s for _, s in paths[2]] == al_greedy["path_scores"]
    assert paths[1] == paths[3]                                        # whole landscape, node order


def test_batched_explore_matches_reference_run(ref_world, monkeypatch, al_greedy, al_fitch):
    from pylogeny_b200 import neighbourhood as nbh, scoring
    w = ref_world
    mods, ali, names, cat, batchedLandscape = w.mods, w.ali, w.names, w.start[:-1], w.batched
    if True:
        ls = batchedLandscape(ali, starting_tree=mods["tree"].tree(cat + ";"), root=True, operator="SPR")
        nbrs = ls.exploreTree(ls.root)
        scores = sorted(ls.getTree(j).score[1] for j in nbrs)
        # the reference's 120 moves = 90 topologies; compare score multisets per distinct topology
        by_topo = {}
        for m in al_fitch["spr_moves"]:
            d, _, rows, _ = nbh.newick_to_desc(m["newick"], {n: i for i, n in enumerate(names)})
            by_topo[nbh.topology_hash(d, rows)] = m["cost"]
        assert len(by_topo) == 90 and len(nbrs) == 90
        assert scores == sorted(by_topo.values())
        assert {ls._b200_hash_of(ls.getTree(j).getNewick(), ls.parsimony_profiles.taxa) for j in nbrs} == set(by_topo)
        assert ls.exploreTree(ls.root) == []                       # explored flag
        # likelihood for many nodes in one call: the engine call is stubbed, the wiring is under test
        calls = []

        def fake_batch(trees, alignment_, updateBranchLengths=True):
            calls.append(len(trees))
            return [-1000.0 - k for k in range(len(trees))]

        monkeypatch.setattr(scoring, "getLogLikelihoodBatch", fake_batch)
        got = ls.scoreLikelihoodFor(nbrs[:7])
        assert calls == [7] and got == [-1000.0 - k for k in range(7)]
        assert [ls.getTree(j).score[0] for j in nbrs[:7]] == got
        assert all(ls.getTree(j).score[1] is not None for j in nbrs[:7])      # parsimony kept
        assert ls.scoreLikelihoodFor(nbrs[:9])[:7] == got and calls == [7, 2]  # only the unscored ones go out
        assert ls.getVertex(nbrs[0]).scoreLikelihood() == got[0]               # the reference finds it cached
        # greedy climb over the batched landscape: same path scores as the unchanged reference run
        ls2 = batchedLandscape(ali, starting_tree=mods["tree"].tree(cat + ";"), root=True, operator="SPR")
        h = mods["heuristic"].parsimonyGreedy(ls2, ls2.getNode(ls2.root))
        h.explore()
        assert [n["tree"].score[1] for n in h.getPath()] == al_greedy["path_scores"]
        assert 1 < len(ls2) <= al_greedy["n_trees"]                # no string-duplicates of one topology
        hashes = [ls2._b200_hash_of(ls2.getTree(j).getNewick(), ls2.parsimony_profiles.taxa) for j in ls2.graph.nodes]
        assert len(set(hashes)) == len(hashes)
