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


def test_batched_explore_matches_reference_run(monkeypatch, al_greedy, al_fitch):
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
    try:
        mods = refshim.install(stubs, want=("base", "newick", "rearrangement", "tree", "landscape", "heuristic"))

        class batchedLandscape(explore.BatchedExploreMixin, mods["landscape"].landscape):
            pass

        names = g["taxa"]
        cat = names[-1]
        for n in reversed(names[:-1]):
            cat = "(%s,%s)" % (n, cat)
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
    finally:
        refshim.uninstall()
