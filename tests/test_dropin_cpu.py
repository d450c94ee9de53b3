"""`landscape.py` and `heuristic.py` drop in unchanged: the reference's own modules (loaded from
/root/reference under the py3 shim -- build container only) run on top of pylogeny_b200.scoring /
pylogeny_b200.parsimony and reproduce the golden parsimonyGreedy run.

No GPU here, so the device call at the very bottom (FitchAlignment.score_newicks) is answered by
the CPU oracle -- test wiring only; the GPU replay of the same 306 scoring calls is
tests/test_fitch_gpu.py::test_reference_greedy_calls_golden."""
import types

import pytest

import oracle
import refshim
from conftest import load_golden

pytestmark = pytest.mark.skipif(not refshim.available(), reason="needs /root/reference (build container)")


class _OracleBackedAlignment(object):
    def __init__(self, profiles, weights):
        self.profiles, self.weights = list(profiles), list(weights)
        self.n_profiles = len(profiles)
        self.calls = 0

    def score_newicks(self, newicks, taxa):
        x = oracle.FitchInputs(self.profiles, self.weights, taxa)
        self.calls += len(newicks)
        return [oracle.fitch_cost(s, inputs=x) for s in newicks]

    def close(self):
        pass


def test_reference_landscape_and_greedy_on_our_scoring(monkeypatch, al_greedy):
    from pylogeny_b200 import alignment, fitch, parsimony, scoring
    made = []

    def factory(profiles, weights):
        made.append(_OracleBackedAlignment(profiles, weights))
        return made[-1]

    monkeypatch.setattr(fitch, "FitchAlignment", factory)
    g = load_golden("al_alignment.json")
    ali = alignment.alignment(names=g["taxa"], seqs=g["seqs"])
    stubs = {"scoring": scoring, "parsimony": parsimony, "alignment": alignment,
             "executable": types.ModuleType("executable")}
    stubs["executable"].raxml = None
    try:
        mods = refshim.install(stubs, want=("base", "newick", "rearrangement", "tree", "landscape", "heuristic"))
        names = g["taxa"]
        cat = names[-1]
        for n in reversed(names[:-1]):
            cat = "(%s,%s)" % (n, cat)
        ls = mods["landscape"].landscape(ali, starting_tree=mods["tree"].tree(cat + ";"), root=True, operator="SPR")
        h = mods["heuristic"].parsimonyGreedy(ls, ls.getNode(ls.root))
        h.explore()
        assert [n["tree"].score[1] for n in h.getPath()] == al_greedy["path_scores"] == [1167, 1149, 1138, 1135]
        assert len(ls) == al_greedy["n_trees"] == 306
        assert h.getBestTree()["tree"].getNewick() == al_greedy["best_newick"]
        # one resident alignment for the whole landscape, one scoring call per tree
        assert len(made) == 1 and made[0].calls == 306
    finally:
        refshim.uninstall()
