"""The reference's Python entry points on the device, end to end (no oracle, no monkeypatching):
pylogeny_b200.parsimony.profile_set -> pylogeny_b200.scoring.getParsimony* -> C-ABI -> CUDA, on the
reference's own fixture (tests/al.fasta re-encoded in tests/golden/al_alignment.json), against the
known answers of SURVEY.md 8c that tests/golden/make_golden.py produced by running the
reference's code (scoring.py:93-155, parsimony.py:11-145, landscape.py:689-761, heuristic.py:65-105)."""
import pytest

from conftest import load_golden

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def ali():
    from pylogeny_b200 import alignment
    g = load_golden("al_alignment.json")
    a = alignment.alignment(names=g["taxa"], seqs=g["seqs"])
    yield a
    a.close()


def test_scoring_entry_points_named_trees(ali, al_fitch, al_profiles):
    from pylogeny_b200 import parsimony, scoring
    named = {c["name"]: c for c in al_fitch["named"]}
    want = {"caterpillar_file_order": 1167, "balanced": 1296, "caterpillar_sorted": 1396, "unrooted_trifurcation": 1103}
    prof = parsimony.profile_set(ali)
    assert [str(p) for p in prof.profiles] == al_profiles["profiles"] and list(prof.weights) == al_profiles["weights"]
    for name, cost in want.items():
        nw = named[name]["newick"]
        assert named[name]["cost"] == cost
        assert scoring.getParsimonyFromProfiles(nw, prof) == cost          # scoring.py:123-137
        assert scoring.getParsimony(nw, ali) == cost                        # scoring.py:93-105 (rebuilds the profiles)

    class Topo(object):  # what getParsimony*ForTopology use of a topology (scoring.py:107-121, 139-155)
        def __init__(self, s):
            self.s = s

        def toNewick(self):
            return self.s

    nw = named["rooted_at_lowest_leaf"]["newick"]
    assert scoring.getParsimonyFromProfilesForTopology(Topo(nw), prof) == named["rooted_at_lowest_leaf"]["cost"]
    assert scoring.getParsimonyForTopology(Topo(nw), ali) == named["rooted_at_lowest_leaf"]["cost"]
    # the deprecated names of parsimony.py:147-189 (binary trees)
    assert parsimony.fitch_cost(Topo(named["balanced"]["newick"]), prof) == 1296
    assert parsimony.fitch(Topo(named["caterpillar_file_order"]["newick"]), ali) == 1167


def test_scoring_replays_the_reference_greedy_run(ali, al_greedy, al_fitch):
    """All 306 scoring calls of the unchanged reference's parsimonyGreedy run, one
    getParsimonyFromProfiles call each (as landscape.py:750 makes them), then in one batch; and the
    reference's 120 SPR moves of the start tree."""
    from pylogeny_b200 import parsimony, scoring
    prof = parsimony.profile_set(ali)
    assert len(al_greedy["calls"]) == 306
    got = [scoring.getParsimonyFromProfiles(nw, prof) for nw, _ in al_greedy["calls"]]
    assert got == [c for _, c in al_greedy["calls"]]
    assert scoring.getParsimonyFromProfilesBatch([nw for nw, _ in al_greedy["calls"]], prof) == got
    moves = al_fitch["spr_moves"]
    assert scoring.getParsimonyFromProfilesBatch([m["newick"] for m in moves], prof) == [m["cost"] for m in moves]
    assert min(got) == 1135 and got[0] == 1167


def test_profile_cache_follows_content(ali):
    """scoring.py:134-137 rebuilds the profile list per call; the resident copy must notice edits."""
    from pylogeny_b200 import parsimony, scoring
    prof = parsimony.profile_set(ali)
    nw = load_golden("al_fitch.json")["named"][0]["newick"]
    base = scoring.getParsimonyFromProfiles(nw, prof)
    prof.weights[0] += 5                       # in-place reweighting (bootstrap)
    bumped = scoring.getParsimonyFromProfiles(nw, prof)
    prof.weights[0] -= 5
    assert scoring.getParsimonyFromProfiles(nw, prof) == base
    assert bumped >= base


def test_batched_explore_on_device(ali, al_fitch, al_greedy):
    """BatchedExploreMixin.exploreTree (landscape.py:689-761's contract) on the device: the 90
    distinct topologies of the reference's 120 SPR moves with the reference's scores, the explored
    flag, the edges, and a greedy climb that follows the golden path 1167 -> 1149 -> 1138 -> 1135."""
    from minilandscape import MiniLandscape, MiniTree
    from pylogeny_b200 import explore, neighbourhood as nbh

    class Batched(explore.BatchedExploreMixin, MiniLandscape):
        pass

    names = ali.getTaxa()
    start = [c for c in al_fitch["named"] if c["name"] == "caterpillar_file_order"][0]
    ls = Batched(ali, MiniTree(start["newick"]))
    taxa = ls.parsimony_profiles.taxa
    nbrs = ls.exploreTree(ls.root)
    by_topo = {}
    for m in al_fitch["spr_moves"]:
        d, _, rows, _ = nbh.newick_to_desc(m["newick"], taxa)
        by_topo[nbh.topology_hash(d, rows)] = m["cost"]
    assert len(by_topo) == 90 and len(nbrs) == 90
    got = {ls._b200_hash_of(ls.getTree(j).getNewick(), taxa): ls.getTree(j).score[1] for j in nbrs}
    assert got == by_topo
    assert all(ls.graph.has_edge(ls.root, j) for j in nbrs) and all(ls.getTree(j).origin == 'SPR' for j in nbrs)
    assert ls.getNode(ls.root)['explored'] and ls.exploreTree(ls.root) == []
    # greedy climb (heuristic.py:65-105): move to the best neighbour while it improves
    ls.getTree(ls.root).score = (None, start["cost"])
    cur, path, seen = ls.root, [start["cost"]], [ls.root]
    while True:
        ls.exploreTree(cur)
        cand = sorted((j for j in ls.graph.neighbors(cur) if j not in seen), key=lambda j: ls.getTree(j).score[1])
        if not cand or ls.getTree(cand[0]).score[1] >= path[-1]:
            break
        cur = cand[0]
        seen.append(cur)
        path.append(ls.getTree(cur).score[1])
    assert path == al_greedy["path_scores"] == [1167, 1149, 1138, 1135]
    hashes = [ls._b200_hash_of(ls.getTree(j).getNewick(), taxa) for j in ls.graph.nodes]
    assert len(set(hashes)) == len(hashes) and len(ls) <= al_greedy["n_trees"]
