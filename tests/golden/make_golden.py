#!/usr/bin/env python
"""Generate the committed golden fixtures from the REFERENCE's own code.

Run in the build container only (needs /root/reference):
    python tests/golden/make_golden.py
It executes
  * the reference's pure-Python modules (parsimony.profile_set, rearrangement.allSPR,
    landscape.exploreTree, heuristic.parsimonyGreedy) under the py3 shim in
    tests/refshim.py, and
  * the reference's compiled Fitch core (oracle/_ref/libfitch_ref.so, built by
    oracle/Makefile from /root/reference/pylogeny/fitch.cpp:6-272),
and writes small JSON fixtures next to this script.  The fixtures travel to the
GPU box; /root/reference does not.
"""
import json
import os
import random
import sys
import types

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import oracle  # noqa: E402
import refshim  # noqa: E402


def read_fasta(path):
    names, seqs = [], []
    with open(path) as fh:
        for line in fh:
            line = line.strip()
            if not line:
                continue
            if line.startswith(">"):
                names.append(line[1:].split()[0])
                seqs.append("")
            else:
                seqs[-1] += line
    return names, seqs


class _Data:
    def __init__(self, seqs):
        self.seqs = seqs

    def sequenceSlice(self, site):  # p4 Alignment.sequenceSlice: column as list of chars
        return [s[site] for s in self.seqs]


class StubAlignment:
    """What reference parsimony.py:25-53 and landscape.py:335-340 need from alignment.py."""

    def __init__(self, names, seqs):
        self.names, self.data = names, _Data(seqs)
        self.paths = {}

    def getSize(self):
        return len(self.data.seqs[0])

    def getNumSeqs(self):
        return len(self.names)

    def getTaxa(self):
        return list(self.names)

    def __len__(self):
        return self.getSize()


def make_scoring_module(counter):
    """scoring.getParsimonyFromProfiles restated from scoring.py:123-137 with the
    compiled reference core standing in for the py2 `fitch` extension."""
    m = types.ModuleType("scoring")

    def getParsimonyFromProfiles(newick, profiles):
        prolist = [str(x) for x in profiles.profiles]
        c = oracle.ref_fitch_cost(newick, prolist, profiles.weights, profiles.taxa)
        counter.append((newick, c))
        return c

    def getLogLikelihood(tree, alignment, updateBranchLengths=True):
        return None

    m.getParsimonyFromProfiles = getParsimonyFromProfiles
    m.getLogLikelihood = getLogLikelihood
    return m


def caterpillar(names):
    s = names[-1]
    for n in reversed(names[:-1]):
        s = "(%s,%s)" % (n, s)
    return s + ";"


def random_newick(rng, labels, multif=0.0, lengths=False):
    nodes = list(labels)
    rng.shuffle(nodes)

    def bl():
        return (":%s" % round(rng.uniform(0.01, 2.0), 4)) if lengths else ""

    while len(nodes) > 1:
        k = 2
        while k < len(nodes) and rng.random() < multif:
            k += 1
        picks = [nodes.pop(rng.randrange(len(nodes))) for _ in range(k)]
        nodes.append("(" + ",".join(p + bl() for p in picks) + ")")
    return nodes[0] + ";"


def main():
    if not refshim.available() or not oracle.ref_available():
        sys.exit("needs /root/reference and oracle/_ref (run `make -C oracle`)")
    names, seqs = read_fasta("/root/reference/tests/al.fasta")
    with open(os.path.join(HERE, "al_alignment.json"), "w") as fh:
        json.dump({"source": "reference tests/al.fasta re-encoded", "taxa": names, "seqs": seqs}, fh)

    calls = []
    stubs = {"scoring": make_scoring_module(calls), "alignment": types.ModuleType("alignment"),
             "executable": types.ModuleType("executable")}
    stubs["executable"].raxml = None
    mods = refshim.install(stubs, want=("base", "newick", "rearrangement", "tree", "parsimony", "landscape",
                                        "heuristic"))
    ali = StubAlignment(names, seqs)

    # --- profiles by the reference's own parsimony.profile_set (parsimony.py:11-53)
    prof = mods["parsimony"].profile_set(ali)
    profiles = [str(p) for p in prof.profiles]
    gold_prof = {"profiles": profiles, "weights": list(prof.weights), "taxa": dict(prof.taxa),
                 "n_sites": prof.numSites}
    with open(os.path.join(HERE, "al_profiles.json"), "w") as fh:
        json.dump(gold_prof, fh)
    inputs = oracle.FitchInputs(profiles, prof.weights, prof.taxa)

    # --- named trees
    srt = sorted(names)
    fixed = {
        "caterpillar_file_order": caterpillar(names),
        "balanced": "(((%s,%s),(%s,%s)),((%s,%s),(%s,%s)));" % tuple(names),
        "caterpillar_sorted": caterpillar(srt),
    }
    treemod, rearr = mods["tree"], mods["rearrangement"]
    start = treemod.tree(caterpillar(names))
    topo = start.toTopology()
    fixed["rooted_at_lowest_leaf"] = topo.toNewick()
    fixed["unrooted_trifurcation"] = topo.toUnrootedNewick()
    fitch_cases = [{"name": k, "newick": v, "cost": oracle.ref_fitch_cost(v, inputs=inputs)}
                   for k, v in fixed.items()]

    # --- the reference's own SPR enumeration (rearrangement.py:580-594), every move
    spr = []
    for mv in topo.allSPR():
        t = mv.toTree()
        nw = t.getNewick()
        spr.append({"newick": nw, "struct": t.getStructure(), "cost": oracle.ref_fitch_cost(nw, inputs=inputs)})
    with open(os.path.join(HERE, "al_fitch.json"), "w") as fh:
        json.dump({"named": fitch_cases, "spr_moves": spr, "n_nni_reference": len(topo.allNNI())}, fh)

    # --- parsimonyGreedy over the SPR landscape, reference landscape.py/heuristic.py unchanged
    del calls[:]
    ls = mods["landscape"].landscape(ali, starting_tree=treemod.tree(caterpillar(names)), root=True,
                                     operator="SPR")
    h = mods["heuristic"].parsimonyGreedy(ls, ls.getNode(ls.root))
    h.explore()
    path = [n["tree"].score[1] for n in h.getPath()]
    with open(os.path.join(HERE, "al_greedy.json"), "w") as fh:
        json.dump({"path_scores": path, "n_trees": len(ls), "calls": calls,
                   "best_newick": h.getBestTree()["tree"].getNewick()}, fh)

    # --- random known-answer cases for the compiled reference core
    rng = random.Random(20261019)
    cases = []
    for ci in range(60):
        n = rng.randint(3, 14)
        P = rng.randint(1, 48)
        nstates = rng.choice([2, 4, 5, 10])
        labels = ["T%d" % i for i in range(n)]
        taxa = {l: i for i, l in enumerate(labels)}
        if ci % 5 == 0:  # permuted taxon->column map
            perm = list(range(n))
            rng.shuffle(perm)
            taxa = {l: perm[i] for i, l in enumerate(labels)}
        profs = ["".join(str(rng.randrange(nstates)) for _ in range(n)) for _ in range(P)]
        if ci % 7 == 3:
            wts = [rng.choice([1, 2, 3, 1000003, 2 ** 31 - 1, 2 ** 33 + 5]) for _ in range(P)]
        else:
            wts = [rng.randint(1, 50) if rng.random() < 0.3 else 1 for _ in range(P)]
        nw = random_newick(rng, labels, multif=(0.4 if ci % 3 == 0 else 0.0), lengths=(ci % 2 == 0))
        cases.append({"newick": nw, "profiles": profs, "weights": wts, "taxa": taxa,
                      "cost": oracle.ref_fitch_cost(nw, profs, wts, taxa)})
    with open(os.path.join(HERE, "random_fitch.json"), "w") as fh:
        json.dump(cases, fh)
    print("profiles:", len(profiles), "sum w:", sum(prof.weights))
    print("named:", [(c["name"], c["cost"]) for c in fitch_cases])
    print("spr moves:", len(spr), "distinct structs:", len({s["struct"] for s in spr}),
          "min/max:", min(s["cost"] for s in spr), max(s["cost"] for s in spr))
    print("greedy path:", path, "trees:", len(ls), "calls:", len(calls))


if __name__ == "__main__":
    main()
