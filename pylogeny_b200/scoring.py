"""Mirror of the reference's scoring.py entry points (scoring.py:41-155) on the B200 engine.

    getParsimony(newick, alignment)                       scoring.py:93-105
    getParsimonyForTopology(topo, alignment)              scoring.py:107-121
    getParsimonyFromProfiles(newick, profiles)            scoring.py:123-137
    getParsimonyFromProfilesForTopology(topology, profs)  scoring.py:139-155
    getLogLikelihood(tree, alignment, updateBranchLengths) scoring.py:41-75

plus the batched entry points the landscape should call instead of one call per
neighbour (landscape.py:724-756): getParsimonyFromProfilesBatch.

Differences from the reference, all on the host side:
  * the profile strings are marshalled and uploaded ONCE per profile_set and stay
    resident in HBM (the reference rebuilds the list per tree, scoring.py:134,
    and fitch.cpp:293-309 copies it per call);
  * getParsimonyForTopology works (the reference raises NameError at :121).
"""
from . import fitch, parsimony


def _resident(profiles):
    """Device-resident alignment cached on the profile_set, keyed on its CONTENT (the reference
    rebuilds the profile list and copies the weights on every call, scoring.py:134-137, so editing
    either in place -- bootstrap reweighting -- must take effect here too).  The key costs O(P)
    cheap operations per call: Python caches each string's hash inside the string object."""
    prolist = [str(x) for x in profiles.profiles]
    weights = [int(w) for w in profiles.weights]
    key = (hash(tuple(prolist)), hash(tuple(weights)), len(prolist))
    cache = getattr(profiles, "_b200_cache", None)
    if cache is None or cache[0] != key:
        if cache is not None:
            cache[1].close()
        cache = (key, fitch.FitchAlignment(prolist, weights))
        profiles._b200_cache = cache
    return cache[1]


def invalidate(profiles):
    """Drop the device-resident copy of a profile_set (it is rebuilt on the next scoring call)."""
    cache = getattr(profiles, "_b200_cache", None)
    if cache is not None:
        cache[1].close()
        profiles._b200_cache = None


def getParsimonyFromProfiles(newick, profiles):
    """Acquire parsimony for a Newick string given a parsimony.profile_set. Returns an integer."""
    return int(_resident(profiles).score_newicks([newick], profiles.taxa)[0])


def getParsimonyFromProfilesBatch(newicks, profiles):
    """Parsimony of many Newick strings in one engine call. Returns a list of integers."""
    return [int(x) for x in _resident(profiles).score_newicks(list(newicks), profiles.taxa)]


def getParsimonyFromProfilesForTopology(topology, profiles):
    return getParsimonyFromProfiles(topology.toNewick(), profiles)


def getParsimony(newick, alignment):
    profiles = parsimony.profile_set(alignment)
    return getParsimonyFromProfiles(newick, profiles)


def getParsimonyForTopology(topo, alignment):
    profiles = parsimony.profile_set(alignment)
    return getParsimonyFromProfiles(topo.toNewick(), profiles)


def getLogLikelihoodBatch(trees, alignment, updateBranchLengths=True):
    """getLogLikelihood for a list of trees on one alignment in a single engine call (all trees are
    smoothed in lockstep on the device: ~0.6 ms per 64-taxon x 10k-site tree against ~2.2 ms through one
    handle per tree).  Returns a list
    of floats (None where the reference path would have returned None for that tree)."""
    from . import libpllWrapper as W, pll
    if not trees:
        return []
    model = pll.model_for(alignment)
    newicks = [t.toTopology().toUnrootedNewick() for t in trees]
    try:
        scores, optimised = W.getLogLikelihoodBatch(alignment.getPhylip(), newicks, model.getFileName())
    except Exception:
        return [getLogLikelihood(t, alignment, updateBranchLengths) for t in trees]  # isolate the failing tree
    if updateBranchLengths:
        for t, nw in zip(trees, optimised):
            try:
                t.updateNewick(nw.strip(), reroot=True)
            except ValueError:
                pass
    return [sc if sc else None for sc in scores]


def getLogLikelihood(tree, alignment, updateBranchLengths=True):
    """Log-likelihood through the libpllWrapper replacement (scoring.py:41-75): returns a float,
    or None when the model could not be built or scored."""
    from . import pll
    topo = tree.toTopology()
    try:
        p = pll.dataModel(topo, alignment)
    except Exception:
        return None
    try:
        sc = p.getLogLikelihood()
    except Exception:
        p.close()
        return None
    if updateBranchLengths:
        try:
            tree.updateNewick(p.getNewickString().strip(), reroot=True)
        except ValueError:
            pass
    p.close()
    return sc
