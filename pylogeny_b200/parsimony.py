"""Mirror of the reference's parsimony.py: column -> canonical pattern ("profile")
compression with integer weights.

Same public surface as /root/reference/pylogeny/parsimony.py:11-145
(`profile_set` with .alignment/.numSites/.taxa/.profiles/.weights/.sites, weight(),
get(), getForTaxa(); `site_profile` with .vector/.alphabet, ==, str()), the same
first-occurrence pattern order and the same weights, but the O(S*P) list scan of
_constructSet (parsimony.py:39-46) is a dictionary lookup and, when the alignment
exposes `as_matrix()` (pylogeny_b200.alignment does), the canonical relabelling of
_buildVector (parsimony.py:132-145) is vectorised over all columns at once.

The deprecated entry points `fitch_cost(topology, profiles)` and `fitch(topology, alignment)`
(parsimony.py:147-189, a pure-Python Fitch for binary trees) keep their names and signatures
but score on the device like everything else (there is no CPU scorer in this package).
"""
import numpy as np


class site_profile(object):
    """One alignment column as a canonical digit string (parsimony.py:97-145)."""

    def __init__(self, alignment, site, _vector=None, _alphabet=None):
        self.alignment = alignment
        self.site = site
        self.alphabet = _alphabet
        self.vector = ''
        if _vector is not None:
            self.vector = _vector
        else:
            self._buildVector()

    def __eq__(self, o):
        if o is None:
            return False
        return self.vector == o.vector

    def __ne__(self, o):
        return not self.__eq__(o)

    def __hash__(self):
        return hash(self.vector)

    def __str__(self):
        return self.vector

    def _buildVector(self):
        sitesli = self.alignment.data.sequenceSlice(self.site)
        comps, out = {}, []
        for it in sitesli:
            if it not in comps:
                comps[it] = len(comps)
            out.append(str(comps[it]))  # str(idx): two chars once idx >= 10
        self.vector = ''.join(out)
        self.alphabet = comps


def canonical_columns(mat):
    """mat: uint8 [n_rows, n_sites] -> int16 [n_rows, n_sites] rank of first appearance per column."""
    n, S = mat.shape
    label = np.full((256, S), -1, dtype=np.int16)
    count = np.zeros(S, dtype=np.int16)
    cols = np.arange(S)
    out = np.empty((n, S), dtype=np.int16)
    for r in range(n):
        sym = mat[r]
        cur = label[sym, cols]
        new = cur < 0
        if new.any():
            label[sym[new], cols[new]] = count[new]
            count[new] += 1
            cur = label[sym, cols]
        out[r] = cur
    return out


class profile_set(object):
    """All site profiles of an alignment, de-duplicated with weights (parsimony.py:11-95)."""

    def __init__(self, alignment):
        self.alignment = alignment
        self.numSites = alignment.getSize()
        self.taxa = {}
        self.profiles = []
        self.weights = []
        self.sites = []
        self._constructSet()
        self._buildTaxaDict()

    def _constructSet(self):
        if hasattr(self.alignment, "as_matrix"):
            return self._constructSetVectorised()
        index = {}
        for site in range(self.numSites):
            pro = site_profile(self.alignment, site)
            self.sites.append(pro)
            at = index.get(pro.vector)
            if at is not None:
                self.weights[at] += 1
            else:
                index[pro.vector] = len(self.profiles)
                self.profiles.append(pro)
                self.weights.append(1)

    def _constructSetVectorised(self):
        mat = self.alignment.as_matrix()  # uint8 [n_rows, n_sites]
        ranks = canonical_columns(mat)
        S = ranks.shape[1]
        if S == 0:
            return
        cols = np.ascontiguousarray(ranks.T)  # [S, n]
        if cols.max() < 10:
            keys = np.ascontiguousarray((cols + 48).astype(np.uint8))
            void = keys.view(np.dtype((np.void, keys.shape[1]))).ravel()
            _, first, inverse, counts = np.unique(void, return_index=True, return_inverse=True, return_counts=True)
            order = np.argsort(first, kind="stable")  # first-occurrence order (parsimony.py:42-46)
            rank_of = np.empty_like(order)
            rank_of[order] = np.arange(len(order))
            vectors = [keys[first[u]].tobytes().decode("ascii") for u in order]
            self.weights = [int(counts[u]) for u in order]
            site_to = rank_of[inverse.ravel()]
        else:  # >= 10 symbols in some column: str(idx) becomes multi-char, take the generic route
            index, vectors, site_to = {}, [], np.empty(S, dtype=np.int64)
            for s in range(S):
                v = ''.join(map(str, cols[s].tolist()))
                at = index.get(v)
                if at is None:
                    at = index[v] = len(vectors)
                    vectors.append(v)
                    self.weights.append(0)
                self.weights[at] += 1
                site_to[s] = at
        self.profiles = [site_profile(self.alignment, -1, _vector=v) for v in vectors]
        firsts = {}
        for s, u in enumerate(site_to.tolist()):
            if u not in firsts:
                firsts[u] = s
                self.profiles[u].site = s
            self.sites.append(site_profile(self.alignment, s, _vector=self.profiles[u].vector))
        self._dense = None

    def _buildTaxaDict(self):
        tax = self.alignment.getTaxa()
        for t in range(len(tax)):
            self.taxa[tax[t]] = t

    def __len__(self):
        return len(self.profiles)

    def weight(self, val):
        return self.weights[val]

    def get(self, val):
        return self.profiles[val]

    def getForTaxa(self, val, tax):
        return self.profiles[val].vector[self.taxa[tax]]


def fitch_cost(topology, profiles):
    """parsimony.fitch_cost (parsimony.py:147-176, deprecated there): Fitch length of a topology
    over a profile_set.  The reference walks `topology.getRoot()` in Python and reads only the first
    two children of every node; here the topology's Newick string (`toNewick()`) goes to the CUDA
    scorer, which equals it on binary trees (and follows fitch.cpp's k-ary rule elsewhere).  The
    result is fitch.cpp's 32-bit total; the Python original does not wrap."""
    from . import scoring
    return scoring.getParsimonyFromProfiles(topology.toNewick(), profiles)


def fitch(topology, alignment):
    """parsimony.fitch (parsimony.py:178-189, deprecated there): profile the alignment, then
    fitch_cost; None for an alignment without columns."""
    profiles = profile_set(alignment)
    if len(profiles) == 0:
        return None
    return fitch_cost(topology, profiles)
