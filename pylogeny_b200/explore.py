"""Batched `exploreTree` for the reference's landscape class.

    class batchedLandscape(BatchedExploreMixin, landscape.landscape): pass

`landscape.exploreTree` (landscape.py:689-761) enumerates a neighbourhood in Python and scores one
tree per call.  The mixin overrides only that method: the whole neighbourhood goes to the engine
in one call (neighbourhood.score_parsimony: enumeration in C++, edge-insertion Fitch on the GPU),
each DISTINCT topology comes back once with its score, and only then are tree objects made.

Contract (checked in tests/test_explore_cpu.py against the golden run of the unchanged
reference): same set of neighbour topologies with the same parsimony scores, same greedy path
scores; fewer nodes where the reference stored string-duplicates of one topology
because identity here is the canonical topology hash, not the child-order-dependent string.
Tie rule: neighbours are inserted in the engine's enumeration order (pruned branch by node id,
then depth-first by target), so among equal scores `sorted(...)[0]` (heuristic.py:92-95) may pick
a different tree than the reference's enumeration order would.

Likelihood: `scoreLikelihoodFor(indices)` scores many nodes in one engine call (batched branch-length
smoothing); with `batch_likelihood = True` every explored neighbourhood is scored that way.

Falls back to the reference method when locks are set (rearrangement.py:228-240 semantics are not
reproduced here) or for move types the engine does not enumerate.
"""
from . import fitch, neighbourhood as nbh
from . import scoring


class BatchedExploreMixin(object):

    def _b200_hash_of(self, newick, taxa):
        d, _, rows, _ = nbh.newick_to_desc(newick, taxa)
        return nbh.topology_hash(d, rows)

    def _b200_sync_index(self, taxa):
        """Topology hash -> node index for every tree already in the landscape."""
        idx = getattr(self, "_b200_index", None)
        if idx is None:
            idx = self._b200_index = {}
            self._b200_seen = set()
        for i in list(self.graph.nodes):
            if i in self._b200_seen:
                continue
            self._b200_seen.add(i)
            idx.setdefault(self._b200_hash_of(self.getNode(i)['tree'].getNewick(), taxa), i)
        return idx

    def exploreTree(self, i, type=nbh.MOVE_SPR):
        if type not in (nbh.MOVE_SPR, nbh.MOVE_NNI) or getattr(self, "locks", None):
            return super(BatchedExploreMixin, self).exploreTree(i, type)
        p = self.parsimony_profiles
        node = self.getNode(i)
        if node['explored']:
            return list()
        tre = node['tree']
        treecls = tre.__class__
        newick = tre.toTopology().toNewick()            # rooted at the lowest-order leaf
        desc, brlen, rows, tips = nbh.newick_to_desc(newick, p.taxa)
        aln = scoring._resident(p)
        moves, costs = nbh.score_parsimony(aln, desc, rows, type)
        descs, brs, hashes = nbh.neighbour_trees(desc, brlen, None, type, rows)
        index = self._b200_sync_index(p.taxa)
        origin = 'SPR' if type == nbh.MOVE_SPR else 'NNI'
        neighbors = list()
        for j in range(len(costs)):
            h = int(hashes[j])
            found = index.get(h)
            if found is not None:
                if found != i and self.graph.has_node(found) and not self.graph.has_edge(found, i):
                    self.graph.add_edge(found, i)
                    self.getEdge(found, i)['weight'] = self.defaultWeight
                continue
            t = treecls(nbh.desc_to_newick(descs[j], tips, brs[j]), check=True)
            if self.findTreeTopologyByStructure(t.getStructure()) is not None:
                continue  # (string already present under another hash: cannot happen for one topology)
            t.score = (None, int(costs[j]))
            t.origin = origin
            k = self._newNode(t, score=True)
            index[h] = k
            self._b200_seen.add(k)
            neighbors.append(k)
            self.graph.add_edge(i, k)
            self.getEdge(i, k)['weight'] = self.defaultWeight
        node['explored'] = True
        if getattr(self, "batch_likelihood", False):
            self.scoreLikelihoodFor(neighbors)
        return neighbors

    def scoreLikelihoodFor(self, indices):
        """`vertex.scoreLikelihood()` (landscape.py:1302-1314) for many nodes in ONE engine call:
        every tree without a likelihood gets the reference's per-tree treatment (3 smoothing passes
        from default branch lengths, lnL, Newick updated) through scoring.getLogLikelihoodBatch.
        likelihoodGreedy / smoothGreedy (heuristic.py:107-260) then find the scores in place.  Set
        `batch_likelihood = True` on the landscape to do this for every explored neighbourhood."""
        nodes = [self.getNode(i) for i in indices]
        todo = [nd for nd in nodes if nd['tree'].score[0] is None]
        scores = scoring.getLogLikelihoodBatch([nd['tree'] for nd in todo], self.alignment)
        for nd, sc in zip(todo, scores):
            nd['tree'].score = (sc, nd['tree'].score[1])
        return [nd['tree'].score[0] for nd in nodes]
