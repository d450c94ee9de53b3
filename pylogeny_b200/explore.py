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
Tie rule: by default neighbours are inserted in the engine's enumeration order (pruned branch by node
id, then depth-first by target), so among equal scores `sorted(...)[0]` (heuristic.py:92-95) may pick
a different tree than the reference's enumeration order would.  Set `reference_order = True` on the
landscape to insert them in the order the reference's own `topol.allType(type)` walks its moves
(first occurrence of every topology) and to build each NEW tree with the reference's own
`rearrangement.toTree()` (its Newick child order decides how the reference lists that tree's
neighbours later), which makes ties resolve as in the reference: same tree at every step of
`parsimonyGreedy`, same node order over the whole landscape (tests/test_explore_cpu.py).

Locks (`landscape.locks`, landscape.py:714-718): the reference's move list IS the rule -- a branch may
only move inside its side of every locked bipartition (rearrangement.py:228-240 `_getPartition`), the
flipped pass re-roots without the locks (rearrangement.py:31-40), and a tree that already violates a
lock only steps to non-violating neighbours (landscape.py:743-746).  With locks set (or
`reference_order`) the mixin therefore asks the reference's own topology object for its lazy move list
(no tree is built, nothing is scored), maps every (moved branch, destination) pair onto the engine's
move table by the leaf sets of the two branches (`neighbourhood.all_moves`), and keeps exactly the
topologies some listed move reaches; scoring stays one engine call for the whole neighbourhood.

Likelihood: `scoreLikelihoodFor(indices)` scores many nodes in one engine call (batched branch-length
smoothing); with `batch_likelihood = True` every explored neighbourhood is scored that way.

Falls back to the reference method only for move types the engine does not enumerate for it (TBR: the
reference has no TBR producer either, rearrangement.py:661-664).
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

    def _b200_reference_moves(self, topol, type, desc, tips):
        """Indices (into the engine's list of distinct neighbours) of the topologies the reference's own
        move list reaches, in the order of their first occurrence in it.  `topol` is the reference's
        topology object of the tree with the landscape's locks applied; its `allType(type)` list is
        lazy: (structure, moved branch, destination branch) triples, rearrangement.py:44-62.  Also
        returns, per index, the first such move (materialised by the caller only for NEW topologies)."""
        n = len(tips)
        bit = {label: 1 << k for k, label in enumerate(tips)}
        full = (1 << n) - 1
        below = [1 << k for k in range(n)] + [0] * (n - 1)
        for k, (a, b) in enumerate(desc):
            below[n + k] = below[a] | below[b]
        # unrooted edge named by its child end (the edge through the root by the root's first child,
        # as the engine names it): leaf set of the moved part -> (edge, 0 = the subtree below it, 1 = the rest)
        edge_of = {}
        for c in list(range(2 * n - 2)):
            if c == desc[-1][1]:
                continue
            edge_of[below[c]] = (c, 0)
            edge_of[full ^ below[c]] = (c, 1)
        table = nbh.all_moves(desc, type)
        uid = {(int(c), int(s_), int(t)): int(u) for c, s_, t, _, u in table}
        masks = {}

        def masks_of(struct):
            m = masks.get(id(struct))
            if m is None:
                m = masks[id(struct)] = {}
                stack = [(br, False) for br in struct.getRoot().children]
                while stack:
                    br, done = stack.pop()
                    kids = br.child.children
                    if not kids:
                        m[br] = bit[br.child.label]
                    elif done:
                        v = 0
                        for k in kids:
                            v |= m[k]
                        m[br] = v
                    else:
                        stack.append((br, True))
                        stack.extend((k, False) for k in kids)
            return m

        order, seen, first = [], set(), {}
        for en in topol.allType(type):
            m = masks_of(en.topol)
            c, side = edge_of[m[en.target]]
            t = edge_of[m[en.destination]][0]
            u = uid.get((c, side, t))
            if u is None:  # (cannot happen: every legal reference move is a (prune, target) pair of the table)
                raise RuntimeError("reference move %s has no counterpart in the engine's move table" % (en,))
            if u not in seen:
                seen.add(u)
                order.append(u)
                first[u] = en
        return order, first

    def exploreTree(self, i, type=nbh.MOVE_SPR):
        if type not in (nbh.MOVE_SPR, nbh.MOVE_NNI):
            return super(BatchedExploreMixin, self).exploreTree(i, type)
        p = self.parsimony_profiles
        node = self.getNode(i)
        if node['explored']:
            return list()
        tre = node['tree']
        treecls = tre.__class__
        topol = tre.toTopology()
        newick = topol.toNewick()            # rooted at the lowest-order leaf
        desc, brlen, rows, tips = nbh.newick_to_desc(newick, p.taxa)
        aln = scoring._resident(p)
        moves, costs = nbh.score_parsimony(aln, desc, rows, type)
        descs, brs, hashes = nbh.neighbour_trees(desc, brlen, None, type, rows)
        index = self._b200_sync_index(p.taxa)
        origin = 'SPR' if type == nbh.MOVE_SPR else 'NNI'
        locks = getattr(self, "locks", None)
        isvi = False
        order, first = range(len(costs)), None
        if locks or getattr(self, "reference_order", False):
            if locks:
                isvi = self._isViolating(topol)
                for lock in locks:           # landscape.py:714-718
                    hasthis = topol.getBranchFromBipartition(lock)
                    if hasthis:
                        topol.lockBranch(hasthis)
            order, first = self._b200_reference_moves(topol, type, desc, tips)
        neighbors = list()
        for j in order:
            h = int(hashes[j])
            found = index.get(h)
            if found is not None:
                if found != i and self.graph.has_node(found) and not self.graph.has_edge(found, i):
                    self.graph.add_edge(found, i)
                    self.getEdge(found, i)['weight'] = self.defaultWeight
                continue
            if first is not None:
                # the reference's own tree object of its first move to this topology: the child order of
                # its Newick decides the order in which the reference would list ITS neighbours later
                t = first[j].toTree()
            else:
                t = treecls(nbh.desc_to_newick(descs[j], tips, brs[j]), check=True)
            if self.findTreeTopologyByStructure(t.getStructure()) is not None:
                continue  # (string already present under another hash: cannot happen for one topology)
            if isvi and self._isViolating(t.toTopology()):
                continue  # landscape.py:743-746
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
