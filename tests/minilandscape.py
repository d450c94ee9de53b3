"""TEST INFRASTRUCTURE: the smallest landscape / tree pair that satisfies what
pylogeny_b200.explore.BatchedExploreMixin uses of the reference's classes, so the mixin can run
on the GPU box, where /root/reference does not exist.  Method names and meaning follow
landscape.py:58-138 (graph, getNode, getEdge, _newNode, findTreeTopologyByStructure, defaultWeight,
parsimony_profiles, locks) and tree.py (getNewick, getStructure, toTopology().toNewick(), score,
origin).  tests/test_explore_cpu.py runs the same mixin over the REAL reference classes."""
import networkx

from pylogeny_b200 import parsimony


def _parse(s):
    """Newick -> nested tuples of labels (branch lengths dropped)."""
    s = s.strip().rstrip(';')
    pos = [0]

    def sub():
        if s[pos[0]] == '(':
            pos[0] += 1
            kids = [sub()]
            while s[pos[0]] == ',':
                pos[0] += 1
                kids.append(sub())
            pos[0] += 1  # ')'
            node = tuple(kids)
        else:
            j = pos[0]
            while j < len(s) and s[j] not in ',:)':
                j += 1
            node = s[pos[0]:j]
            pos[0] = j
        while pos[0] < len(s) and s[pos[0]] not in ',)':
            pos[0] += 1  # label / length after a subtree
        return node

    return sub()


def _structure(node):
    """Child-sorted Newick without lengths (tree.py getStructure's role: string identity)."""
    if isinstance(node, str):
        return node
    return "(" + ",".join(sorted(_structure(c) for c in node)) + ")"


class MiniTree(object):
    def __init__(self, newick, check=False):
        self.newick = newick
        self.score = (None, None)
        self.origin = None

    def getNewick(self):
        return self.newick

    def getStructure(self):
        return _structure(_parse(self.newick)) + ";"

    def toTopology(self):
        return self

    def toNewick(self):
        return self.newick

    def toTree(self):
        return self


class MiniLandscape(object):
    defaultWeight = 0.0

    def __init__(self, ali, starting_tree):
        self.alignment = ali
        self.graph = networkx.Graph()
        self.parsimony_profiles = parsimony.profile_set(ali)
        self.locks = []
        self.nextTree = 0
        self.root = self._newNode(starting_tree)

    def _newNode(self, t, score=False):
        i = self.nextTree
        self.graph.add_node(i, tree=t, explored=False)
        self.nextTree += 1
        return i

    def getNode(self, i):
        return self.graph.nodes[i]

    def getTree(self, i):
        return self.graph.nodes[i]['tree']

    def getEdge(self, i, j):
        return self.graph.get_edge_data(i, j)

    def findTreeTopologyByStructure(self, struct):
        for i in self.graph.nodes:
            if self.getTree(i).getStructure() == struct:
                return i
        return None

    def __len__(self):
        return self.graph.number_of_nodes()
