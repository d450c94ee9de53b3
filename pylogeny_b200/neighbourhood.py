"""Whole-neighbourhood scoring: the batched consumer the landscape should call instead of one
scoring call per neighbour (landscape.py:724-756).

Trees are rooted binary post-order descriptors (int32 [n_tips-1, 2], see
include/pylogeny_b200.h); `newick_to_desc` / `desc_to_newick` convert from and to the Newick
strings the rest of Pylogeny uses.  Enumeration and materialisation run on the host (C++ in the
C-ABI library); scoring runs on the GPU by edge insertion.
"""
import ctypes as C

import numpy as np

from . import _lib
from ._lib import MOVE_NNI, MOVE_SPR, MOVE_SPR_WITHIN, MOVE_TBR  # noqa: F401


def _desc(desc):
    d = np.ascontiguousarray(desc, dtype=np.int32)
    if d.ndim != 2 or d.shape[1] != 2:
        raise ValueError("descriptor must be int32 [n_tips-1, 2]")
    return d


def neighbourhood_size(desc, move_type=MOVE_SPR):
    d = _desc(desc)
    n = C.c_int64()
    _lib.check(_lib.lib().plg_neighbourhood_size(move_type, d.shape[0] + 1, d.ctypes.data, C.byref(n)))
    return n.value


def _rows(tip_rows):
    return np.ascontiguousarray(tip_rows, dtype=np.int32) if tip_rows is not None else None


def topology_hash(desc, tip_rows=None):
    """Canonical (rooting- and leaf-order-independent) hash; tips identified by tip_rows[id]."""
    d = _desc(desc)
    rows = _rows(tip_rows)
    h = C.c_uint64()
    _lib.check(_lib.lib().plg_topology_hash(d.shape[0] + 1, d.ctypes.data,
                                            rows.ctypes.data if rows is not None else None, C.byref(h)))
    return h.value


def all_moves(desc, move_type=MOVE_SPR):
    """Every (pruned branch, target branch) pair of the NNI / SPR neighbourhood, repeats included: int32
    [A, 5] = (child end of pruned branch, 0 subtree / 1 complement moves, child end of target branch,
    distance, index of the move's topology among the distinct neighbours)."""
    d = _desc(desc)
    n = C.c_int64()
    _lib.check(_lib.lib().plg_neighbourhood_all_moves(move_type, d.shape[0] + 1, d.ctypes.data, C.byref(n), None))
    out = np.empty((n.value, 5), dtype=np.int32)
    _lib.check(_lib.lib().plg_neighbourhood_all_moves(move_type, d.shape[0] + 1, d.ctypes.data, C.byref(n), out.ctypes.data))
    return out


def neighbour_trees(desc, brlen=None, sel=None, move_type=MOVE_SPR, tip_rows=None):
    """Materialise neighbours `sel` (default: all). Returns (descs [m, n-1, 2], brlens or None, hashes [m])."""
    d = _desc(desc)
    rows = _rows(tip_rows)
    if sel is None:
        sel = np.arange(neighbourhood_size(d, move_type), dtype=np.int64)
    sel = np.ascontiguousarray(sel, dtype=np.int64)
    b = np.ascontiguousarray(brlen, dtype=np.float64) if brlen is not None else None
    out_d = np.zeros((len(sel),) + d.shape, dtype=np.int32)
    out_b = np.zeros((len(sel),) + d.shape, dtype=np.float64) if b is not None else None
    out_h = np.zeros(len(sel), dtype=np.uint64)
    _lib.check(_lib.lib().plg_neighbourhood_trees(
        move_type, d.shape[0] + 1, d.ctypes.data, b.ctypes.data if b is not None else None,
        rows.ctypes.data if rows is not None else None, len(sel), sel.ctypes.data, out_d.ctypes.data, out_b.ctypes.data if out_b is not None else None, out_h.ctypes.data))
    return out_d, out_b, out_h


def score_parsimony(aln, desc, tip_rows=None, move_type=MOVE_SPR, shard=(0, 1), want_moves=True):
    """Fitch length of every distinct neighbour. aln: fitch.FitchAlignment.
    Returns (moves int32 [M, 4], costs int32 [M]); with shard=(i, c) only block i of c is filled.
    want_moves=False skips the move table (moves is None): a climb only needs the argmin, which
    `neighbour_trees(desc, sel=[i])` materialises."""
    d = _desc(desc)
    rows = np.ascontiguousarray(tip_rows, dtype=np.int32) if tip_rows is not None else None
    n = C.c_int64()
    L = _lib.lib()
    M = neighbourhood_size(d, move_type)
    moves = np.empty((M, 4), dtype=np.int32) if want_moves else None  # every row is written by the call
    costs = np.zeros(M, dtype=np.int32)
    _lib.check(L.plg_fitch_score_neighbourhood(aln._h, move_type, d.shape[0] + 1, d.ctypes.data,
                                               rows.ctypes.data if rows is not None else None, shard[0], shard[1],
                                               C.byref(n), moves.ctypes.data if want_moves else None,
                                               costs.ctypes.data, None))
    return (moves[:n.value] if want_moves else None), costs[:n.value]


def score_likelihood(aln, model, desc, brlen, tip_rows=None, move_type=MOVE_SPR, shard=(0, 1), want_moves=True):
    """lnL (fixed branch lengths) of every distinct neighbour. aln: likelihood.LikAlignment.
    Returns (moves int32 [M, 4] or None, lnl float64 [M])."""
    d = _desc(desc)
    b = np.ascontiguousarray(brlen, dtype=np.float64)
    rows = np.ascontiguousarray(tip_rows, dtype=np.int32) if tip_rows is not None else None
    n = C.c_int64()
    M = neighbourhood_size(d, move_type)
    moves = np.empty((M, 4), dtype=np.int32) if want_moves else None  # every row is written by the call
    lnl = np.zeros(M, dtype=np.float64)
    _lib.check(_lib.lib().plg_lik_score_neighbourhood(
        aln._h, model._h, move_type, d.shape[0] + 1, d.ctypes.data, b.ctypes.data,
        rows.ctypes.data if rows is not None else None, shard[0], shard[1], C.byref(n),
        moves.ctypes.data if want_moves else None, lnl.ctypes.data, None))
    return (moves[:n.value] if want_moves else None), lnl[:n.value]


def newick_to_desc(newick, taxa):
    """Newick string -> (desc int32 [n-1, 2], brlen float64 [n-1, 2], tip_rows int32 [n]).

    `taxa` maps leaf label -> alignment row (the profile_set.taxa dict).  Tips are numbered in order
    of appearance.  A multifurcating node (e.g. the trifurcating root of toUnrootedNewick,
    rearrangement.py:715-732) is resolved left to right with zero-length branches, which leaves the
    unrooted topology and every likelihood unchanged."""
    s = newick.strip().rstrip(';')
    pos = [0]
    tips, rows, desc, brl = [], [], [], []

    def length():
        if pos[0] < len(s) and s[pos[0]] == ':':
            j = pos[0] + 1
            while j < len(s) and s[j] not in ',)':
                j += 1
            try:
                v = float(s[pos[0] + 1:j])
            except ValueError:
                v = 0.0
            pos[0] = j
            return v
        return 0.0

    def label():
        j = pos[0]
        while j < len(s) and s[j] not in ',:;)':
            j += 1
        lab = s[pos[0]:j]
        pos[0] = j
        return lab

    def sub():
        if s[pos[0]] == '(':
            pos[0] += 1
            kids = [sub()]
            while s[pos[0]] == ',':
                pos[0] += 1
                kids.append(sub())
            if s[pos[0]] != ')':
                raise ValueError("malformed Newick string")
            pos[0] += 1
            label()
            ln = length()
            node, nl = kids[0]
            for k, kl in kids[1:]:
                desc.append((node, k))
                brl.append((nl, kl))
                node, nl = ('i', len(desc) - 1), 0.0
            return node, ln
        lab = label()
        ln = length()
        tips.append(lab)
        rows.append(taxa[lab])
        return ('t', len(tips) - 1), ln

    sub()
    n = len(tips)
    ids = lambda x: x[1] if x[0] == 't' else n + x[1]
    d = np.array([(ids(a), ids(b)) for a, b in desc], dtype=np.int32).reshape(-1, 2)
    return d, np.array(brl, dtype=np.float64).reshape(-1, 2), np.array(rows, dtype=np.int32), tips


def desc_to_newick(desc, labels, brlen=None):
    """Rooted descriptor -> Newick string; children sorted by label like newick.py:40-46 does."""
    n = len(labels)
    text = {}
    for i, (a, b) in enumerate(np.asarray(desc)):
        parts = []
        for k, c in enumerate((int(a), int(b))):
            is_tip = c < n
            t = labels[c] if is_tip else text[c]
            key = labels[c] if is_tip else ''
            if brlen is not None and brlen[i][k] > 0:
                t += ":%s" % str(float(brlen[i][k]))
            parts.append((key, t))
        parts.sort(key=lambda p: p[0])
        text[n + i] = "(%s)" % ",".join(p[1] for p in parts)
    return text[n + len(desc) - 1] + ";"


def parsimony_greedy(aln, desc, tip_rows=None, move_type=MOVE_SPR, max_steps=10000):
    """Greedy hill-climb by parsimony over the rearrangement landscape, one engine call per step
    (what heuristic.parsimonyGreedy.explore does through landscape.exploreTree, heuristic.py:65-105:
    move to the best neighbour while it is strictly better).  Ties: first minimum in the engine's
    enumeration order.  Returns (path of costs, final descriptor)."""
    d = _desc(desc)
    cur = int(aln.score_trees(d[None], tip_rows)[0])
    path = [cur]
    for _ in range(max_steps):
        _, costs = score_parsimony(aln, d, tip_rows, move_type, want_moves=False)
        if len(costs) == 0:
            break
        best = int(np.argmin(costs))
        if int(costs[best]) >= cur:
            break
        descs, _, _ = neighbour_trees(d, None, [best], move_type)
        d, cur = descs[0], int(costs[best])
        path.append(cur)
    return path, d


def likelihood_greedy(ali, newick, shortlist=None, max_steps=1000, passes=3):
    """Greedy hill-climb by likelihood over the SPR landscape with the reference's per-tree scoring
    semantics: EVERY neighbour of the current tree is scored the way `vertex.scoreLikelihood` ->
    `scoring.getLogLikelihood` scores it (landscape.py:1302-1314, scoring.py:41-75, as
    heuristic.likelihoodGreedy does at heuristic.py:148-153) -- default branch lengths, `passes`
    smoothing passes, lnL -- in ONE batched engine call per step (`libpllWrapper.getLogLikelihoodBatch`),
    and the climb moves to the best neighbour while it beats the current tree's smoothed lnL.

    shortlist=k (an APPROXIMATION, not the reference's semantics -- "lazy SPR"): the neighbourhood is
    first ranked by GTR+G4 lnL at the current tree's branch lengths (fixed-length edge-insertion scan,
    `score_likelihood`, ~1e3 x cheaper per neighbour) and only the k best are smoothed.  The path can
    differ from shortlist=None when the scan's ranking and the smoothed ranking disagree beyond rank k.

    `ali`: alignment.phylipFriendlyAlignment (DNA); `newick`: start tree on its taxa.
    Returns (path of smoothed lnL, final Newick with optimised branch lengths)."""
    import ctypes as C_
    from . import likelihood, libpllWrapper as W, pll
    phy, part = ali.getPhylip(), pll.model_for(ali).getFileName()
    labels = ali.getTaxa()
    taxa = {l: i for i, l in enumerate(labels)}
    la = model = None
    if shortlist is not None:
        # the model the libpllWrapper drop-in builds for this alignment (GTR, all rates 1, empirical
        # frequencies, alpha = 1): read it back from a handle so both calls score under the same model
        h = W.new(phy, newick, part)
        K, P = C_.c_int(), C_.c_int64()
        freqs, rates = np.zeros(4), np.zeros(4)
        dp = C_.POINTER(C_.c_double)
        _lib.check(_lib.lib().plg_pll_info(h.ptr, C_.byref(K), C_.byref(P), freqs.ctypes.data_as(dp), rates.ctypes.data_as(dp)))
        W.destroy(h)
        if K.value != 4:
            raise ValueError("the shortlist scan is written for DNA alignments")
        codes, w = likelihood.compress_columns(likelihood.encode_sequences(ali.toStrList(), _lib.DATA_DNA))
        la = likelihood.LikAlignment(codes, w)
        model = likelihood.Model(np.ones(6), freqs, 1.0)
    try:
        (cur,), (cur_nw,) = W.getLogLikelihoodBatch(phy, [newick], part, passes)
        path = [cur]
        for _ in range(max_steps):
            d, br, rows, tips = newick_to_desc(cur_nw, taxa)
            if neighbourhood_size(d) == 0:
                break
            sel = None
            if shortlist is not None:
                _, scan = score_likelihood(la, model, d, br, rows, want_moves=False)
                sel = np.argsort(-scan, kind="stable")[:shortlist].astype(np.int64)
            descs, _, _ = neighbour_trees(d, None, sel, tip_rows=rows)
            cands = [desc_to_newick(x, tips) for x in descs]
            lnl, opt = W.getLogLikelihoodBatch(phy, cands, part, passes)
            j = int(np.argmax(lnl))
            if not lnl[j] > cur:
                break
            cur, cur_nw = lnl[j], opt[j]
            path.append(cur)
        return path, cur_nw
    finally:
        if la is not None:
            la.close()
            model.close()
