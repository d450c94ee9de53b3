"""The BASELINE.json configs as parity cases (sizes from SURVEY.md 8: C1..C5)."""
import numpy as np
import pytest

import oracle
from util import desc_to_newick, random_desc

pytestmark = pytest.mark.gpu


def test_c1_greedy_climb_al_fasta(al_profiles, al_fitch, al_greedy):
    """C1 / C3 logic at C1 size: the batched greedy climb reproduces the reference's golden
    parsimonyGreedy path 1167 -> 1149 -> 1138 -> 1135."""
    from pylogeny_b200 import fitch, neighbourhood as nbh
    a = fitch.FitchAlignment(al_profiles["profiles"], al_profiles["weights"])
    named = {c["name"]: c for c in al_fitch["named"]}
    d, _, rows, _ = nbh.newick_to_desc(named["caterpillar_file_order"]["newick"], al_profiles["taxa"])
    path, final = nbh.parsimony_greedy(a, d, rows)
    assert path == al_greedy["path_scores"] == [1167, 1149, 1138, 1135]
    assert a.score_trees(final[None], rows)[0] == 1135
    a.close()


def test_c2_full_size_lnl_spr_neighbourhood():
    """C2: 64 taxa x 10k sites, GTR+G4 lnL of all 14,762 SPR neighbours.  Size-independent
    properties at full size: (i) a sample of neighbours equals the full re-traversal of the
    materialised tree and the FP64 oracle (1e-8); (ii) lnL is linear in the pattern weights --
    doubling them doubles every lnL exactly, splitting them splits it; (iii) relabelling the tips
    (another descriptor of the same tree on the same taxa) gives the same lnL per topology hash."""
    from pylogeny_b200 import likelihood, neighbourhood as nbh
    from util import simulate_dna
    rng = np.random.default_rng(1001)
    n, S = 64, 10_000
    exch, freqs, alpha = np.array([1.2, 3.4, 0.8, 1.1, 4.1, 1.0]), np.array([0.28, 0.22, 0.24, 0.26]), 0.7
    hidden = random_desc(rng, n)
    seqs = simulate_dna(rng, hidden, rng.exponential(0.05, hidden.shape) + 0.005, S, exch, freqs)
    codes = np.array([[{"A": 1, "C": 2, "G": 4, "T": 8}[c] for c in s] for s in seqs], dtype=np.uint8)
    holes = rng.random(codes.shape)
    codes[holes < 0.01] = 15                     # gaps
    codes[(holes >= 0.01) & (holes < 0.015)] = 5  # R = A or G
    w = rng.integers(1, 4, S).astype(np.float64)
    d = random_desc(rng, n)
    br = rng.exponential(0.1, d.shape)
    a = likelihood.LikAlignment(codes, w)
    m = likelihood.Model(exch, freqs, alpha)
    moves, lnl = nbh.score_likelihood(a, m, d, br)
    M = 2 * (n - 3) * (2 * n - 7)
    assert len(lnl) == M == 14762 and np.all(np.isfinite(lnl)) and np.all(lnl < 0)
    # 96 neighbours spread over the whole move list AND over every regraft distance the neighbourhood
    # holds (moves[:, 3] = steps from the prune point: the depth of the visit inside its search walk):
    # each against the full re-traversal of the materialised tree and against the FP64 oracle
    dist = moves[:, 3]
    per_depth = [np.flatnonzero(dist == k) for k in np.unique(dist)]
    sel = np.unique(np.concatenate([np.linspace(0, M - 1, 40).astype(np.int64)] +
                                   [ids[np.linspace(0, len(ids) - 1, min(len(ids), 3)).astype(np.int64)] for ids in per_depth]))
    if len(sel) < 96:
        sel = np.unique(np.concatenate([sel, rng.choice(M, 96 - len(sel), replace=False)]))
    assert len(sel) >= 64 and set(np.unique(dist[sel]).tolist()) == set(np.unique(dist).tolist())
    descs, brs, hashes = nbh.neighbour_trees(d, br, sel)
    np.testing.assert_allclose(lnl[sel], a.score_trees(m, descs, brs), rtol=1e-11)
    for i in range(len(sel)):
        want = oracle.lnl(codes, w, exch, freqs, alpha, descs[i], brs[i])
        assert abs(lnl[sel[i]] - want) <= 1e-8 * abs(want)   # north_star tolerance
    a2 = likelihood.LikAlignment(codes, 2.0 * w)
    assert nbh.score_likelihood(a2, m, d, br, want_moves=False)[1].tolist() == (2.0 * lnl).tolist()
    wa = np.where(rng.random(S) < 0.5, w, 0.0)
    a3, a4 = likelihood.LikAlignment(codes, wa), likelihood.LikAlignment(codes, w - wa)
    parts = nbh.score_likelihood(a3, m, d, br, want_moves=False)[1] + nbh.score_likelihood(a4, m, d, br, want_moves=False)[1]
    np.testing.assert_allclose(parts, lnl, rtol=1e-12)
    # the same tree with its tips renumbered: descriptor ids permuted, tip_rows maps them back
    sigma = rng.permutation(n)
    d2 = np.where(d < n, sigma[np.minimum(d, n - 1)], d).astype(np.int32)
    rows2 = np.empty(n, dtype=np.int32)
    rows2[sigma] = np.arange(n, dtype=np.int32)
    _, lnl2 = nbh.score_likelihood(a, m, d2, br, rows2)
    _, _, h1 = nbh.neighbour_trees(d, br)
    _, _, h2 = nbh.neighbour_trees(d2, br, tip_rows=rows2)
    assert len(set(h1.tolist())) == M and set(h1.tolist()) == set(h2.tolist())
    by_hash = dict(zip(h1.tolist(), lnl.tolist()))
    np.testing.assert_allclose(lnl2, [by_hash[h] for h in h2.tolist()], rtol=1e-11)
    for x in (a, a2, a3, a4):
        x.close()
    m.close()


def test_c3_greedy_climb_full_size():
    """C3: 256 taxa x 100k sites, parsimonyGreedy over the SPR landscape.  Size-independent
    properties: the path is strictly decreasing, every accepted score equals a from-scratch
    re-traversal of the accepted tree, the end point is a local optimum of its neighbourhood."""
    from pylogeny_b200 import fitch, neighbourhood as nbh
    rng = np.random.default_rng(1003)
    n, S = 256, 100_000
    # sites evolved on a hidden tree so that the climb has somewhere to go
    hidden = random_desc(rng, n)
    states = np.zeros((2 * n - 1, S), dtype=np.uint8)
    states[2 * n - 2] = rng.integers(0, 4, S)
    for i in range(n - 2, -1, -1):
        for c in hidden[i]:
            mut = rng.random(S) < 0.04
            states[c] = np.where(mut, rng.integers(0, 4, S), states[n + i])
    dense = np.ascontiguousarray(states[:n].T + 48).astype(np.uint8)
    a = fitch.FitchAlignment.from_dense(dense, np.ones(S, dtype=np.int64))
    start = random_desc(rng, n)
    path, final = nbh.parsimony_greedy(a, start, max_steps=6)
    assert len(path) >= 3 and all(x > y for x, y in zip(path, path[1:]))
    assert int(a.score_trees(final[None])[0]) == path[-1]
    _, costs = nbh.score_parsimony(a, final)
    assert len(costs) == 2 * (n - 3) * (2 * n - 7) == 255530
    # spot-check the engine's neighbour scores at full width against the oracle on a pattern subset
    sub = 2000
    a2 = fitch.FitchAlignment.from_dense(dense[:sub], np.ones(sub, dtype=np.int64))
    _, c2 = nbh.score_parsimony(a2, final)
    labels = ["T%d" % i for i in range(n)]
    x = oracle.FitchInputs([bytes(r).decode() for r in dense[:sub]], [1] * sub, {l: i for i, l in enumerate(labels)})
    sel = np.linspace(0, len(c2) - 1, 12).astype(np.int64)
    descs, _, _ = nbh.neighbour_trees(final, None, sel)
    assert [oracle.fitch_cost(desc_to_newick(d, labels), inputs=x) for d in descs] == c2[sel].tolist()
    a.close(); a2.close()


def test_c4_shape_two_trees_full_width():
    """C4 shape: 128 taxa x 1M sites, GTR+G4 (two random trees instead of 100k).  Properties:
    additivity over a split of the sites, pulley principle; oracle on a 20k-site slice."""
    from pylogeny_b200 import likelihood
    rng = np.random.default_rng(1004)
    n, S = 128, 1_000_000
    codes = (1 << rng.integers(0, 4, (n, S))).astype(np.uint8)
    exch, f = np.array([1.0, 2.0, 0.5, 1.5, 3.0, 1.0]), np.array([0.25, 0.3, 0.2, 0.25])
    m = likelihood.Model(exch, f, 0.8)
    descs = np.stack([random_desc(rng, n) for _ in range(2)])
    brs = rng.exponential(0.1, descs.shape)
    brs2 = brs.copy()
    tot = brs[:, -1].sum(1)
    brs2[:, -1, 0], brs2[:, -1, 1] = 0.9 * tot, 0.1 * tot
    a = likelihood.LikAlignment(codes, np.ones(S))
    full = a.score_trees(m, descs, brs)
    moved = a.score_trees(m, descs, brs2)
    np.testing.assert_allclose(full, moved, rtol=1e-12)
    a.close()
    cut = 20_000
    lo = likelihood.LikAlignment(codes[:, :cut], np.ones(cut))
    hi = likelihood.LikAlignment(codes[:, cut:], np.ones(S - cut))
    part_lo, part_hi = lo.score_trees(m, descs, brs), hi.score_trees(m, descs, brs)
    np.testing.assert_allclose(part_lo + part_hi, full, rtol=1e-12)
    want = oracle.lnl(codes[:, :cut], np.ones(cut), exch, f, 0.8, descs[0], brs[0])
    assert abs(part_lo[0] - want) <= 1e-8 * abs(want)
    lo.close(); hi.close(); m.close()


def test_c5_protein_tbr_neighbourhood():
    """C5: synthetic protein 50 taxa x 5k sites, the SHIPPED WAG table (likelihood.protein_model("WAG"),
    what a "WAG, p1 = ..." partition file selects) + G4, every TBR neighbour.  Checks: a sample against
    the oracle (given the same table), a sample against re-traversal, SPR neighbours are found again
    inside the TBR neighbourhood; then a sample again under JTT."""
    from pylogeny_b200 import _lib, likelihood, neighbourhood as nbh
    rng = np.random.default_rng(1005)
    n, S = 50, 5000
    codes = rng.integers(0, 20, (n, S)).astype(np.uint8)
    exch, f = likelihood.protein_model("WAG")
    a = likelihood.LikAlignment(codes, np.ones(S), _lib.DATA_AA)
    m = likelihood.Model(exch, f, 0.9)
    d = random_desc(rng, n)
    br = rng.exponential(0.1, d.shape)
    _, lnl = nbh.score_likelihood(a, m, d, br, move_type=nbh.MOVE_TBR)
    M = len(lnl)
    assert M == nbh.neighbourhood_size(d, nbh.MOVE_TBR) and M > 2 * (n - 3) * (2 * n - 7)
    sel = np.linspace(0, M - 1, 12).astype(np.int64)
    descs, brs, hashes = nbh.neighbour_trees(d, br, sel, nbh.MOVE_TBR)
    np.testing.assert_allclose(a.score_trees(m, descs, brs), lnl[sel], rtol=1e-11)
    for j in range(0, 12, 2):
        want = oracle.lnl(codes, np.ones(S), exch, f, 0.9, descs[j], brs[j])
        assert abs(lnl[sel[j]] - want) <= 1e-8 * abs(want)
    ej, fj = likelihood.protein_model("JTT")
    mj = likelihood.Model(ej, fj, 0.9)
    _, lj = nbh.score_likelihood(a, mj, d, br, move_type=nbh.MOVE_TBR, want_moves=False)
    assert len(lj) == M and not np.allclose(lj[sel], lnl[sel], rtol=1e-6)
    for j in (1, 6, 11):
        want = oracle.lnl(codes, np.ones(S), ej, fj, 0.9, descs[j], brs[j])
        assert abs(lj[sel[j]] - want) <= 1e-8 * abs(want)
    mj.close()
    # SPR neighbours are TBR neighbours (same topology; branch lengths may differ when the first
    # TBR move that reaches a topology is not the SPR one, so only identity is compared)
    _, _, h_spr = nbh.neighbour_trees(d, br, np.arange(0, 2 * (n - 3) * (2 * n - 7), 97), nbh.MOVE_SPR)
    _, _, h_all = nbh.neighbour_trees(d, br, None, nbh.MOVE_TBR)
    assert set(h_spr.tolist()) <= set(h_all.tolist())
    a.close(); m.close()


def test_likelihood_climb_scan_then_batched_smoothing():
    """Likelihood hill-climb, shortlist mode (an approximation: a fixed-length scan of the whole SPR
    neighbourhood ranks it, the best few are smoothed in one batch).  The path increases strictly, the
    final lnL is what one libpllWrapper handle returns for the final topology, and the climb recovers
    (most of) the gap to the tree the data were simulated on."""
    from pylogeny_b200 import alignment, libpllWrapper as W, neighbourhood as nbh, pll
    from util import simulate_dna
    rng = np.random.default_rng(77)
    n, S = 16, 1500
    true = random_desc(rng, n)
    br = rng.exponential(0.07, true.shape) + 0.01
    seqs = simulate_dna(rng, true, br, S, np.array([1.0, 3.0, 0.7, 1.3, 4.0, 1.0]), np.array([0.3, 0.2, 0.2, 0.3]))
    ali = alignment.phylipFriendlyAlignment(names=["s%d" % i for i in range(n)], seqs=seqs)
    labels = ali.getTaxa()
    start = desc_to_newick(random_desc(rng, n), labels)
    path, final = nbh.likelihood_greedy(ali, start, shortlist=12, max_steps=6)
    assert len(path) >= 3 and all(b > a for a, b in zip(path, path[1:]))
    pm = ali._pllmodel
    h = W.new(ali.getPhylip(), final, pm.getFileName())
    again = W.getLogLikelihood(h)                       # same topology, smoothed again from z = 0.9
    W.destroy(h)
    assert abs(again - path[-1]) <= 1e-9 * abs(again)
    (truth,), _ = W.getLogLikelihoodBatch(ali.getPhylip(), [desc_to_newick(true, labels)], pm.getFileName())
    assert path[-1] > path[0] + 0.5 * (truth - path[0])  # at least half of the way to the generating tree
    pm.close(); ali.close()


def test_likelihood_climb_reference_semantics_smooths_every_neighbour():
    """likelihood_greedy() by default gives EVERY neighbour the reference's per-tree treatment
    (heuristic.py:148-153 -> vertex.scoreLikelihood -> scoring.getLogLikelihood: default lengths, 3
    smoothing passes, lnL).  One climb step on a small case equals doing exactly that through one
    libpllWrapper handle per neighbour (the unbatched drop-in path), and picks the same winner."""
    from pylogeny_b200 import alignment, libpllWrapper as W, neighbourhood as nbh, pll
    from util import simulate_dna
    rng = np.random.default_rng(78)
    n, S = 7, 600
    true = random_desc(rng, n)
    br = rng.exponential(0.08, true.shape) + 0.01
    seqs = simulate_dna(rng, true, br, S, np.array([1.0, 3.0, 0.7, 1.3, 4.0, 1.0]), np.array([0.3, 0.2, 0.2, 0.3]))
    ali = alignment.phylipFriendlyAlignment(names=["s%d" % i for i in range(n)], seqs=seqs)
    labels = ali.getTaxa()
    taxa = {l: i for i, l in enumerate(labels)}
    start = desc_to_newick(random_desc(rng, n), labels)
    path, final = nbh.likelihood_greedy(ali, start, max_steps=1)
    part = pll.model_for(ali).getFileName()
    # the same step by hand, one handle per tree
    h = W.new(ali.getPhylip(), start, part)
    cur = W.getLogLikelihood(h)
    cur_nw = W.getNewickString(h)
    W.destroy(h)
    assert abs(path[0] - cur) <= 1e-9 * abs(cur)
    d, b, rows, tips = nbh.newick_to_desc(cur_nw, taxa)
    descs, _, _ = nbh.neighbour_trees(d, None, None, tip_rows=rows)
    assert len(descs) == 2 * (n - 3) * (2 * n - 7)
    per_tree = []
    for x in descs:
        h = W.new(ali.getPhylip(), nbh.desc_to_newick(x, tips), part)
        per_tree.append(W.getLogLikelihood(h))
        W.destroy(h)
    best = max(per_tree)
    if best > cur:
        assert len(path) == 2 and abs(path[1] - best) <= 1e-9 * abs(best)
        d2, _, r2, _ = nbh.newick_to_desc(final, taxa)
        assert nbh.topology_hash(d2, r2) == nbh.topology_hash(descs[int(np.argmax(per_tree))], rows)
    else:
        assert len(path) == 1
    pll.model_for(ali).close(); ali.close()
