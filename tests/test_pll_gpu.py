"""libpllWrapper / pll.py / scoring.getLogLikelihood compat path on the GPU: evaluation at the
default branch lengths and after 3 smoothing passes, both checked against the FP64 oracle at the
branch lengths the handle reports (1e-8 relative)."""
import ctypes as C

import numpy as np
import pytest

import oracle
from util import desc_to_newick, random_desc, simulate_dna

pytestmark = pytest.mark.gpu


def _oracle_at(newick, taxa, codes, w, exch, freqs, alpha):
    from pylogeny_b200 import neighbourhood as nbh
    d, br, rows, _ = nbh.newick_to_desc(newick, taxa)
    return oracle.lnl(codes[rows], w, exch, freqs, alpha, d, br)


def _case(n=12, S=1500, seed=3):
    from pylogeny_b200 import alignment
    rng = np.random.default_rng(seed)
    d = random_desc(rng, n)
    br = rng.exponential(0.08, d.shape) + 0.01
    exch = np.array([1.0, 3.0, 0.7, 1.3, 4.0, 1.0])
    f = np.array([0.3, 0.2, 0.2, 0.3])
    seqs = simulate_dna(rng, d, br, S, exch, f)
    seqs[1] = seqs[1][:20] + "N-RY" + seqs[1][24:]      # ambiguity codes and gaps
    ali = alignment.phylipFriendlyAlignment(names=["tax%d" % i for i in range(n)], seqs=seqs)
    return ali, d


def test_new_evaluate_smooth_newick():
    from pylogeny_b200 import _lib, libpllWrapper as W, likelihood, pll
    ali, d = _case()
    n = ali.getNumSeqs()
    labels = ali.getTaxa()
    # unrooted Newick with a trifurcating top node, as toUnrootedNewick() hands over (rearrangement.py:715-732)
    rooted = desc_to_newick(d, labels)
    pm = pll.partitionModel(ali)
    pm.createSimpleModel()
    h = W.new(ali.getPhylip(), rooted, pm.getFileName())
    K, P = C.c_int(), C.c_int64()
    freqs, rates = np.zeros(4), np.zeros(4)
    dp = C.POINTER(C.c_double)
    _lib.check(_lib.lib().plg_pll_info(h.ptr, C.byref(K), C.byref(P), freqs.ctypes.data_as(dp), rates.ctypes.data_as(dp)))
    assert K.value == 4 and P.value <= ali.getSize()
    np.testing.assert_allclose(rates, oracle.gamma_rates(1.0, 4), rtol=1e-12)
    codes = likelihood.encode_sequences(ali.toStrList(), _lib.DATA_DNA)
    counts = np.array([(codes == m).sum() for m in (1, 2, 4, 8)], float)
    assert np.allclose(freqs, counts / counts.sum(), atol=2e-3)        # empirical (ambiguities shared out)
    taxa = {l: i for i, l in enumerate(labels)}
    w = np.ones(ali.getSize())
    l0 = W.evaluate(h)
    nw0 = W.getNewickString(h)
    assert abs(l0 - _oracle_at(nw0, taxa, codes, w, np.ones(6), freqs, 1.0)) <= 1e-8 * abs(l0)
    from pylogeny_b200 import neighbourhood as nbh
    _, br0, _, _ = nbh.newick_to_desc(nw0, taxa)
    assert np.allclose(br0[br0 > 0], -np.log(0.9))                      # default z = 0.9 on every branch
    l1 = W.getLogLikelihood(h)
    nw1 = W.getNewickString(h)
    assert l1 > l0 + 1.0
    assert abs(l1 - _oracle_at(nw1, taxa, codes, w, np.ones(6), freqs, 1.0)) <= 1e-8 * abs(l1)
    l2 = W.getLogLikelihood(h)                                          # three more passes: nearly converged
    assert l2 >= l1 - 1e-6 and (l2 - l1) < 1e-3 * abs(l1)
    W.destroy(h)
    with pytest.raises(ReferenceError):
        W.getLogLikelihood(h)
    with pytest.raises(IOError):
        W.new("/nonexistent.phy", rooted, pm.getFileName())
    pm.close()
    ali.close()


def test_scoring_getLogLikelihood_dropin():
    """scoring.getLogLikelihood(tree, alignment) with the reference's call pattern (scoring.py:41-75)."""
    from pylogeny_b200 import scoring
    ali, d = _case(n=8, S=600, seed=9)
    labels = ali.getTaxa()

    class Topo(object):
        def toUnrootedNewick(self):
            return desc_to_newick(d, labels)

    class Tree(object):
        newick = None

        def toTopology(self):
            return Topo()

        def updateNewick(self, nw, reroot=True):
            self.newick = nw

    t = Tree()
    sc = scoring.getLogLikelihood(t, ali)
    assert isinstance(sc, float) and sc < 0
    assert t.newick.endswith(";") and t.newick.count(":") == 2 * 8 - 3
    assert scoring.getLogLikelihood(t, "not an alignment") is None       # any failure -> None (scoring.py:59-60)
    ali.close()


def test_phylip_forms(tmp_path):
    """The alignment reader takes the three PHYLIP forms in use -- one line per taxon (what
    alignment.py writes), interleaved blocks, sequential with sequences wrapped over several lines --
    and scores them identically; a truncated file is an IOError as in the reference wrapper."""
    from pylogeny_b200 import libpllWrapper as W, pll
    ali, d = _case(n=9, S=230, seed=17)
    labels, seqs = ali.getTaxa(), ali.toStrList()
    nw = desc_to_newick(d, labels)
    pm = pll.partitionModel(ali)
    pm.createSimpleModel()
    h = W.new(ali.getPhylip(), nw, pm.getFileName())
    want = W.evaluate(h)
    W.destroy(h)
    head = "%d %d\n" % (len(labels), len(seqs[0]))
    inter = head + "".join("%-12s%s\n" % (l, q[:70]) for l, q in zip(labels, seqs))
    for off in range(70, 230, 70):
        inter += "\n" + "".join(" ".join(q[off + k:off + k + 10] for k in range(0, min(70, 230 - off), 10)) + "\n" for q in seqs)
    seq = head + "".join("%s\n%s" % (l, "".join(q[k:k + 60] + "\n" for k in range(0, 230, 60))) for l, q in zip(labels, seqs))
    seq2 = head + "".join("%-10s %s\n%s" % (l, q[:50], "".join(q[k:k + 60] + "\n" for k in range(50, 230, 60))) for l, q in zip(labels, seqs))
    for name, text in (("interleaved", inter), ("sequential", seq), ("sequential2", seq2)):
        path = tmp_path / (name + ".phy")
        path.write_text(text)
        h = W.new(str(path), nw, pm.getFileName())
        assert W.evaluate(h) == want, name
        W.destroy(h)
    bad = tmp_path / "short.phy"
    bad.write_text(seq[:len(seq) // 2])
    with pytest.raises(IOError):
        W.new(str(bad), nw, pm.getFileName())
    pm.close(); ali.close()


def test_wag_model_is_usable():
    """The shipped WAG table builds a valid reversible model (provenance flagged unverified)."""
    from pylogeny_b200 import _lib, alignment, libpllWrapper as W, pll
    rng = np.random.default_rng(4)
    aa = "ARNDCQEGHILKMFPSTWYV"
    seqs = ["".join(aa[i] for i in rng.integers(0, 20, 120)) for _ in range(6)]
    ali = alignment.phylipFriendlyAlignment(names=list("abcdef"), seqs=seqs)
    assert ali.getDataType() == 'protein'
    pm = pll.partitionModel(ali)
    pm.createSimpleModel()
    h = W.new(ali.getPhylip(), "(T0,T1,(T2,(T3,(T4,T5))));", pm.getFileName())
    l0 = W.evaluate(h)
    l1 = W.getLogLikelihood(h)
    assert np.isfinite(l0) and l1 >= l0
    W.destroy(h)
    pm.close(); ali.close()


def test_moves_in_distance_vs_oracle():
    """get{SPR,NNI,All}MovesInDistance (libpllWrapper.c:241-319): every returned Newick string scores,
    under the FP64 oracle, the lnL that came with it (float-rounded like :54); counts follow the
    radius-limited enumeration; best first."""
    from pylogeny_b200 import _lib, libpllWrapper as W, likelihood, neighbourhood as nbh, pll
    ali, d = _case(n=9, S=400, seed=11)
    n = ali.getNumSeqs()
    labels = ali.getTaxa()
    pm = pll.partitionModel(ali)
    pm.createSimpleModel()
    h = W.new(ali.getPhylip(), desc_to_newick(d, labels), pm.getFileName())
    W.getLogLikelihood(h)                         # neighbours are scored at the optimised lengths
    freqs = np.zeros(4)
    _lib.check(_lib.lib().plg_pll_info(h.ptr, None, None, freqs.ctypes.data_as(C.POINTER(C.c_double)), None))
    codes = likelihood.encode_sequences(ali.toStrList(), _lib.DATA_DNA)
    taxa = {l: i for i, l in enumerate(labels)}
    w = np.ones(ali.getSize())
    base_nw = W.getNewickString(h)
    bd, _, brows, _ = nbh.newick_to_desc(base_nw, taxa)
    assert W.getSPRMovesInDistance(h, 0) == []
    nni = W.getNNIMovesInDistance(h, 1)
    assert len(nni) == 2 * (n - 3) and all(k == "NNI" for k, _, _ in nni)
    for radius in (1, 2, 50):
        spr = W.getSPRMovesInDistance(h, radius)
        assert len(spr) == nbh.neighbourhood_size(bd, nbh.MOVE_SPR_WITHIN(radius))
        assert all(k == "SPR" for k, _, _ in spr)
        vals = [v for _, v, _ in spr]
        assert vals == sorted(vals, reverse=True)
        hashes = set()
        for _, v, nw in spr[:: max(1, len(spr) // 25)]:
            want = _oracle_at(nw, taxa, codes, w, np.ones(6), freqs, 1.0)
            assert abs(v - want) <= 4e-7 * abs(want)             # float32 rounding of the reference's tuple
        for _, _, nw in spr:
            dd, _, rows, _ = nbh.newick_to_desc(nw, taxa)
            hashes.add(nbh.topology_hash(dd, rows))
        assert len(hashes) == len(spr) and nbh.topology_hash(bd, brows) not in hashes
    both = W.getAllMovesInDistance(h, 50)
    assert len(both) == 2 * (n - 3) * (2 * n - 7)
    assert sum(k == "NNI" for k, _, _ in both) == 2 * (n - 3)
    assert {nw for _, _, nw in both if _ == _} >= {nw for _, _, nw in nni}
    assert W.getNewickString(h) == base_nw                      # the handle's tree is left as it was (rollback, :62)
    with pytest.raises(ReferenceError):
        W.getSPRMovesInDistance(h, "2")
    W.destroy(h)
    pm.close(); ali.close()


def test_alignment_cache_is_keyed_by_file_identity(tmp_path):
    """Handles on the same PHYLIP file share one parsed, device-resident alignment (the reference
    re-parses it per tree, libpllWrapper.c:138); rewriting the file must not serve stale data."""
    import os
    import time
    from pylogeny_b200 import libpllWrapper as W, pll
    ali, d = _case(n=9, S=600, seed=11)
    labels = ali.getTaxa()
    nw = desc_to_newick(d, labels)
    pm = pll.partitionModel(ali)
    pm.createSimpleModel()
    path = str(tmp_path / "a.phy")
    text = open(ali.getPhylip()).read()
    open(path, "w").write(text)
    h1 = W.new(path, nw, pm.getFileName())
    h2 = W.new(path, nw, pm.getFileName())          # cache hit
    l1, l2 = W.evaluate(h1), W.evaluate(h2)
    assert l1 == l2
    W.destroy(h1)
    assert W.evaluate(h2) == l2                      # the shared alignment outlives the first handle
    # same path, different content (first sequence altered): a fresh parse, a different likelihood
    lines = text.split("\n")
    name, seq = lines[1].split()
    flipped = seq.translate(str.maketrans("ACGT", "CATG"))
    lines[1] = "%s %s" % (name, flipped)
    time.sleep(0.01)
    open(path, "w").write("\n".join(lines))
    os.utime(path, None)
    h3 = W.new(path, nw, pm.getFileName())
    l3 = W.evaluate(h3)
    assert l3 != l2
    os.environ["PLG_PLL_ALN_CACHE"] = "0"
    try:
        h4 = W.new(path, nw, pm.getFileName())       # cache off: same answer as the cached fresh parse
        assert W.evaluate(h4) == l3
        W.destroy(h4)
    finally:
        del os.environ["PLG_PLL_ALN_CACHE"]
    W.destroy(h2); W.destroy(h3)
    pm.close(); ali.close()


def test_handle_lockstep_equals_host_loop(monkeypatch):
    """plg_pll_loglik on one handle runs the whole smoothing program in one cooperative launch (DNA x
    Gamma-4, pll_smooth1.cuh); PLG_PLL_HANDLE_MODE=batch runs the lockstep program as a batch of one,
    =host keeps the literal loop with one device round trip per branch.  Same
    lnL and branch lengths, also when getLogLikelihood is called again on the same handle (the reference
    mutates the instance: the second call continues from the first call's lengths, libpllWrapper.c:205)."""
    from pylogeny_b200 import libpllWrapper as W, neighbourhood as nbh, pll
    ali, d = _case(n=13, S=900, seed=9)
    labels = ali.getTaxa()
    taxa = {l: i for i, l in enumerate(labels)}
    nw = desc_to_newick(d, labels)
    pm = pll.partitionModel(ali)
    pm.createSimpleModel()
    res = {}
    for mode in ("host", "batch", "device"):
        if mode != "device":
            monkeypatch.setenv("PLG_PLL_HANDLE_MODE", mode)
        else:
            monkeypatch.delenv("PLG_PLL_HANDLE_MODE")
        h = W.new(ali.getPhylip(), nw, pm.getFileName())
        first = W.getLogLikelihood(h)
        nw1 = W.getNewickString(h)
        second = W.getLogLikelihood(h)
        nw2 = W.getNewickString(h)
        ev = W.evaluate(h)
        W.destroy(h)
        res[mode] = (first, second, ev, nbh.newick_to_desc(nw1, taxa)[1], nbh.newick_to_desc(nw2, taxa)[1])
    a = res["host"]
    for mode in ("batch", "device"):   # the lockstep program as a batch of one; the single cooperative launch
        b = res[mode]
        np.testing.assert_allclose(b[:3], a[:3], rtol=1e-9)
        assert b[1] >= b[0] - 1e-9 * abs(b[0]) and abs(b[2] - b[1]) <= 1e-9 * abs(b[1])   # smoothing on, evaluate agrees
        np.testing.assert_allclose(b[3], a[3], rtol=1e-6, atol=1e-9)
        np.testing.assert_allclose(b[4], a[4], rtol=1e-6, atol=1e-9)
        assert not np.allclose(b[3], b[4])                                               # the second call moved the branches on
    pm.close(); ali.close()


def test_batched_getLogLikelihood_equals_per_tree_path(monkeypatch):
    """plg_pll_loglik_batch (all trees smoothed in lockstep on the device) against one handle per
    tree on the literal host loop: same lnL and branch lengths up to the summation order of the
    derivative sums, and the returned lnL equals the oracle's at the returned branch lengths (1e-8)."""
    monkeypatch.setenv("PLG_PLL_HANDLE_MODE", "host")
    from pylogeny_b200 import _lib, libpllWrapper as W, likelihood, neighbourhood as nbh, pll
    ali, d = _case(n=14, S=1200, seed=5)
    labels = ali.getTaxa()
    taxa = {l: i for i, l in enumerate(labels)}
    rng = np.random.default_rng(8)
    descs = [d] + [random_desc(rng, len(labels)) for _ in range(11)] + [d]
    newicks = [desc_to_newick(x, labels) for x in descs]
    pm = pll.partitionModel(ali)
    pm.createSimpleModel()
    per_tree, per_nw = [], []
    for nw in newicks:
        h = W.new(ali.getPhylip(), nw, pm.getFileName())
        per_tree.append(W.getLogLikelihood(h))
        per_nw.append(W.getNewickString(h))
        W.destroy(h)
    lnl, opt = W.getLogLikelihoodBatch(ali.getPhylip(), newicks, pm.getFileName())
    assert len(lnl) == len(opt) == len(newicks)
    np.testing.assert_allclose(lnl, per_tree, rtol=1e-9)
    assert lnl[0] == lnl[-1]                                    # the same tree twice in one batch
    codes = likelihood.encode_sequences(ali.toStrList(), _lib.DATA_DNA)
    h = W.new(ali.getPhylip(), newicks[0], pm.getFileName())
    K, P = C.c_int(), C.c_int64()
    freqs, rates = np.zeros(4), np.zeros(4)
    dp = C.POINTER(C.c_double)
    _lib.check(_lib.lib().plg_pll_info(h.ptr, C.byref(K), C.byref(P), freqs.ctypes.data_as(dp), rates.ctypes.data_as(dp)))
    W.destroy(h)
    w = np.ones(ali.getSize())
    for i in (0, 3, 7):
        want = _oracle_at(opt[i], taxa, codes, w, np.ones(6), freqs, 1.0)
        assert abs(lnl[i] - want) <= 1e-8 * abs(want)
        _, b_batch, _, _ = nbh.newick_to_desc(opt[i], taxa)
        _, b_tree, _, _ = nbh.newick_to_desc(per_nw[i], taxa)
        np.testing.assert_allclose(b_batch, b_tree, rtol=1e-6, atol=1e-9)
    one, _ = W.getLogLikelihoodBatch(ali.getPhylip(), newicks[:1], pm.getFileName(), passes=0)
    h = W.new(ali.getPhylip(), newicks[0], pm.getFileName())
    assert abs(one[0] - W.evaluate(h)) <= 1e-10 * abs(one[0])   # no smoothing: lnL at z = 0.9
    W.destroy(h)
    with pytest.raises(IOError):
        W.getLogLikelihoodBatch(ali.getPhylip(), ["(a,b,c);"], pm.getFileName())
    pm.close(); ali.close()


def test_batched_smoother_fused_reorientation_is_bit_identical(monkeypatch):
    """DNA batches of more than four trees run a re-orientation fused with the sum-table pass of the Newton
    step that follows it on the same vector (lik.h ClvOptOp); PLG_PLL_NO_FUSE=1 keeps the two launches.
    Same tile mapping, same arithmetic, same order of the sums: identical lnL and Newick strings."""
    from pylogeny_b200 import libpllWrapper as W, pll
    ali, d = _case(n=17, S=2100, seed=31)
    labels = ali.getTaxa()
    rng = np.random.default_rng(5)
    newicks = [desc_to_newick(random_desc(rng, 17), labels) for _ in range(11)]
    pm = pll.partitionModel(ali)
    pm.createSimpleModel()
    got = W.getLogLikelihoodBatch(ali.getPhylip(), newicks, pm.getFileName())
    monkeypatch.setenv("PLG_PLL_NO_FUSE", "1")
    want = W.getLogLikelihoodBatch(ali.getPhylip(), newicks, pm.getFileName())
    assert list(got[0]) == list(want[0]) and list(got[1]) == list(want[1])
    pm.close(); ali.close()


@pytest.mark.parametrize("n,S,signal,max_blocks", [(3, 200, True, 0), (4, 64, True, 0), (5, 300, True, 0), (9, 1500, True, 0),
                                                   (33, 2500, True, 0), (10, 12, False, 0), (10, 40, False, 0), (7, 25, False, 0),
                                                   (12, 55, False, 0), (9, 1500, True, 1), (33, 2500, True, 3), (14, 3000, False, 2),
                                                   (150, 60, True, 0), (180, 60, True, 0), (300, 200, True, 0),
                                                   (200, 300, True, 1)])
def test_single_launch_smoother_equals_host_loop(n, S, signal, max_blocks, monkeypatch):
    """The per-handle getLogLikelihood of a DNA tree -- ONE cooperative launch that walks the whole
    smoothing program with a grid barrier per Newton evaluation (pll_smooth1.cuh) -- against the literal
    per-branch host loop: smallest trees, alignments of a few patterns (one block), several blocks, and
    random sequences on few sites, where the bad-curvature guard retries and the z clamps do the work;
    max_blocks caps the grid (PLG_S1_MAX_BLOCKS) so that a block walks several 128-pattern tiles, as on
    alignments too long for one co-resident tile per block; 150 taxa: the action list no longer fits
    shared memory beside the matrices; 180 and more: the matrices do not all fit either, and the rest live in the
    blocks' own global regions.  lnL, branch lengths, and the oracle at the
    returned lengths."""
    if max_blocks:
        monkeypatch.setenv("PLG_S1_MAX_BLOCKS", str(max_blocks))
    from pylogeny_b200 import _lib, alignment, libpllWrapper as W, likelihood, neighbourhood as nbh, pll
    rng = np.random.default_rng(1000 * n + S)
    if signal:
        ali, d = _case(n=max(n, 4), S=S, seed=n + S)
        if n == 3:
            seqs = ali.toStrList()[:3]
            ali.close()
            ali = alignment.phylipFriendlyAlignment(names=["tax%d" % i for i in range(3)], seqs=seqs)
            d = np.array([[0, 1], [3, 2]], dtype=np.int32)
    else:
        seqs = ["".join(rng.choice(list("ACGT"), S)) for _ in range(n)]
        ali = alignment.phylipFriendlyAlignment(names=["tax%d" % i for i in range(n)], seqs=seqs)
        d = random_desc(rng, n)
    labels = ali.getTaxa()
    taxa = {l: i for i, l in enumerate(labels)}
    nw = desc_to_newick(d, labels)
    pm = pll.partitionModel(ali)
    pm.createSimpleModel()
    res = {}
    for mode in ("host", "device"):
        if mode == "host":
            monkeypatch.setenv("PLG_PLL_HANDLE_MODE", "host")
        else:
            monkeypatch.delenv("PLG_PLL_HANDLE_MODE")
        h = W.new(ali.getPhylip(), nw, pm.getFileName())
        lnl = W.getLogLikelihood(h)
        res[mode] = (lnl, W.getNewickString(h), W.evaluate(h))
        if mode == "device":
            K, P = C.c_int(), C.c_int64()
            freqs, rates = np.zeros(4), np.zeros(4)
            dp = C.POINTER(C.c_double)
            _lib.check(_lib.lib().plg_pll_info(h.ptr, C.byref(K), C.byref(P), freqs.ctypes.data_as(dp), rates.ctypes.data_as(dp)))
        W.destroy(h)
    (l0, nw0, e0), (l1, nw1, e1) = res["host"], res["device"]
    assert abs(l1 - l0) <= 1e-9 * abs(l0) and abs(e1 - l1) <= 1e-9 * abs(l1)
    np.testing.assert_allclose(nbh.newick_to_desc(nw1, taxa)[1], nbh.newick_to_desc(nw0, taxa)[1], rtol=1e-6, atol=1e-9)
    codes = likelihood.encode_sequences(ali.toStrList(), _lib.DATA_DNA)
    want = _oracle_at(nw1, taxa, codes, np.ones(ali.getSize()), np.ones(6), freqs, 1.0)
    assert abs(l1 - want) <= 1e-8 * abs(want)
    pm.close(); ali.close()


@pytest.mark.parametrize("n", [3, 4, 5, 9])
def test_batched_getLogLikelihood_small_trees_and_chunks(n, monkeypatch):
    """Smallest trees (one inner node at n = 3), a trifurcating input Newick, and a batch cut into
    several device chunks: same answers as one handle per tree (host loop)."""
    from pylogeny_b200 import libpllWrapper as W, pll
    monkeypatch.setenv("PLG_PLL_HANDLE_MODE", "host")
    ali, d = _case(n=max(n, 4), S=400, seed=20 + n)
    if n == 3:
        from pylogeny_b200 import alignment
        seqs = ali.toStrList()[:3]
        ali.close()
        ali = alignment.phylipFriendlyAlignment(names=["tax%d" % i for i in range(3)], seqs=seqs)
    labels = ali.getTaxa()
    rng = np.random.default_rng(n)
    if n == 3:
        newicks = ["(%s,%s,%s);" % tuple(labels), "((%s,%s),%s);" % (labels[2], labels[0], labels[1])]
    else:
        newicks = [desc_to_newick(random_desc(rng, n), labels) for _ in range(11)]
        newicks.append("(%s,%s,(%s));" % (labels[0], labels[1], ",".join(labels[2:])) if n == 4 else newicks[0])
        if n == 4:
            newicks[-1] = "(%s,%s,(%s,%s));" % tuple(labels)   # trifurcating root, as toUnrootedNewick writes
    pm = pll.partitionModel(ali)
    pm.createSimpleModel()
    want = []
    for nw in newicks:
        h = W.new(ali.getPhylip(), nw, pm.getFileName())
        want.append(W.getLogLikelihood(h))
        W.destroy(h)
    monkeypatch.setenv("PLG_PLL_BATCH_CHUNK", "5")
    got, opt = W.getLogLikelihoodBatch(ali.getPhylip(), newicks, pm.getFileName())
    np.testing.assert_allclose(got, want, rtol=1e-9)
    assert len(opt) == len(newicks) and all(s.endswith(";") for s in opt)
    assert W.getLogLikelihoodBatch(ali.getPhylip(), [], pm.getFileName()) == ([], [])
    pm.close(); ali.close()


def test_batched_getLogLikelihood_wag(monkeypatch):
    """The 20-state family of the batched smoother (clv20 / eval20 kernels, 80-entry sum tables)
    against one handle per tree (host loop)."""
    monkeypatch.setenv("PLG_PLL_HANDLE_MODE", "host")
    from pylogeny_b200 import alignment, libpllWrapper as W, pll
    rng = np.random.default_rng(44)
    aa = "ARNDCQEGHILKMFPSTWYV"
    n = 9
    base = rng.integers(0, 20, 300)
    seqs = []
    for _ in range(n):
        s = base.copy()
        mut = rng.random(300) < 0.25
        s[mut] = rng.integers(0, 20, mut.sum())
        seqs.append("".join(aa[i] for i in s))
    seqs[2] = seqs[2][:10] + "XB-Z" + seqs[2][14:]
    ali = alignment.phylipFriendlyAlignment(names=["p%d" % i for i in range(n)], seqs=seqs)
    labels = ali.getTaxa()
    newicks = [desc_to_newick(random_desc(rng, n), labels) for _ in range(7)]
    pm = pll.partitionModel(ali)
    pm.createSimpleModel()
    want, want_nw = [], []
    for nw in newicks:
        h = W.new(ali.getPhylip(), nw, pm.getFileName())
        want.append(W.getLogLikelihood(h))
        want_nw.append(W.getNewickString(h))
        W.destroy(h)
    got, opt = W.getLogLikelihoodBatch(ali.getPhylip(), newicks, pm.getFileName())
    np.testing.assert_allclose(got, want, rtol=1e-9)
    from pylogeny_b200 import neighbourhood as nbh
    taxa = {l: i for i, l in enumerate(labels)}
    for a, b in zip(opt, want_nw):
        np.testing.assert_allclose(nbh.newick_to_desc(a, taxa)[1], nbh.newick_to_desc(b, taxa)[1], rtol=1e-6, atol=1e-9)
    pm.close(); ali.close()


@pytest.mark.parametrize("seed", [1, 2, 3, 4])
def test_batched_getLogLikelihood_no_signal_data(seed, monkeypatch):
    """Random sequences on few sites: flat or wrongly curved likelihood surfaces on many branches,
    where makenewz's bad-curvature guard and the z clamps do the work.  Batch == per-tree host loop."""
    monkeypatch.setenv("PLG_PLL_HANDLE_MODE", "host")
    from pylogeny_b200 import alignment, libpllWrapper as W, pll
    rng = np.random.default_rng(100 + seed)
    n, S = 10, int(rng.integers(8, 60))
    seqs = ["".join("ACGT"[i] for i in rng.integers(0, 4, S)) for _ in range(n)]
    if seed % 2:
        seqs[3] = seqs[2]                      # identical sequences: a zero-length optimum (z -> zmax)
    ali = alignment.phylipFriendlyAlignment(names=["r%d" % i for i in range(n)], seqs=seqs)
    labels = ali.getTaxa()
    newicks = [desc_to_newick(random_desc(rng, n), labels) for _ in range(9)]
    pm = pll.partitionModel(ali)
    pm.createSimpleModel()
    want = []
    for nw in newicks:
        h = W.new(ali.getPhylip(), nw, pm.getFileName())
        want.append(W.getLogLikelihood(h))
        W.destroy(h)
    got, _ = W.getLogLikelihoodBatch(ali.getPhylip(), newicks, pm.getFileName())
    np.testing.assert_allclose(got, want, rtol=1e-9)
    pm.close(); ali.close()
