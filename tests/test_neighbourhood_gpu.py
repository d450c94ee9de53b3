"""Whole-neighbourhood scoring by edge insertion vs (a) full re-traversal of the materialised
neighbour trees on the GPU, (b) the CPU oracles, (c) golden numbers generated from the reference."""
import numpy as np
import pytest

import oracle
from util import desc_to_newick, random_desc, random_states, states_to_profiles

pytestmark = pytest.mark.gpu


def test_al_fasta_spr_and_nni_golden(al_profiles, al_fitch):
    """BASELINE config 1: al.fasta, every SPR / NNI neighbour of the file-order caterpillar."""
    from pylogeny_b200 import fitch, neighbourhood as nbh
    a = fitch.FitchAlignment(al_profiles["profiles"], al_profiles["weights"])
    named = {c["name"]: c for c in al_fitch["named"]}
    d, _, rows, tips = nbh.newick_to_desc(named["caterpillar_file_order"]["newick"], al_profiles["taxa"])
    assert a.score_trees(d[None], rows)[0] == 1167
    moves, costs = nbh.score_parsimony(a, d, rows, nbh.MOVE_SPR)
    assert len(costs) == 90 and costs.min() == 1149 and costs.max() == 1416 and int(costs.sum()) == 115698
    # the reference's own 120 SPR moves cover the same set of (topology -> score)
    ref_scores = sorted(set((m["cost"]) for m in al_fitch["spr_moves"]))
    assert sorted(set(costs.tolist())) == ref_scores
    _, nni = nbh.score_parsimony(a, d, rows, nbh.MOVE_NNI)
    assert sorted(nni.tolist()) == [1161, 1162, 1164, 1165, 1167, 1171, 1179, 1186, 1292, 1292]
    # every neighbour, materialised and re-scored from scratch, and through the oracle via Newick
    descs, _, _ = nbh.neighbour_trees(d, None, None, nbh.MOVE_SPR)
    assert a.score_trees(descs, rows).tolist() == costs.tolist()
    x = oracle.FitchInputs(al_profiles["profiles"], al_profiles["weights"], al_profiles["taxa"])
    want = [oracle.fitch_cost(nbh.desc_to_newick(dd, tips), inputs=x) for dd in descs]
    assert want == costs.tolist()
    a.close()


@pytest.mark.parametrize("n,P,ns", [(4, 50, 4), (9, 300, 5), (33, 2000, 4)])
def test_fitch_edge_insertion_vs_retraversal(n, P, ns):
    from pylogeny_b200 import fitch, neighbourhood as nbh
    rng = np.random.default_rng(n + P)
    states = random_states(rng, n, P, ns, 0.02)
    w = np.where(rng.random(P) < 0.3, rng.integers(1, 100, P), 1).astype(np.int64)
    a = fitch.FitchAlignment.from_dense(states, w)
    d = random_desc(rng, n)
    perm = rng.permutation(n).astype(np.int32)  # tips -> alignment rows
    moves, costs = nbh.score_parsimony(a, d, perm)
    descs, _, _ = nbh.neighbour_trees(d)
    full = a.score_trees(descs, perm)
    assert costs.tolist() == full.tolist()
    # shards tile the move list
    parts = [nbh.score_parsimony(a, d, perm, shard=(i, 3))[1] for i in range(3)]
    M = len(costs)
    merged = np.concatenate([parts[i][M * i // 3: M * (i + 1) // 3] for i in range(3)])
    assert merged.tolist() == costs.tolist()
    # a few against the oracle
    labels = ["T%d" % i for i in range(n)]
    taxa = {labels[i]: int(perm[i]) for i in range(n)}
    x = oracle.FitchInputs(states_to_profiles(states), w.tolist(), taxa)
    for i in range(0, M, max(1, M // 25)):
        assert oracle.fitch_cost(desc_to_newick(descs[i], labels), inputs=x) == costs[i]
    a.close()


@pytest.mark.parametrize("n,P", [(4, 64), (7, 500), (20, 1500)])
def test_lnl_edge_insertion_vs_retraversal_and_oracle(n, P):
    from pylogeny_b200 import likelihood, neighbourhood as nbh
    rng = np.random.default_rng(7 * n + P)
    codes = (1 << rng.integers(0, 4, (n, P))).astype(np.uint8)
    codes = np.where(rng.random((n, P)) < 0.03, 15, codes).astype(np.uint8)
    w = rng.integers(1, 5, P).astype(float)
    exch, f = rng.uniform(0.3, 3, 6), rng.dirichlet(np.ones(4) * 6)
    a = likelihood.LikAlignment(codes, w)
    m = likelihood.Model(exch, f, 0.7)
    d = random_desc(rng, n)
    br = rng.exponential(0.1, d.shape)
    moves, lnl = nbh.score_likelihood(a, m, d, br)
    descs, brs, _ = nbh.neighbour_trees(d, br)
    full = a.score_trees(m, descs, brs)
    np.testing.assert_allclose(lnl, full, rtol=1e-11)
    M = len(lnl)
    for i in range(0, M, max(1, M // 12)):
        want = oracle.lnl(codes, w, exch, f, 0.7, descs[i], brs[i])
        assert abs(lnl[i] - want) <= 1e-8 * abs(want)
    _, nni = nbh.score_likelihood(a, m, d, br, move_type=nbh.MOVE_NNI)
    assert len(nni) == 2 * (n - 3)
    a.close(); m.close()


@pytest.mark.parametrize("n,P", [(5, 100), (9, 400), (14, 900)])
def test_tbr_edge_insertion_vs_retraversal(n, P):
    """TBR: part roots joined pairwise == full re-traversal of the materialised neighbour (Fitch exact, lnL 1e-11)."""
    from pylogeny_b200 import fitch, likelihood, neighbourhood as nbh
    rng = np.random.default_rng(31 * n + P)
    states = random_states(rng, n, P, 4, 0.03)
    w = np.where(rng.random(P) < 0.3, rng.integers(1, 50, P), 1).astype(np.int64)
    fa = fitch.FitchAlignment.from_dense(states, w)
    d = random_desc(rng, n)
    br = rng.exponential(0.1, d.shape)
    moves, costs = nbh.score_parsimony(fa, d, None, nbh.MOVE_TBR)
    descs, brs, _ = nbh.neighbour_trees(d, br, None, nbh.MOVE_TBR)
    assert len(costs) == len(descs) > 0
    assert fa.score_trees(descs).tolist() == costs.tolist()
    codes = (1 << rng.integers(0, 4, (n, P))).astype(np.uint8)
    la = likelihood.LikAlignment(codes, w.astype(float))
    m = likelihood.Model(rng.uniform(0.5, 2, 6), rng.dirichlet(np.ones(4) * 5), 0.8)
    _, lnl = nbh.score_likelihood(la, m, d, br, move_type=nbh.MOVE_TBR)
    np.testing.assert_allclose(lnl, la.score_trees(m, descs, brs), rtol=1e-11)
    parts = [nbh.score_likelihood(la, m, d, br, move_type=nbh.MOVE_TBR, shard=(i, 2))[1] for i in range(2)]
    M = len(lnl)
    np.testing.assert_array_equal(np.concatenate([parts[0][:M // 2], parts[1][M // 2:]]), lnl)
    fa.close(); la.close(); m.close()


@pytest.mark.parametrize("n,P,scale", [(4, 30, 1.0), (5, 64, 1.0), (9, 333, 1.0), (16, 700, 1.0), (30, 150, 40.0), (60, 90, 1.0)])
def test_tbr_20_states_fused_steps_and_pair_grid(n, P, scale, monkeypatch):
    """20 states x Gamma-4 TBR neighbourhoods: the fused candidate steps + DMMA pair grid (lik_tbr20.cuh,
    default) against the level program of plain updates and row evaluations (PLG_TBR20=0), against the
    re-traversal of materialised neighbours and against the oracle -- ambiguity codes, weights (also
    zero), the underflow regime (long branches: 2^-256 factors inside a fused step), shards."""
    from pylogeny_b200 import _lib, likelihood, neighbourhood as nbh
    rng = np.random.default_rng(7 * n + P)
    codes = rng.integers(0, 20, (n, P)).astype(np.uint8)
    codes = np.where(rng.random((n, P)) < 0.03, rng.integers(20, 23, (n, P)), codes).astype(np.uint8)
    w = rng.integers(0, 4, P).astype(float)
    exch, f = likelihood.protein_model("WAG") if n % 2 else (rng.uniform(0.05, 5.0, 190), rng.dirichlet(np.ones(20) * 10))
    a = likelihood.LikAlignment(codes, w, _lib.DATA_AA)
    m = likelihood.Model(exch, f, 0.8)
    d = random_desc(rng, n)
    br = rng.exponential(0.1 * scale, d.shape)
    br[rng.random(br.shape) < 0.1] = 1e-9   # (an exact zero between two different tip states is lnL = -inf)
    moves, lnl = nbh.score_likelihood(a, m, d, br, move_type=nbh.MOVE_TBR)
    M = len(lnl)
    assert M == nbh.neighbourhood_size(d, nbh.MOVE_TBR) and np.all(np.isfinite(lnl)) and np.all(lnl < 0)
    monkeypatch.setenv("PLG_TBR20", "0")
    moves0, lnl0 = nbh.score_likelihood(a, m, d, br, move_type=nbh.MOVE_TBR)
    monkeypatch.delenv("PLG_TBR20")
    assert moves.tolist() == moves0.tolist()
    np.testing.assert_allclose(lnl, lnl0, rtol=1e-12)
    sel = np.unique(np.linspace(0, M - 1, 24).astype(np.int64)) if M else np.zeros(0, np.int64)
    descs, brs, _ = nbh.neighbour_trees(d, br, sel, nbh.MOVE_TBR)
    if M:
        np.testing.assert_allclose(lnl[sel], a.score_trees(m, descs, brs), rtol=1e-11)
    for j in range(0, len(sel), 5):
        want = oracle.lnl(codes, w, exch, f, 0.8, descs[j], brs[j])
        assert abs(lnl[sel[j]] - want) <= 1e-8 * abs(want)
    if M >= 3:
        parts = [nbh.score_likelihood(a, m, d, br, move_type=nbh.MOVE_TBR, shard=(i, 3), want_moves=False)[1] for i in range(3)]
        got = np.concatenate([parts[i][M * i // 3:M * (i + 1) // 3] for i in range(3)])
        np.testing.assert_array_equal(got, lnl)
        # the neighbourhood cut into chunks of bisections that fit a (tiny) memory budget: same numbers
        for mode in ("1", "0"):
            monkeypatch.setenv("PLG_TBR20", mode)
            monkeypatch.setenv("PLG_TBR_MAX_SLOTS", str(12 * n))
            _, chunked = nbh.score_likelihood(a, m, d, br, move_type=nbh.MOVE_TBR, want_moves=False)
            monkeypatch.delenv("PLG_TBR_MAX_SLOTS")
            np.testing.assert_array_equal(chunked, lnl if mode == "1" else lnl0)
        monkeypatch.delenv("PLG_TBR20")
    a.close(); m.close()


@pytest.mark.parametrize("n,P,scale,wmax", [(6, 40, 1.0, 4), (24, 700, 1.0, 4), (70, 300, 1.0, 4), (40, 200, 60.0, 4),
                                            (24, 700, 1.0, 1), (70, 300, 1.0, 2), (40, 200, 60.0, 1), (9, 33, 200.0, 1)])
def test_lnl_planners_agree(n, P, scale, wmax, monkeypatch):
    """The DNA x Gamma-4 neighbourhood planners (PLG_LIK_NB_MODE=ops: one level-synchronous update +
    evaluation per neighbour; default: depth-first search walks on the FP64 tensor pipe) give the
    same lnL per neighbour, also in the underflow regime (long branches => 2^-256 rescaling inside
    the walk), for shards and for radius-limited neighbourhoods; walk == oracle on a sample.
    wmax = 1: unit weights, every slice takes the walk's mantissa/exponent path (no log in the
    kernel); wmax = 2: weights 0/1 mixed with heavier ones, so both paths run in one call; wmax = 4:
    weighted slices only."""
    from pylogeny_b200 import likelihood, neighbourhood as nbh
    rng = np.random.default_rng(11 * n + P)
    codes = (1 << rng.integers(0, 4, (n, P))).astype(np.uint8)
    codes = np.where(rng.random((n, P)) < 0.02, 15, codes).astype(np.uint8)
    w = rng.integers(1, 4, P).astype(float) if wmax == 4 else np.ones(P)
    if wmax == 2:  # some 32-pattern slices all 0/1, others weighted
        w[rng.random(P) < 0.1] = 0.0
        w[P // 2:][rng.random(P - P // 2) < 0.2] = 3.0
    exch, f = rng.uniform(0.3, 3, 6), rng.dirichlet(np.ones(4) * 6)
    a = likelihood.LikAlignment(codes, w)
    m = likelihood.Model(exch, f, 0.6)
    d = random_desc(rng, n)
    br = rng.exponential(0.1 * scale, d.shape)
    res = {}
    for mode in ("ops", "walkm"):
        monkeypatch.setenv("PLG_LIK_NB_MODE", mode)
        res[mode] = nbh.score_likelihood(a, m, d, br)[1]
    np.testing.assert_allclose(res["walkm"], res["ops"], rtol=1e-12)
    res["walk"] = res["walkm"]
    M = len(res["walk"])
    parts = [nbh.score_likelihood(a, m, d, br, shard=(i, 3))[1] for i in range(3)]
    merged = np.concatenate([parts[i][M * i // 3: M * (i + 1) // 3] for i in range(3)])
    np.testing.assert_array_equal(merged, res["walk"])
    descs, brs, _ = nbh.neighbour_trees(d, br)
    for i in range(0, M, max(1, M // 10)):
        want = oracle.lnl(codes, w, exch, f, 0.6, descs[i], brs[i])
        assert abs(res["walk"][i] - want) <= 1e-8 * abs(want)
    monkeypatch.delenv("PLG_LIK_NB_MODE")
    np.testing.assert_array_equal(nbh.score_likelihood(a, m, d, br)[1], res["walkm"])  # walkm is the default
    mt = nbh.MOVE_SPR_WITHIN(3)
    _, near = nbh.score_likelihood(a, m, d, br, move_type=mt)
    _, _, near_h = nbh.neighbour_trees(d, br, None, mt)
    _, _, all_h = nbh.neighbour_trees(d, br)
    by_hash = dict(zip(all_h.tolist(), res["walk"].tolist()))
    np.testing.assert_allclose(near, [by_hash[h] for h in near_h.tolist()], rtol=1e-12)
    a.close(); m.close()


@pytest.mark.parametrize("n,P,states,gaps", [(4, 30, 4, 0.0), (6, 200, 4, 0.1), (33, 1500, 4, 0.02), (90, 700, 4, 0.0),
                                              (25, 300, 9, 0.0), (18, 260, 18, 0.0)])
def test_fitch_search_walks_equal_level_ops(n, P, states, gaps, monkeypatch):
    """Parsimony of a neighbourhood by depth-first search walks (default) == one level-synchronous op
    per search step (PLG_FITCH_NB_MODE=ops) == full re-traversal of the materialised neighbours,
    bit for bit; also for shards, NNI and radius-limited SPR."""
    from pylogeny_b200 import fitch, neighbourhood as nbh
    rng = np.random.default_rng(13 * n + P)
    st = random_states(rng, n, P, states, gaps)
    w = np.where(rng.random(P) < 0.3, rng.integers(1, 3000, P), 1).astype(np.int64)
    a = fitch.FitchAlignment.from_dense(st, w)
    d = random_desc(rng, n)
    res = {}
    for mode in ("walk", "ops"):
        monkeypatch.setenv("PLG_FITCH_NB_MODE", mode)
        res[mode] = {mt: nbh.score_parsimony(a, d, None, mt)[1] for mt in (nbh.MOVE_SPR, nbh.MOVE_NNI, nbh.MOVE_SPR_WITHIN(3))}
    for mt in res["walk"]:
        assert res["walk"][mt].tolist() == res["ops"][mt].tolist()
    monkeypatch.setenv("PLG_FITCH_NB_MODE", "walk")
    spr = res["walk"][nbh.MOVE_SPR]
    descs, _, _ = nbh.neighbour_trees(d, None)
    assert a.score_trees(descs).tolist() == spr.tolist()
    M = len(spr)
    parts = [nbh.score_parsimony(a, d, None, nbh.MOVE_SPR, shard=(i, 3))[1] for i in range(3)]
    merged = np.concatenate([parts[i][M * i // 3: M * (i + 1) // 3] for i in range(3)])
    assert merged.tolist() == spr.tolist()
    perm = rng.permutation(n).astype(np.int32)
    monkeypatch.setenv("PLG_FITCH_NB_MODE", "ops")
    want = nbh.score_parsimony(a, d, perm)[1]
    monkeypatch.setenv("PLG_FITCH_NB_MODE", "walk")
    assert nbh.score_parsimony(a, d, perm)[1].tolist() == want.tolist()
    a.close()


@pytest.mark.parametrize("n,P,shape", [(4, 40, "random"), (5, 64, "random"), (12, 333, "random"), (40, 1000, "caterpillar"),
                                        (64, 900, "balanced"), (130, 520, "random"), (257, 300, "random")])
def test_fitch_device_planned_spr_equals_host_planner(n, P, shape, monkeypatch):
    """Full SPR neighbourhoods are laid out on the device (spr_plan_kernel, default) from O(n) host
    records; the host planner (PLG_FITCH_NB_PLAN=host: enumerate_spr + build_fitch_nb_walk_program)
    must give the same move list (every field, same order) and the same costs, bit for bit."""
    from pylogeny_b200 import fitch, neighbourhood as nbh
    rng = np.random.default_rng(7 * n + P)
    st = random_states(rng, n, P, 4, 0.03)
    w = np.where(rng.random(P) < 0.4, rng.integers(1, 900, P), 1).astype(np.int64)
    a = fitch.FitchAlignment.from_dense(st, w)
    if shape == "caterpillar":
        d = np.array([(0, 1)] + [(n + i - 1, i + 1) for i in range(1, n - 1)], dtype=np.int32)
    elif shape == "balanced":
        ids = list(range(n))
        rows = []
        while len(ids) > 1:
            nxt = []
            for i in range(0, len(ids) - 1, 2):
                rows.append((ids[i], ids[i + 1]))
                nxt.append(n + len(rows) - 1)
            if len(ids) % 2:
                nxt.append(ids[-1])
            ids = nxt
        d = np.array(rows, dtype=np.int32)
    else:
        d = random_desc(rng, n)
    perm = rng.permutation(n).astype(np.int32)
    got = {}
    for plan in ("device", "host"):
        monkeypatch.setenv("PLG_FITCH_NB_PLAN", plan)
        got[plan] = nbh.score_parsimony(a, d, perm)
    assert got["device"][0].shape == got["host"][0].shape == (nbh.neighbourhood_size(d), 4)
    assert got["device"][0].tolist() == got["host"][0].tolist()
    assert got["device"][1].tolist() == got["host"][1].tolist()
    monkeypatch.setenv("PLG_FITCH_NB_PLAN", "device")
    M = len(got["device"][1])
    for count in (2, 7):
        parts = [nbh.score_parsimony(a, d, perm, shard=(i, count)) for i in range(count)]
        merged = np.concatenate([parts[i][1][M * i // count: M * (i + 1) // count] for i in range(count)])
        assert merged.tolist() == got["device"][1].tolist(), count
        assert all(p[0].tolist() == got["device"][0].tolist() for p in parts)
    descs, _, _ = nbh.neighbour_trees(d)
    assert a.score_trees(descs, perm).tolist() == got["device"][1].tolist()
    a.close()


@pytest.mark.parametrize("n,P", [(4, 40), (5, 70), (13, 333), (64, 500), (131, 200)])
def test_lnl_device_planned_spr_equals_host_planner(n, P, monkeypatch):
    """GTR+G4 lnL of a full SPR neighbourhood: search walks laid out on the device
    (spr_lik_plan_kernel, default) vs by the host planner (PLG_LIK_NB_PLAN=host) -- same move table,
    bit-identical lnL, and the same byte accounting in the profile records."""
    import ctypes as C
    from pylogeny_b200 import _lib, likelihood, neighbourhood as nbh
    rng = np.random.default_rng(11 * n + P)
    codes = (1 << rng.integers(0, 4, (n, P))).astype(np.uint8)
    codes[rng.random((n, P)) < 0.02] = 15
    w = rng.integers(1, 4, P).astype(np.float64)
    a = likelihood.LikAlignment(codes, w)
    m = likelihood.Model(rng.uniform(0.5, 2.0, 6), rng.dirichlet(np.ones(4) * 8), 0.6)
    d = random_desc(rng, n)
    br = rng.exponential(0.1, d.shape)
    perm = rng.permutation(n).astype(np.int32)
    L = _lib.lib()
    got, acct = {}, {}
    for plan in ("device", "host"):
        monkeypatch.setenv("PLG_LIK_NB_PLAN", plan)
        L.plg_profile_enable(1)
        L.plg_profile_read(4, None, None, None, None, None)
        got[plan] = nbh.score_likelihood(a, m, d, br, perm)
        ms, by, un, nl, rq = C.c_double(), C.c_double(), C.c_double(), C.c_int64(), C.c_double()
        L.plg_profile_read(4, C.byref(ms), C.byref(by), C.byref(un), C.byref(nl), C.byref(rq))
        L.plg_profile_enable(0)
        acct[plan] = (by.value, un.value, nl.value, rq.value)
        assert nbh.score_likelihood(a, m, d, br, perm, want_moves=False)[1].tolist() == got[plan][1].tolist()
    assert got["device"][0].tolist() == got["host"][0].tolist()
    assert got["device"][1].tolist() == got["host"][1].tolist()  # bit-identical
    assert acct["device"] == pytest.approx(acct["host"], rel=1e-12)
    # shards of the neighbourhood: each plans and walks only the searches that hold its ids
    M = len(got["device"][1])
    for count in (2, 5):
        for plan in ("device", "host"):
            monkeypatch.setenv("PLG_LIK_NB_PLAN", plan)
            parts = [nbh.score_likelihood(a, m, d, br, perm, shard=(i, count)) for i in range(count)]
            merged = np.concatenate([parts[i][1][M * i // count: M * (i + 1) // count] for i in range(count)])
            assert merged.tolist() == got["device"][1].tolist(), (plan, count)
            assert all(p[0].tolist() == got["device"][0].tolist() for p in parts)   # the move table is whole in every shard
    a.close(); m.close()


def test_device_planner_global_memory_fallback(monkeypatch):
    """Edge records too large for shared memory are read from global memory (forced here): same results."""
    from pylogeny_b200 import fitch, likelihood, neighbourhood as nbh
    rng = np.random.default_rng(99)
    n, P = 37, 260
    st = random_states(rng, n, P, 4, 0.02)
    fa = fitch.FitchAlignment.from_dense(st, np.ones(P, dtype=np.int64))
    codes = (1 << rng.integers(0, 4, (n, P))).astype(np.uint8)
    la = likelihood.LikAlignment(codes, np.ones(P))
    m = likelihood.Model(rng.uniform(0.5, 2.0, 6), rng.dirichlet(np.ones(4) * 8), 0.8)
    d = random_desc(rng, n)
    br = rng.exponential(0.1, d.shape)
    want_f = nbh.score_parsimony(fa, d)
    want_l = nbh.score_likelihood(la, m, d, br)
    monkeypatch.setenv("PLG_SPR_PLAN_GLOBAL", "1")
    got_f = nbh.score_parsimony(fa, d)
    got_l = nbh.score_likelihood(la, m, d, br)
    assert got_f[0].tolist() == want_f[0].tolist() and got_f[1].tolist() == want_f[1].tolist()
    assert got_l[0].tolist() == want_l[0].tolist() and got_l[1].tolist() == want_l[1].tolist()
    fa.close(); la.close(); m.close()


def test_device_planner_many_small_trees(monkeypatch):
    """Every tree size from 4 to 24 tips, three random shapes each: device-planned neighbourhoods
    (parsimony and GTR+G4) equal the host-planned ones, move tables included."""
    from pylogeny_b200 import fitch, likelihood, neighbourhood as nbh
    rng = np.random.default_rng(2024)
    P = 96
    for n in range(4, 25):
        st = random_states(rng, n, P, 4, 0.05)
        w = rng.integers(1, 6, P).astype(np.int64)
        fa = fitch.FitchAlignment.from_dense(st, w)
        codes = (1 << rng.integers(0, 4, (n, P))).astype(np.uint8)
        la = likelihood.LikAlignment(codes, w.astype(float))
        m = likelihood.Model(rng.uniform(0.5, 2.0, 6), rng.dirichlet(np.ones(4) * 8), 0.9)
        for _ in range(3):
            d = random_desc(rng, n)
            br = rng.exponential(0.1, d.shape)
            res = {}
            for plan in ("device", "host"):
                monkeypatch.setenv("PLG_FITCH_NB_PLAN", plan)
                monkeypatch.setenv("PLG_LIK_NB_PLAN", plan)
                res[plan] = (nbh.score_parsimony(fa, d), nbh.score_likelihood(la, m, d, br))
            for k in (0, 1):
                assert res["device"][k][0].tolist() == res["host"][k][0].tolist(), (n, k)
                assert res["device"][k][1].tolist() == res["host"][k][1].tolist(), (n, k)
            assert len(res["device"][0][1]) == nbh.neighbourhood_size(d)
        fa.close(); la.close(); m.close()
