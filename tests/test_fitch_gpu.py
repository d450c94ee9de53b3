"""Parity of the CUDA Fitch path (through the C-ABI) with the reference: golden fixtures
generated from the reference's own code, and the pinned C oracle on seeded random inputs."""
import numpy as np
import pytest

import oracle
from util import desc_to_newick, random_desc, random_kary_newick, random_states, states_to_profiles

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def al(al_profiles):
    from pylogeny_b200 import fitch
    a = fitch.FitchAlignment(al_profiles["profiles"], al_profiles["weights"])
    yield a
    a.close()


def test_named_golden(al, al_profiles, al_fitch):
    got = al.score_newicks([c["newick"] for c in al_fitch["named"]], al_profiles["taxa"])
    assert got.tolist() == [c["cost"] for c in al_fitch["named"]]


def test_calculateCost_compat(al_profiles, al_fitch):
    from pylogeny_b200 import fitch
    for c in al_fitch["named"]:
        assert fitch.calculateCost(c["newick"], al_profiles["profiles"], al_profiles["weights"],
                                   al_profiles["taxa"]) == c["cost"]


def test_reference_spr_moves_golden(al, al_profiles, al_fitch):
    got = al.score_newicks([m["newick"] for m in al_fitch["spr_moves"]], al_profiles["taxa"])
    assert got.tolist() == [m["cost"] for m in al_fitch["spr_moves"]]


def test_reference_greedy_calls_golden(al, al_profiles, al_greedy):
    got = al.score_newicks([c[0] for c in al_greedy["calls"]], al_profiles["taxa"])
    assert got.tolist() == [c[1] for c in al_greedy["calls"]]


def test_random_golden_one_shot(random_fitch):
    """60 reference-scored cases: k-ary trees, permuted taxa maps, huge weights (int32 wrap)."""
    from pylogeny_b200 import fitch
    for c in random_fitch:
        assert fitch.calculateCost(c["newick"], c["profiles"], c["weights"], c["taxa"]) == c["cost"], c["newick"]


@pytest.mark.parametrize("n,P,ns,gap", [(5, 1, 2, 0.0), (17, 129, 4, 0.05), (64, 5000, 4, 0.01), (40, 777, 9, 0.1)])
def test_descriptor_batch_vs_oracle(n, P, ns, gap):
    from pylogeny_b200 import fitch
    rng = np.random.default_rng(n * 1000 + P)
    states = random_states(rng, n, P, ns, gap)
    w = np.where(rng.random(P) < 0.3, rng.integers(1, 2000, P), 1).astype(np.int64)
    labels = ["T%d" % i for i in range(n)]
    taxa = {l: i for i, l in enumerate(labels)}
    descs = np.stack([random_desc(rng, n) for _ in range(12)])
    a = fitch.FitchAlignment.from_dense(states, w)
    got = a.score_trees(descs)
    x = oracle.FitchInputs(states_to_profiles(states), w.tolist(), taxa)
    want = [oracle.fitch_cost(desc_to_newick(d, labels), inputs=x) for d in descs]
    assert got.tolist() == want
    # same trees through the Newick entry point
    assert a.score_newicks([desc_to_newick(d, labels) for d in descs], taxa).tolist() == want
    a.close()


def test_kary_vs_oracle():
    from pylogeny_b200 import fitch
    rng = np.random.default_rng(99)
    n, P = 30, 1000
    states = random_states(rng, n, P, 5, 0.0)
    w = rng.integers(1, 5, P).astype(np.int64)
    labels = ["T%d" % i for i in range(n)]
    taxa = {l: i for i, l in enumerate(labels)}
    nws = [random_kary_newick(rng, labels, 0.5) for _ in range(20)]
    a = fitch.FitchAlignment.from_dense(states, w)
    x = oracle.FitchInputs(states_to_profiles(states), w.tolist(), taxa)
    assert a.score_newicks(nws, taxa).tolist() == [oracle.fitch_cost(s, inputs=x) for s in nws]
    a.close()


def test_edge_cases():
    from pylogeny_b200 import fitch
    # no patterns at all: every tree costs 0 (fitch.cpp:249 loop body never runs)
    assert fitch.calculateCost("(A,B);", [], [], {"A": 0, "B": 1}) == 0
    # single leaf, two leaves
    assert fitch.calculateCost("A;", ["0"], [3], {"A": 0}) == 0
    assert fitch.calculateCost("(A,B);", ["01", "00"], [3, 5], {"A": 0, "B": 1}) == 3
    # unknown label / mismatched lengths follow the documented error mapping
    with pytest.raises(KeyError):
        fitch.calculateCost("(A,C);", ["01"], [1], {"A": 0, "B": 1})
    with pytest.raises(SystemError):
        fitch.calculateCost("(A,B);", ["01"], [1, 2], {"A": 0, "B": 1})
    # zero and negative weights, 32-bit wrap (fitch.cpp:247,268)
    x = oracle.FitchInputs(["01", "01"], [0, -7], {"A": 0, "B": 1})
    assert fitch.calculateCost("(A,B);", ["01", "01"], [0, -7], {"A": 0, "B": 1}) == oracle.fitch_cost("(A,B);", inputs=x) == -7
    big = [2 ** 31 - 1, 2 ** 31 - 1, 5]
    x = oracle.FitchInputs(["01"] * 3, big, {"A": 0, "B": 1})
    assert fitch.calculateCost("(A,B);", ["01"] * 3, big, {"A": 0, "B": 1}) == oracle.fitch_cost("(A,B);", inputs=x)


def test_full_size_properties():
    """BASELINE config 3 shape (256 taxa x 100k sites): re-rooting and child-order invariance of
    binary trees (size-independent properties) plus one tree against the oracle."""
    from pylogeny_b200 import fitch
    rng = np.random.default_rng(1003)
    n, P = 256, 100_000
    states = (rng.integers(0, 4, (P, n)) + 48).astype(np.uint8)
    w = np.ones(P, dtype=np.int64)
    a = fitch.FitchAlignment.from_dense(states, w)
    d = random_desc(rng, n)
    flipped = d[:, ::-1].copy()
    # re-root: hang the old root's two subtrees differently -- move the root onto child 0's edge
    r = len(d) - 1
    a0, b0 = d[r]
    rer = d.copy()
    if a0 >= n:  # (X,Y) root with X=(p,q): new root (p,(q,Y))
        p, q = d[a0 - n]
        rer[a0 - n] = (q, b0)
        rer[r] = (p, a0)
    got = a.score_trees(np.stack([d, flipped, rer]))
    assert got[0] == got[1] == got[2]
    labels = ["T%d" % i for i in range(n)]
    x = oracle.FitchInputs(states_to_profiles(states), w.tolist(), {l: i for i, l in enumerate(labels)})
    assert int(got[0]) == oracle.fitch_cost(desc_to_newick(d, labels), inputs=x)
    a.close()
