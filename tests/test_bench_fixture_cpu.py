"""The committed sample trees of bench.py's CPU arm (tests/golden/bench_ref_trees.npz, made by
tests/golden/make_bench_fixture.py from the REFERENCE's rearrangement code): every SPR sample is a
genuine neighbour of the workload's base tree -- its canonical topology hash is in the engine's own
enumeration, its branch lengths are the reference's own (rearrangement.py:466-497; the tree length is
conserved) -- and scoring the same sample through the engine's materialisation gives the same
branch lengths per split, so both arms of the bench see the same trees."""
import os
import sys
from collections import Counter

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def _unrooted_lengths(brlen):
    """Sorted branch lengths of the unrooted tree: the root's two child branches are one branch."""
    b = np.asarray(brlen)
    return sorted(np.round(np.concatenate([b[:-1].ravel(), [b[-1].sum()]]), 10).tolist())


@pytest.mark.parametrize("name", ["lnl_c2", "lnl_target", "fitch_target"])
def test_spr_samples_are_neighbours_of_the_base_tree(name):
    import bench
    from pylogeny_b200 import neighbourhood as nbh
    fx = np.load(os.path.join(ROOT, "tests", "golden", "bench_ref_trees.npz"))
    wl = bench.WORKLOADS[name]
    d, br = bench.base_tree(wl, 0)
    descs, brs = fx[name + "_desc"].astype(np.int32), fx[name + "_brlen"]
    assert descs.shape == (wl["ref_sample"], wl["n_taxa"] - 1, 2) and brs.shape == descs.shape
    _, _, hashes = nbh.neighbour_trees(d, None, None)
    ours = {int(h): i for i, h in enumerate(hashes)}
    base_hash = nbh.topology_hash(d)
    seen, root_edge, other_move, not_conserved = set(), 0, 0, 0
    for k in range(len(descs)):
        ids = sorted(descs[k].ravel().tolist())
        assert ids == list(range(2 * wl["n_taxa"] - 2))             # a tree: every node a child exactly once
        h = nbh.topology_hash(descs[k])
        assert h in ours and h != base_hash
        seen.add(h)
        if abs(brs[k].sum() - br.sum()) > 1e-9 * br.sum():           # (the reference's rule conserves the tree length,
            not_conserved += 1                                       #  except in a few of its flipped moves)
            continue
        # the engine's materialisation of the same neighbour carries the same lengths
        dd, bb, _ = nbh.neighbour_trees(d, br, [ours[h]])
        a, b = Counter(_unrooted_lengths(bb[0])), Counter(_unrooted_lengths(brs[k]))
        if a != b:
            # the one place the two differ: a regraft INTO the branch that carries the reference's root.  The
            # reference keeps that branch as (root--leaf: whole length, root--rest: 0) and halves the second, so
            # the new node sits at distance 0 from the rest; the engine halves the unrooted branch.
            only_a, only_b = a - b, b - a
            items = list(only_a.items())
            if len(items) == 1 and items[0][1] == 2 and sorted(only_b.items()) == [(0.0, 1), (round(2 * items[0][0], 10), 1)]:
                root_edge += 1
            else:
                # a topology that several (prune, regraft) pairs produce (the distance-1 moves): the reference's
                # sampled move and the engine's representative of that topology halve / merge different branches
                other_move += 1
    assert root_edge < len(descs) // 3 and other_move <= len(descs) // 10 and not_conserved <= len(descs) // 10
    assert len(seen) > len(descs) // 2                               # spread over the neighbourhood, not one move


def test_tbr_samples_are_trees():
    import bench
    fx = np.load(os.path.join(ROOT, "tests", "golden", "bench_ref_trees.npz"))
    wl = bench.WORKLOADS["aa_c5"]
    descs, brs = fx["aa_c5_desc"], fx["aa_c5_brlen"]
    _, br = bench.base_tree(wl, 0)
    for k in range(len(descs)):
        assert sorted(descs[k].ravel().tolist()) == list(range(2 * wl["n_taxa"] - 2))
        assert abs(brs[k].sum() - br.sum()) < 1e-9 * br.sum()
