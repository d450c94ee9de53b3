import os, sys, numpy as np
sys.path.insert(0, '/root/repo'); sys.path.insert(0, '/root/repo/tests')
from util import random_desc, random_states
from pylogeny_b200 import fitch, neighbourhood as nbh
for (n, P, ns) in [(9, 300, 5), (9, 300, 4), (9, 100, 4), (6, 100, 4), (12, 100, 4)]:
    rng = np.random.default_rng(n + P)
    states = random_states(rng, n, P, ns, 0.02)
    w = np.where(rng.random(P) < 0.3, rng.integers(1, 100, P), 1).astype(np.int64)
    a = fitch.FitchAlignment.from_dense(states, w)
    d = random_desc(rng, n)
    os.environ["PLG_FITCH_NB_MODE"] = "walk"
    moves, costs = nbh.score_parsimony(a, d)
    os.environ["PLG_FITCH_NB_MODE"] = "ops"
    _, ops = nbh.score_parsimony(a, d)
    descs, _, _ = nbh.neighbour_trees(d)
    os.environ["PLG_TREE_MODE"] = "walk"
    full = a.score_trees(descs)
    os.environ["PLG_TREE_MODE"] = "levels"
    lev = a.score_trees(descs)
    bad = np.nonzero(costs != ops)[0]
    print(n, P, ns, "walk!=ops at", bad.tolist(), "ops!=full", np.nonzero(ops != full)[0].tolist(), "full!=lev", np.nonzero(full != lev)[0].tolist())
    for i in bad[:12]:
        print("   move", i, moves[i].tolist(), "walk", costs[i], "ops", ops[i])
    print("   desc", d.tolist())
