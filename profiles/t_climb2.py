import os, sys, time, numpy as np
sys.path.insert(0, '/root/repo'); sys.path.insert(0, '/root/repo/tests')
from util import random_desc
from pylogeny_b200 import fitch, neighbourhood as nbh
n, S = 256, 100_000
rng = np.random.default_rng(1002)
states = (rng.integers(0, 4, (S, n)) + 48).astype(np.uint8)
a = fitch.FitchAlignment.from_dense(states, np.ones(S, dtype=np.int64))
d = random_desc(np.random.default_rng(1002 + 7919), n)
nbh.score_parsimony(a, d)
T = {"score": 0.0, "argmin": 0.0, "trees": 0.0}
steps = 0
cur = int(a.score_trees(d[None])[0])
for _ in range(12):
    t0 = time.perf_counter(); _, costs = nbh.score_parsimony(a, d, want_moves=False); t1 = time.perf_counter()
    best = int(np.argmin(costs)); t2 = time.perf_counter()
    if int(costs[best]) >= cur: break
    descs, _, _ = nbh.neighbour_trees(d, None, [best]); t3 = time.perf_counter()
    d, cur = descs[0], int(costs[best]); steps += 1
    T["score"] += t1 - t0; T["argmin"] += t2 - t1; T["trees"] += t3 - t2
print({k: round(1e3 * v / steps, 3) for k, v in T.items()}, steps)
os.environ["PLG_TIMING"] = "1"
nbh.score_parsimony(a, d, want_moves=False)
