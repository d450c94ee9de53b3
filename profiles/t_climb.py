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
t = time.perf_counter(); path, final = nbh.parsimony_greedy(a, d, max_steps=20); dt = time.perf_counter() - t
print("C3 climb: %d steps in %.1f ms = %.2f ms/step; path %s ..." % (len(path) - 1, dt * 1e3, dt * 1e3 / max(1, len(path) - 1), path[:4]))
