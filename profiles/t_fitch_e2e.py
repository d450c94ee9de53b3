"""Fitch handle creation and one SPR-neighbourhood step at the configs[2] shape: where the e2e milliseconds go."""
import os, sys, time, numpy as np
sys.path.insert(0, 'tests'); sys.path.insert(0, '.')
import torch
from util import random_desc
from pylogeny_b200 import fitch, neighbourhood as nbh
n, S = 256, 100_000
rng = np.random.default_rng(1002)
states = (rng.integers(0, 4, (S, n)) + 48).astype(np.uint8)
pin = torch.from_numpy(states).pin_memory().numpy()
w = np.ones(S, dtype=np.int64)
d = random_desc(np.random.default_rng(1002 + 7919), n)
def t(f, reps=7):
    ts = []
    for i in range(reps):
        torch.cuda.synchronize(); t0 = time.perf_counter(); f(); torch.cuda.synchronize(); ts.append((time.perf_counter() - t0) * 1e3)
    return np.median(ts)
a = fitch.FitchAlignment.from_dense(states, w)
for _ in range(3): nbh.score_parsimony(a, d)
print("resident, moves: %.2f ms; no moves: %.2f ms" % (t(lambda: nbh.score_parsimony(a, d)), t(lambda: nbh.score_parsimony(a, d, want_moves=False))))
def cc(x):
    b = fitch.FitchAlignment.from_dense(x, w); b.close()
print("create+close, pageable: %.2f ms; pinned: %.2f ms" % (t(lambda: cc(states)), t(lambda: cc(pin))))
os.environ["PLG_TIMING"] = "1"
cc(pin)
