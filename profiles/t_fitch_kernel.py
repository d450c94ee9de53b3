"""Fitch SPR neighbourhood at 256 x 100k: device time of the search-walk launch (library events)."""
import os, sys, time, numpy as np, ctypes as C
sys.path.insert(0, '/root/repo'); sys.path.insert(0, '/root/repo/tests')
from util import random_desc
from pylogeny_b200 import fitch, neighbourhood as nbh, _lib
L = _lib.lib()
n, S = 256, 100_000
rng = np.random.default_rng(1002)
states = (rng.integers(0, 4, (S, n)) + 48).astype(np.uint8)
a = fitch.FitchAlignment.from_dense(states, np.ones(S, dtype=np.int64))
d = random_desc(np.random.default_rng(1002 + 7919), n)
for _ in range(3): nbh.score_parsimony(a, d, want_moves=False)
ts = []
for _ in range(10):
    t = time.perf_counter(); m, c = nbh.score_parsimony(a, d, want_moves=False); ts.append((time.perf_counter() - t) * 1e3)
import torch
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
print("step ms %.3f (min %.3f) checksum %d" % (np.median(ts), min(ts), int(c.astype(np.int64).sum())))
