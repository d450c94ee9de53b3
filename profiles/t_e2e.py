import os, sys, time, numpy as np
sys.path.insert(0, 'tests'); sys.path.insert(0, '.')
import torch
from util import random_desc
from pylogeny_b200 import likelihood, neighbourhood as nbh, _lib
n, S = 64, 10000
rng = np.random.default_rng(1001)
codes = (1 << rng.integers(0, 4, (n, S))).astype(np.uint8)
w = np.ones(S)
m = likelihood.Model(np.array([1.2, 3.4, 0.8, 1.1, 4.1, 1.0]), np.array([0.28, 0.22, 0.24, 0.26]), 0.7)
d = random_desc(np.random.default_rng(6), n); br = np.random.default_rng(7).exponential(0.1, d.shape)
la = likelihood.LikAlignment(codes, w)
for i in range(3): nbh.score_likelihood(la, m, d, br)
def t(f, reps=5):
    ts = []
    for i in range(reps):
        torch.cuda.synchronize(); t0 = time.perf_counter(); f(); torch.cuda.synchronize(); ts.append((time.perf_counter() - t0) * 1e3)
    return np.median(ts)
print("persistent aln, moves: %.2f ms" % t(lambda: nbh.score_likelihood(la, m, d, br)))
print("persistent aln, no moves: %.2f ms" % t(lambda: nbh.score_likelihood(la, m, d, br, want_moves=False)))
def e2e(want):
    a2 = likelihood.LikAlignment(codes, w); nbh.score_likelihood(a2, m, d, br, want_moves=want); a2.close()
print("create+score+close, moves: %.2f ms" % t(lambda: e2e(True)))
print("create+score+close, no moves: %.2f ms" % t(lambda: e2e(False)))
def cc():
    a2 = likelihood.LikAlignment(codes, w); a2.close()
print("create+close only: %.2f ms" % t(cc))
os.environ["PLG_LIK_NB_PLAN"] = "host"
print("host plan: create+score+close: %.2f ms" % t(lambda: e2e(True)))
os.environ["PLG_LIK_NB_PLAN"] = "device"
os.environ["PLG_TIMING"] = "1"
e2e(True)
