"""Likelihood alignment handle at the configs[3] shape (128 x 1M): create + close from pageable and pinned host memory."""
import sys, time, numpy as np
sys.path.insert(0, '.')
import torch
from pylogeny_b200 import likelihood
n, S = 128, 1_000_000
rng = np.random.default_rng(4)
codes = (1 << rng.integers(0, 4, (n, S))).astype(np.uint8)
w = np.ones(S)
pc, pw = torch.from_numpy(codes).pin_memory().numpy(), torch.from_numpy(w).pin_memory().numpy()
def t(f, reps=7):
    ts = []
    for i in range(reps):
        torch.cuda.synchronize(); t0 = time.perf_counter(); f(); torch.cuda.synchronize(); ts.append((time.perf_counter() - t0) * 1e3)
    return np.median(ts), np.max(ts)
def cc(c, ww):
    a = likelihood.LikAlignment(c, ww); a.close()
cc(codes, w)
print("create+close 128 x 1M: pageable %.2f ms (max %.2f); pinned %.2f ms (max %.2f)" % (t(lambda: cc(codes, w)) + t(lambda: cc(pc, pw))))
