"""north_star target shape for likelihood: every SPR neighbour of a 256-taxon x 100k-site tree, GTR+G4."""
import os, sys, time, numpy as np
sys.path.insert(0, 'tests'); sys.path.insert(0, '.')
from util import random_desc
from pylogeny_b200 import likelihood, neighbourhood as nbh, _lib
import ctypes as C
n, S = int(os.environ.get("N", 256)), int(os.environ.get("S", 100000))
rng = np.random.default_rng(1002)
codes = (1 << rng.integers(0, 4, (n, S))).astype(np.uint8)
la = likelihood.LikAlignment(codes, np.ones(S))
m = likelihood.Model(np.array([1.2, 3.4, 0.8, 1.1, 4.1, 1.0]), np.array([0.28, 0.22, 0.24, 0.26]), 0.7)
d = random_desc(np.random.default_rng(6), n); br = np.random.default_rng(7).exponential(0.1, d.shape)
L = _lib.lib()
nbh.score_likelihood(la, m, d, br)
L.plg_profile_enable(1); L.plg_profile_read(4, None, None, None, None, None)
reps = 2
t = time.perf_counter()
for i in range(reps):
    mv, l = nbh.score_likelihood(la, m, d, br)
dt = (time.perf_counter() - t) / reps
ms, by, un, nl = C.c_double(), C.c_double(), C.c_double(), C.c_int64()
L.plg_profile_read(4, C.byref(ms), C.byref(by), C.byref(un), C.byref(nl), None)
print("%d x %d: %d neighbours, %.1f ms/step = %.3e tree-sites/s; walk kernel %.1f ms/step, %.0f GB/s algorithmic; checksum %.6f"
      % (n, S, len(l), dt * 1e3, len(l) * S / dt, ms.value / reps, by.value / ms.value / 1e6, l.sum()))
# spot check against full re-traversal of materialised neighbours
sel = np.linspace(0, len(l) - 1, 8).astype(np.int64)
descs, brs, _ = nbh.neighbour_trees(d, br, sel)
full = la.score_trees(m, descs, brs)
print("max rel diff vs re-traversal:", np.max(np.abs(full - l[sel]) / np.abs(full)))
