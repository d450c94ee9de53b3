import os, sys, time, numpy as np
sys.path.insert(0, 'tests'); sys.path.insert(0, '.')
from util import random_desc
from pylogeny_b200 import likelihood, neighbourhood as nbh, _lib
import ctypes as C
n, S = 64, 10000
rng = np.random.default_rng(1001)
codes = (1 << rng.integers(0, 4, (n, S))).astype(np.uint8)
la = likelihood.LikAlignment(codes, np.ones(S))
m = likelihood.Model(np.array([1.2, 3.4, 0.8, 1.1, 4.1, 1.0]), np.array([0.28, 0.22, 0.24, 0.26]), 0.7)
d = random_desc(np.random.default_rng(6), n); br = np.random.default_rng(7).exponential(0.1, d.shape)
L = _lib.lib()
for i in range(3):
    nbh.score_likelihood(la, m, d, br)
L.plg_profile_enable(1); L.plg_profile_read(1, None, None, None, None, None); L.plg_profile_read(4, None, None, None, None, None)
ts = []
REPS = 10
for i in range(REPS):
    t = time.perf_counter(); mv, l = nbh.score_likelihood(la, m, d, br); ts.append((time.perf_counter() - t) * 1e3)
ms, by, un, nl = C.c_double(), C.c_double(), C.c_double(), C.c_int64()
L.plg_profile_read(1, C.byref(ms), C.byref(by), C.byref(un), C.byref(nl), None)
wms = C.c_double()
L.plg_profile_read(4, C.byref(wms), None, None, None, None)
print("call ms %.2f (min %.2f)  walk kernel ms %.3f  base-tree kernels ms/step %.2f  checksum %.6f" % (np.median(ts), min(ts), wms.value / REPS, ms.value / REPS, l.sum()))
