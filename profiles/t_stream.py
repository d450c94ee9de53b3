"""Full re-traversal (streaming) of descriptor batches: the working set is far beyond L2, so this is
where the kernels meet the HBM roofline.  Prints per-kernel algorithmic GB/s from the library's events."""
import sys, time, numpy as np, ctypes as C
sys.path.insert(0, 'tests'); sys.path.insert(0, '.')
from util import random_desc
from pylogeny_b200 import fitch, likelihood, _lib
L = _lib.lib()

def prof(kind):
    ms, by, un, nl = C.c_double(), C.c_double(), C.c_double(), C.c_int64()
    L.plg_profile_read(kind, C.byref(ms), C.byref(by), C.byref(un), C.byref(nl), None)
    return ms.value, by.value, nl.value

rng = np.random.default_rng(4)
# likelihood: C4 shape, 128 taxa x 1M sites, 3 random trees
n, S = 128, 1_000_000
codes = (1 << rng.integers(0, 4, (n, S))).astype(np.uint8)
la = likelihood.LikAlignment(codes, np.ones(S))
m = likelihood.Model(np.array([1.0, 2.0, 0.5, 1.5, 3.0, 1.0]), np.array([0.25, 0.3, 0.2, 0.25]), 0.8)
descs = np.stack([random_desc(rng, n) for _ in range(3)]); brs = rng.exponential(0.1, descs.shape)
la.score_trees(m, descs, brs)
L.plg_profile_enable(1); prof(1); prof(2)
t = time.perf_counter(); out = la.score_trees(m, descs, brs); dt = time.perf_counter() - t
ms, by, nl = prof(1); ems, eby, enl = prof(2)
print("lnL 128x1M x3 trees: call %.1f ms (%.3g tree-sites/s); clv_update %.1f ms %.0f GB/s (%d launches); eval %.2f ms %.0f GB/s" % (
    dt * 1e3, 3 * S / dt, ms, by / ms / 1e6, nl, ems, eby / max(ems, 1e-9) / 1e6))
la.close()
# parsimony: C3 shape, 256 taxa x 100k sites, 2048 random trees
n, S, T = 256, 100_000, 2048
states = (rng.integers(0, 4, (S, n)) + 48).astype(np.uint8)
fa = fitch.FitchAlignment.from_dense(states, np.ones(S, dtype=np.int64))
descs = np.stack([random_desc(rng, n) for _ in range(T)])
fa.score_trees(descs[:256])
prof(0)
t = time.perf_counter(); out = fa.score_trees(descs); dt = time.perf_counter() - t
ms, by, nl = prof(0)
print("Fitch 256x100k x%d trees: call %.1f ms (%.3g tree-sites/s); fitch_level %.1f ms %.0f GB/s (%d launches)" % (
    T, dt * 1e3, T * S / dt, ms, by / ms / 1e6, nl))
