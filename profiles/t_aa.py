import os, sys, time, numpy as np
sys.path.insert(0, 'tests'); sys.path.insert(0, '.')
from util import random_desc
from pylogeny_b200 import likelihood, neighbourhood as nbh, _lib
import ctypes as C
rng = np.random.default_rng(1005)
n, S = 50, 5000
codes = rng.integers(0, 20, (n, S)).astype(np.uint8)
exch, f = rng.uniform(0.05, 5.0, 190), rng.dirichlet(np.ones(20) * 10)
a = likelihood.LikAlignment(codes, np.ones(S), _lib.DATA_AA)
m = likelihood.Model(exch, f, 0.9)
d = random_desc(rng, n); br = rng.exponential(0.1, d.shape)
L = _lib.lib()
for mt, name in ((nbh.MOVE_SPR, "SPR"), (nbh.MOVE_TBR, "TBR")):
    nbh.score_likelihood(a, m, d, br, move_type=mt)
    L.plg_profile_enable(1)
    for k in range(5): L.plg_profile_read(k, None, None, None, None, None)
    t = time.perf_counter(); mv, l = nbh.score_likelihood(a, m, d, br, move_type=mt); dt = (time.perf_counter() - t) * 1e3
    out = []
    for k, kn in enumerate(("fitch", "clv_update", "root_eval", "pmatrix")):
        ms, by, un, nl = C.c_double(), C.c_double(), C.c_double(), C.c_int64()
        L.plg_profile_read(k, C.byref(ms), C.byref(by), C.byref(un), C.byref(nl), None)
        if nl.value: out.append("%s %.2f ms %.0f GB/s (%d launches)" % (kn, ms.value, by.value / max(ms.value, 1e-9) / 1e6, nl.value))
    L.plg_profile_enable(0)
    print("AA 50x5k %s: M=%d call %.1f ms -> %.3g tree-sites/s | %s" % (name, len(l), dt, len(l) * S / dt * 1e3, "; ".join(out)))
