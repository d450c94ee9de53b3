"""Descriptor batches (plg_*_score_trees) at BASELINE config-3/4 shapes: tree walks (default) against
the level-synchronous programs (PLG_TREE_MODE=levels), wall time per call and device time of the
scoring kernels from the library's events.  Usage: python profiles/t_stream.py [walk] [levels]"""
import os, sys, time, numpy as np, ctypes as C
sys.path.insert(0, 'tests'); sys.path.insert(0, '.')
from util import random_desc
from pylogeny_b200 import fitch, likelihood, _lib
L = _lib.lib()
modes = sys.argv[1:] or ["walk", "levels"]

def prof(kind):
    ms, by, un, nl = C.c_double(), C.c_double(), C.c_double(), C.c_int64()
    L.plg_profile_read(kind, C.byref(ms), C.byref(by), C.byref(un), C.byref(nl), None)
    return ms.value, by.value, nl.value

rng = np.random.default_rng(4)
n, S = 128, 1_000_000
codes = (1 << rng.integers(0, 4, (n, S))).astype(np.uint8)
la = likelihood.LikAlignment(codes, np.ones(S))
m = likelihood.Model(np.array([1.0, 2.0, 0.5, 1.5, 3.0, 1.0]), np.array([0.25, 0.3, 0.2, 0.25]), 0.8)
T = int(os.environ.get("T_LNL", "8"))
descs = np.stack([random_desc(rng, n) for _ in range(T)]); brs = rng.exponential(0.1, descs.shape)
ref = None
for mode in modes:
    os.environ["PLG_TREE_MODE"] = mode
    Tm = T if mode == "walk" else min(T, 3)
    la.score_trees(m, descs[:Tm], brs[:Tm])
    L.plg_profile_enable(1); prof(1); prof(2); prof(4)
    t = time.perf_counter(); out = la.score_trees(m, descs[:Tm], brs[:Tm]); dt = time.perf_counter() - t
    ms, by, nl = prof(1); ems, eby, enl = prof(2); wms, wby, wnl = prof(4); ms += wms; by += wby; nl += wnl
    L.plg_profile_enable(0)
    ref = out if ref is None else ref
    print("lnL 128x1M x%d trees [%s]: call %.1f ms = %.2f ms/tree (%.3g tree-sites/s); kernels %.1f ms, %.0f GB/s algorithmic (%d launches); maxrel %.1e" % (
        Tm, mode, dt * 1e3, dt * 1e3 / Tm, Tm * S / dt, ms + ems, (by + eby) / (ms + ems) / 1e6, nl + enl,
        np.max(np.abs(out - ref[:Tm]) / np.abs(ref[:Tm]))), flush=True)
la.close()
n, S, T = 256, 100_000, 2048
states = (rng.integers(0, 4, (S, n)) + 48).astype(np.uint8)
fa = fitch.FitchAlignment.from_dense(states, np.ones(S, dtype=np.int64))
descs = np.stack([random_desc(rng, n) for _ in range(T)])
ref = None
for mode in modes:
    os.environ["PLG_TREE_MODE"] = mode
    fa.score_trees(descs[:256])
    L.plg_profile_enable(1); prof(0)
    t = time.perf_counter(); out = fa.score_trees(descs); dt = time.perf_counter() - t
    ms, by, nl = prof(0)
    L.plg_profile_enable(0)
    ref = out if ref is None else ref
    print("Fitch 256x100k x%d trees [%s]: call %.1f ms (%.3g tree-sites/s); kernels %.2f ms, %.0f GB/s algorithmic (%d launches); equal %s" % (
        T, mode, dt * 1e3, T * S / dt, ms, by / ms / 1e6, nl, bool((out == ref).all())), flush=True)
