"""A/B of the DNA x Gamma-4 neighbourhood planners on the bench workload (64 taxa x 10k sites):
wall ms per step and device ms of the CLV/search kernels (library events)."""
import ctypes as C, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import bench
from pylogeny_b200 import likelihood, neighbourhood as nbh, _lib
import torch
L = _lib.lib()
codes, w = bench.make_alignment()
d, br = bench.make_tree(bench.SEED)
model = likelihood.Model(bench.EXCH, bench.FREQS, bench.ALPHA)
aln = likelihood.LikAlignment(codes, w)
ref = None
reps = int(os.environ.get("REPS", "10"))
for mode in sys.argv[1:] or ["visits", "walk"]:
    os.environ["PLG_LIK_NB_MODE"] = mode
    for _ in range(3):
        _, lnl = nbh.score_likelihood(aln, model, d, br)
    torch.cuda.synchronize()
    L.plg_profile_enable(1)
    L.plg_profile_read(1, None, None, None, None, None)
    L.plg_profile_read(4, None, None, None, None, None)
    t0 = time.perf_counter()
    for _ in range(reps):
        _, lnl = nbh.score_likelihood(aln, model, d, br)
    torch.cuda.synchronize()
    dt = (time.perf_counter() - t0) / reps
    ms, by, un, n, rq = C.c_double(), C.c_double(), C.c_double(), C.c_int64(), C.c_double()
    L.plg_profile_read(1, C.byref(ms), C.byref(by), C.byref(un), C.byref(n), C.byref(rq))
    tot = [ms.value, by.value, n.value, rq.value]
    L.plg_profile_read(4, C.byref(ms), C.byref(by), C.byref(un), C.byref(n), C.byref(rq))
    wms = ms.value
    ms.value += tot[0]; by.value += tot[1]; n.value += tot[2]; rq.value += tot[3]
    L.plg_profile_enable(0)
    if ref is None:
        ref = lnl
    print("%-7s %.3f ms/step wall, kernels %.3f ms/step (walk %.3f) in %d launches, alg %.1f GB/s, requested %.1f GB/s, checksum %.9f maxrel %.2e" % (
        mode, dt * 1e3, ms.value / reps, wms / reps, n.value // reps, by.value / ms.value / 1e6, rq.value / ms.value / 1e6, lnl.sum(),
        np.max(np.abs(lnl - ref) / np.abs(ref))), flush=True)
