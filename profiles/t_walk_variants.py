"""Time walk4m_kernel of several library builds (variants/*.so, profiles/build_variants.sh) on the
BASELINE configs[1] workload: one subprocess per library (PLG_LIB_PATH), kernel ms from the
library's own CUDA events, lnL against the level-program planner (PLG_LIK_NB_MODE=ops).
usage: python profiles/t_walk_variants.py variants/a.so variants/b.so ..."""
import os, subprocess, sys
CHILD = r'''
import ctypes as C, os, sys, time
sys.path.insert(0, os.getcwd()); sys.path.insert(0, os.path.join(os.getcwd(), "tests"))
import numpy as np
import bench
from pylogeny_b200 import likelihood, neighbourhood as nbh, _lib
L = _lib.lib()
n, S = int(os.environ.get("NT", "64")), int(os.environ.get("NS", "10000"))
wl = bench.WORKLOADS["lnl_c2" if n == 64 else "lnl_target"]
codes, w = bench.make_alignment(wl["seed"], n, S)
d, br = bench.base_tree(dict(wl, n_taxa=n), 0)
model = likelihood.Model(bench.EXCH, bench.FREQS, bench.ALPHA)
aln = likelihood.LikAlignment(codes, w)
reps = int(os.environ.get("REPS", "10"))
for _ in range(3):
    _, lnl = nbh.score_likelihood(aln, model, d, br, want_moves=False)
L.plg_profile_enable(1)
L.plg_profile_read(4, None, None, None, None, None)
t0 = time.perf_counter()
for _ in range(reps):
    _, lnl = nbh.score_likelihood(aln, model, d, br, want_moves=False)
dt = (time.perf_counter() - t0) / reps
ms, by, un, nl, rq = C.c_double(), C.c_double(), C.c_double(), C.c_int64(), C.c_double()
L.plg_profile_read(4, C.byref(ms), C.byref(by), C.byref(un), C.byref(nl), C.byref(rq))
L.plg_profile_enable(0)
rel = float("nan")
if n <= 64:
    os.environ["PLG_LIK_NB_MODE"] = "ops"
    _, ref = nbh.score_likelihood(aln, model, d, br, want_moves=False)
    rel = float(np.max(np.abs(lnl - ref) / np.abs(ref)))
print("%-28s walk %.3f ms  call %.3f ms  checksum %.6f  maxrel vs ops %.2e" % (
    os.path.basename(os.environ.get("PLG_LIB_PATH", "shipped")), ms.value / reps, dt * 1e3, lnl.sum(), rel), flush=True)
'''
for lib in sys.argv[1:] or [""]:
    env = dict(os.environ)
    if lib:
        env["PLG_LIB_PATH"] = os.path.abspath(lib)
    r = subprocess.run([sys.executable, "-c", CHILD], env=env, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    sys.stdout.write(r.stdout[-2000:] if r.returncode else r.stdout)
    sys.stdout.flush()
