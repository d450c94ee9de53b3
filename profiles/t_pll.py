"""Per-tree cost of the libpllWrapper drop-in (new + getLogLikelihood with 3 smoothing passes + destroy) --
what scoring.getLogLikelihood pays per tree, as the reference does (scoring.py:41-75)."""
import os, sys, time, numpy as np
sys.path.insert(0, 'tests'); sys.path.insert(0, '.')
from util import desc_to_newick, random_desc, simulate_dna
from pylogeny_b200 import alignment, libpllWrapper as W, pll
n, S = int(os.environ.get("N", 64)), int(os.environ.get("S", 10000))
rng = np.random.default_rng(3)
d = random_desc(rng, n); br = rng.exponential(0.08, d.shape) + 0.01
seqs = simulate_dna(rng, d, br, S, np.array([1.0, 3.0, 0.7, 1.3, 4.0, 1.0]), np.array([0.3, 0.2, 0.2, 0.3]))
ali = alignment.phylipFriendlyAlignment(names=["tax%d" % i for i in range(n)], seqs=seqs)
nw = desc_to_newick(d, ali.getTaxa())
pm = pll.partitionModel(ali); pm.createSimpleModel()
for rep in range(3):
    t0 = time.perf_counter(); h = W.new(ali.getPhylip(), nw, pm.getFileName()); t1 = time.perf_counter()
    l0 = W.evaluate(h); t2 = time.perf_counter()
    l1 = W.getLogLikelihood(h); t3 = time.perf_counter()
    W.destroy(h); t4 = time.perf_counter()
    print("%d x %d: new %.1f ms, evaluate %.2f ms, getLogLikelihood (3 smoothing passes) %.1f ms, destroy %.2f ms; lnL %.4f -> %.4f" % (
        n, S, (t1 - t0) * 1e3, (t2 - t1) * 1e3, (t3 - t2) * 1e3, (t4 - t3) * 1e3, l0, l1))
