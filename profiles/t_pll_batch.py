"""Batched scoring.getLogLikelihood semantics (default z, 3 smoothing passes, lnL) for the first T SPR
neighbours of a random tree, against the per-tree handle path."""
import os, sys, time, numpy as np
sys.path.insert(0, 'tests'); sys.path.insert(0, '.')
from util import desc_to_newick, random_desc, simulate_dna
from pylogeny_b200 import alignment, libpllWrapper as W, neighbourhood as nbh, pll
n, S, T = int(os.environ.get("N", 64)), int(os.environ.get("S", 10000)), int(os.environ.get("T", 1024))
rng = np.random.default_rng(3)
d = random_desc(rng, n); br = rng.exponential(0.08, d.shape) + 0.01
seqs = simulate_dna(rng, d, br, S, np.array([1.0, 3.0, 0.7, 1.3, 4.0, 1.0]), np.array([0.3, 0.2, 0.2, 0.3]))
ali = alignment.phylipFriendlyAlignment(names=["tax%d" % i for i in range(n)], seqs=seqs)
labels = ali.getTaxa()
descs, _, _ = nbh.neighbour_trees(d, None, np.arange(T, dtype=np.int64))
newicks = [desc_to_newick(x, labels) for x in descs]
pm = pll.partitionModel(ali); pm.createSimpleModel()
W.getLogLikelihoodBatch(ali.getPhylip(), newicks[:8], pm.getFileName())
t0 = time.perf_counter(); lnl, opt = W.getLogLikelihoodBatch(ali.getPhylip(), newicks, pm.getFileName()); dt = time.perf_counter() - t0
print("batch: %d trees of %d x %d in %.1f ms = %.3f ms/tree; best %.4f" % (T, n, S, dt * 1e3, dt * 1e3 / T, max(lnl)))
t0 = time.perf_counter()
ref = []
for nw in newicks[:16]:
    h = W.new(ali.getPhylip(), nw, pm.getFileName()); ref.append(W.getLogLikelihood(h)); W.destroy(h)
dt1 = (time.perf_counter() - t0) / 16
print("per tree: %.2f ms/tree; max rel diff on 16 trees %.2e; speed-up %.1fx" % (dt1 * 1e3, np.max(np.abs((np.array(ref) - lnl[:16]) / np.array(ref))), dt1 / (dt / T)))
