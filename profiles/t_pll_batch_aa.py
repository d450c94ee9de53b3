"""Batched getLogLikelihood semantics for protein data (WAG + G4): T SPR neighbours of a random 50-taxon tree,
5k sites, against the per-tree handle path."""
import os, sys, time, numpy as np
sys.path.insert(0, 'tests'); sys.path.insert(0, '.')
from util import desc_to_newick, random_desc
from pylogeny_b200 import alignment, libpllWrapper as W, neighbourhood as nbh, pll
n, S, T = int(os.environ.get("N", 50)), int(os.environ.get("S", 5000)), int(os.environ.get("T", 256))
rng = np.random.default_rng(4)
aa = "ARNDCQEGHILKMFPSTWYV"
seqs = ["".join(rng.choice(list(aa), S)) for _ in range(n)]
ali = alignment.phylipFriendlyAlignment(names=["tax%d" % i for i in range(n)], seqs=seqs)
labels = ali.getTaxa()
d = random_desc(rng, n)
descs, _, _ = nbh.neighbour_trees(d, None, np.arange(T, dtype=np.int64))
newicks = [desc_to_newick(x, labels) for x in descs]
pm = pll.partitionModel(ali); pm.createSimpleModel()
W.getLogLikelihoodBatch(ali.getPhylip(), newicks[:8], pm.getFileName())
t0 = time.perf_counter(); lnl, opt = W.getLogLikelihoodBatch(ali.getPhylip(), newicks, pm.getFileName()); dt = time.perf_counter() - t0
print("AA batch: %d trees of %d x %d in %.1f ms = %.3f ms/tree; best %.4f" % (T, n, S, dt * 1e3, dt * 1e3 / T, max(lnl)))
t0 = time.perf_counter()
ref = []
for nw in newicks[:4]:
    h = W.new(ali.getPhylip(), nw, pm.getFileName()); ref.append(W.getLogLikelihood(h)); W.destroy(h)
dt1 = (time.perf_counter() - t0) / 4
print("AA per tree: %.2f ms/tree; max rel diff %.2e" % (dt1 * 1e3, np.max(np.abs((np.array(ref) - lnl[:4]) / np.array(ref)))))
