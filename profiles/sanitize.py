"""Small end-to-end pass over every kernel family for compute-sanitizer (memcheck)."""
import sys
import numpy as np
sys.path.insert(0, 'tests'); sys.path.insert(0, '.')
from util import random_desc, random_states
from pylogeny_b200 import _lib, fitch, likelihood, neighbourhood as nbh
rng = np.random.default_rng(0)
n, P = 9, 333
st = random_states(rng, n, P, 5, 0.05)
w = rng.integers(1, 9, P).astype(np.int64)
fa = fitch.FitchAlignment.from_dense(st, w)
d = random_desc(rng, n); br = rng.exponential(0.1, d.shape)
print("fitch trees", fa.score_trees(np.stack([d, d])))
for mt in (nbh.MOVE_NNI, nbh.MOVE_SPR, nbh.MOVE_TBR):
    print("fitch nb", mt, nbh.score_parsimony(fa, d, None, mt)[1].sum())
print("kary", fa.score_newicks(["((T0,T1,T2),(T3,T4),T5,(T6,(T7,T8)));"], {"T%d" % i: i for i in range(n)}))
codes = rng.integers(1, 16, (n, P)).astype(np.uint8)
la = likelihood.LikAlignment(codes, w.astype(float))
m = likelihood.Model(rng.uniform(0.5, 2, 6), rng.dirichlet(np.ones(4) * 5), 0.8)
print("lnl trees", la.score_trees(m, d, br))
for mt in (nbh.MOVE_NNI, nbh.MOVE_SPR, nbh.MOVE_TBR):
    print("lnl nb", mt, nbh.score_likelihood(la, m, d, br, move_type=mt)[1].sum())
aa = rng.integers(0, 23, (n, P)).astype(np.uint8)
pa = likelihood.LikAlignment(aa, w.astype(float), _lib.DATA_AA)
pm = likelihood.Model(rng.uniform(0.1, 3, 190), rng.dirichlet(np.ones(20) * 5), 1.1)
print("aa trees", pa.score_trees(pm, d, br), "aa spr", nbh.score_likelihood(pa, pm, d, br)[1].sum())
m1 = likelihood.Model(rng.uniform(0.5, 2, 6), rng.dirichlet(np.ones(4) * 5), 1.0, 1)
print("r1", la.score_trees(m1, d, br))
print("sanitize pass done")
# the older planners / level programs behind their switches, and a deeper tree (stack levels > 2)
import os
n2 = 40
st2 = random_states(rng, n2, 700, 4, 0.02)
fa2 = fitch.FitchAlignment.from_dense(st2, np.ones(700, dtype=np.int64))
d2 = random_desc(rng, n2); br2 = rng.exponential(0.1, d2.shape)
codes2 = rng.integers(1, 16, (n2, 700)).astype(np.uint8)
la2 = likelihood.LikAlignment(codes2, np.ones(700))
for env, vals in (("PLG_TREE_MODE", ("walk", "levels")), ("PLG_LIK_NB_MODE", ("walkm", "ops")),
                  ("PLG_FITCH_NB_MODE", ("walk", "ops"))):
    for v in vals:
        os.environ[env] = v
        print(env, v, fa2.score_trees(np.stack([d2, d2])).sum(), la2.score_trees(m, np.stack([d2, d2]), np.stack([br2, br2])).sum(),
              nbh.score_parsimony(fa2, d2)[1].sum(), nbh.score_likelihood(la2, m, d2, br2)[1].sum())
    os.environ.pop(env)
print("sanitize pass 2 done")
# device-planned vs host-planned neighbourhoods, the global-memory fallback of the plan kernels, and
# the 20-state TBR path (hoisted reconnecting branch + row evaluations)
for env, vals in (("PLG_FITCH_NB_PLAN", ("device", "host")), ("PLG_LIK_NB_PLAN", ("device", "host")),
                  ("PLG_SPR_PLAN_GLOBAL", ("1",)), ("PLG_AA_EVAL_ROWS", ("1", "0"))):
    for v in vals:
        os.environ[env] = v
        print(env, v, nbh.score_parsimony(fa2, d2)[1].sum(), nbh.score_likelihood(la2, m, d2, br2)[1].sum(),
              nbh.score_likelihood(pa, pm, d, br, move_type=nbh.MOVE_TBR)[1].sum())
    os.environ.pop(env)
print("sanitize pass 3 done")
# round 2: 20-state TBR (fused candidate steps + pair grid; PLG_TBR20=0 = the level program), a deeper
# protein tree with long branches (scaling factors inside the fused steps), the batched smoother (stored
# table for 20 states, none for DNA), the single-launch smoother and the lockstep batch of one
for v in ("1", "0"):
    os.environ["PLG_TBR20"] = v
    print("PLG_TBR20", v, nbh.score_likelihood(pa, pm, d, br, move_type=nbh.MOVE_TBR)[1].sum())
os.environ.pop("PLG_TBR20")
n3 = 26
aa3 = rng.integers(0, 20, (n3, 150)).astype(np.uint8)
pa3 = likelihood.LikAlignment(aa3, rng.integers(0, 3, 150).astype(float), _lib.DATA_AA)
d3 = random_desc(rng, n3); br3 = rng.exponential(3.0, d3.shape)
print("aa tbr deep", nbh.score_likelihood(pa3, pm, d3, br3, move_type=nbh.MOVE_TBR)[1].sum())
from util import desc_to_newick
from pylogeny_b200 import alignment, libpllWrapper as W, pll
for letters, nt in (("ACGT", 11), ("ARNDCQEGHILKMFPSTWYV", 7)):
    seqs = ["".join(rng.choice(list(letters), 300)) for _ in range(nt)]
    ali = alignment.phylipFriendlyAlignment(names=["t%d" % i for i in range(nt)], seqs=seqs)
    pmod = pll.partitionModel(ali); pmod.createSimpleModel()
    nws = [desc_to_newick(random_desc(rng, nt), ali.getTaxa()) for _ in range(6)]
    print("batch smoother", letters[:4], W.getLogLikelihoodBatch(ali.getPhylip(), nws, pmod.getFileName())[0][:2])
    for mode in ("", "batch", "host"):
        if mode:
            os.environ["PLG_PLL_HANDLE_MODE"] = mode
        h = W.new(ali.getPhylip(), nws[0], pmod.getFileName())
        print("handle", mode or "single launch", W.getLogLikelihood(h))
        W.destroy(h)
        os.environ.pop("PLG_PLL_HANDLE_MODE", None)
    pmod.close(); ali.close()
print("sanitize pass 4 done")
