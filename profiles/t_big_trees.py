"""Larger trees than the BASELINE configs: 1024-taxon Fitch and 512-taxon GTR+G4 SPR neighbourhoods,
device-planned vs host-planned (equality) and timing."""
import os, sys, time, numpy as np
sys.path.insert(0, 'tests'); sys.path.insert(0, '.')
from util import random_desc, random_states
from pylogeny_b200 import fitch, likelihood, neighbourhood as nbh
rng = np.random.default_rng(5)
n, P = 1024, 10000
st = random_states(rng, n, P, 4, 0.01)
fa = fitch.FitchAlignment.from_dense(st, np.ones(P, dtype=np.int64))
d = random_desc(rng, n)
res = {}
for plan in ("device", "host"):
    os.environ["PLG_FITCH_NB_PLAN"] = plan
    nbh.score_parsimony(fa, d)
    t = time.perf_counter(); m, c = nbh.score_parsimony(fa, d); dt = time.perf_counter() - t
    t = time.perf_counter(); nbh.score_parsimony(fa, d, want_moves=False); dt0 = time.perf_counter() - t
    res[plan] = (m, c)
    print("Fitch %d taxa x %d patterns, %d neighbours [%s planner]: %.1f ms with the move table, %.1f ms without" % (n, P, len(c), plan, dt * 1e3, dt0 * 1e3))
assert res["device"][1].tolist() == res["host"][1].tolist() and np.array_equal(res["device"][0], res["host"][0])
n, P = 512, 1000
codes = (1 << rng.integers(0, 4, (n, P))).astype(np.uint8)
la = likelihood.LikAlignment(codes, np.ones(P))
mdl = likelihood.Model(np.array([1.2, 3.4, 0.8, 1.1, 4.1, 1.0]), np.array([0.28, 0.22, 0.24, 0.26]), 0.7)
d = random_desc(rng, n); br = rng.exponential(0.1, d.shape)
res = {}
for plan in ("device", "host"):
    os.environ["PLG_LIK_NB_PLAN"] = plan
    nbh.score_likelihood(la, mdl, d, br)
    t = time.perf_counter(); m, l = nbh.score_likelihood(la, mdl, d, br); dt = time.perf_counter() - t
    t = time.perf_counter(); nbh.score_likelihood(la, mdl, d, br, want_moves=False); dt0 = time.perf_counter() - t
    res[plan] = (m, l)
    print("GTR+G4 %d taxa x %d patterns, %d neighbours [%s planner]: %.1f ms with the move table, %.1f ms without" % (n, P, len(l), plan, dt * 1e3, dt0 * 1e3))
assert res["device"][1].tolist() == res["host"][1].tolist() and np.array_equal(res["device"][0], res["host"][0])
print("big trees ok")
