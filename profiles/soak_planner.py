"""Soak: many random trees / sizes, device-planned vs host-planned neighbourhoods (Fitch and GTR+G4),
interleaved, repeated -- looks for rare mismatches (stream ordering, buffer reuse across calls)."""
import os, sys, numpy as np
sys.path.insert(0, 'tests'); sys.path.insert(0, '.')
from util import random_desc, random_states
from pylogeny_b200 import fitch, likelihood, neighbourhood as nbh
rng = np.random.default_rng(int(os.environ.get("SEED", 1)))
bad = 0
for it in range(int(os.environ.get("ITERS", 120))):
    n = int(rng.integers(4, 90)); P = int(rng.integers(1, 3000))
    st = random_states(rng, n, P, int(rng.integers(2, 7)), 0.03)
    w = np.where(rng.random(P) < 0.3, rng.integers(1, 1000, P), 1).astype(np.int64)
    fa = fitch.FitchAlignment.from_dense(st, w)
    codes = rng.integers(1, 16, (n, P)).astype(np.uint8)
    la = likelihood.LikAlignment(codes, w.astype(float))
    m = likelihood.Model(rng.uniform(0.5, 2, 6), rng.dirichlet(np.ones(4) * 5), float(rng.uniform(0.3, 2)))
    d = random_desc(rng, n); br = rng.exponential(0.1, d.shape)
    res = {}
    for plan in ("device", "host", "device"):
        os.environ["PLG_FITCH_NB_PLAN"] = plan; os.environ["PLG_LIK_NB_PLAN"] = plan
        want = bool(rng.integers(0, 2))
        f = nbh.score_parsimony(fa, d, want_moves=want)
        l = nbh.score_likelihood(la, m, d, br, want_moves=want)
        key = (f[1].tolist(), l[1].tolist())
        if "ref" not in res: res["ref"] = key
        elif key != res["ref"]:
            bad += 1; print("MISMATCH it", it, "n", n, "P", P, plan)
        if want:
            if "mv" not in res: res["mv"] = (f[0].tolist(), l[0].tolist())
            elif (f[0].tolist(), l[0].tolist()) != res["mv"]:
                bad += 1; print("MOVE MISMATCH it", it, n, P, plan)
    fa.close(); la.close(); m.close()
print("soak done, mismatches:", bad)
