"""Fitch SPR neighbourhood step at the BASELINE configs[2] shape: device-planned vs host-planned."""
import os, sys, time, numpy as np
sys.path.insert(0, '/root/repo'); sys.path.insert(0, '/root/repo/tests')
from util import random_desc
from pylogeny_b200 import fitch, neighbourhood as nbh
n, S = 256, 100_000
rng = np.random.default_rng(1002)
states = (rng.integers(0, 4, (S, n)) + 48).astype(np.uint8)
a = fitch.FitchAlignment.from_dense(states, np.ones(S, dtype=np.int64))
d = random_desc(np.random.default_rng(1002 + 7919), n)
ref = None
for plan in ("host", "device"):
    os.environ["PLG_FITCH_NB_PLAN"] = plan
    for want in (True, False):
        for _ in range(3):
            m, c = nbh.score_parsimony(a, d, want_moves=want)
        t = time.perf_counter()
        for _ in range(10):
            m, c = nbh.score_parsimony(a, d, want_moves=want)
        dt = (time.perf_counter() - t) / 10
        print("plan=%s moves=%s: %.2f ms/step, checksum %d" % (plan, want, dt * 1e3, int(c.astype(np.int64).sum())))
        if ref is None:
            ref = (m, c)
        assert c.tolist() == ref[1].tolist()
        if want:
            assert m.tolist() == ref[0].tolist()
os.environ["PLG_TIMING"] = "1"
nbh.score_parsimony(a, d, want_moves=False)
