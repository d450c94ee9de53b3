"""Sharded / device-output entry points on one GPU (plg_*_score_trees_shard, plg_*_score_neighbourhood_out,
pylogeny_b200/distributed.py with world = 1, and the blocks of an emulated world = 3 written one after the
other into ONE device tensor): identical to the unsharded host-output calls, bit for bit."""
import ctypes as C

import numpy as np
import pytest

from util import random_desc, random_states

pytestmark = pytest.mark.gpu


def _setup(n=19, P=700, n_trees=23, seed=3):
    from pylogeny_b200 import fitch, likelihood
    rng = np.random.default_rng(seed)
    st = random_states(rng, n, P, 4, 0.02)
    w = rng.integers(1, 5, P).astype(np.int64)
    fa = fitch.FitchAlignment.from_dense(st, w)
    codes = (1 << rng.integers(0, 4, (n, P))).astype(np.uint8)
    la = likelihood.LikAlignment(codes, np.where(rng.random(P) < 0.5, 1.0, w.astype(float)))
    m = likelihood.Model(rng.uniform(0.5, 2, 6), rng.dirichlet(np.ones(4) * 5), 0.7)
    descs = np.stack([random_desc(rng, n) for _ in range(n_trees)])
    brs = rng.exponential(0.1, descs.shape)
    return fa, la, m, descs, brs


def test_tree_batches_shard_into_one_device_tensor():
    import torch
    from pylogeny_b200 import _lib, distributed as D
    fa, la, m, descs, brs = _setup()
    n, ni = descs.shape[0], descs.shape[1]
    want_l, want_f = la.score_trees(m, descs, brs), fa.score_trees(descs)
    dev = torch.device("cuda", 0)
    L = _lib.lib()
    for world in (1, 3, 23, 40):
        out_l = torch.zeros(n, dtype=torch.float64, device=dev)
        out_f = torch.zeros(n, dtype=torch.int32, device=dev)
        host_l = np.zeros(n)
        torch.cuda.synchronize()
        for r in range(world):
            _lib.check(L.plg_lik_score_trees_shard(la._h, m._h, n, ni + 1, descs.ctypes.data, brs.ctypes.data, None, r, world,
                                                   C.c_void_p(out_l.data_ptr()), _lib.OUT_DEVICE, None))
            _lib.check(L.plg_fitch_score_trees_shard(fa._h, n, ni + 1, descs.ctypes.data, None, r, world,
                                                     C.c_void_p(out_f.data_ptr()), _lib.OUT_DEVICE, None))
            _lib.check(L.plg_lik_score_trees_shard(la._h, m._h, n, ni + 1, descs.ctypes.data, brs.ctypes.data, None, r, world,
                                                   host_l.ctypes.data, _lib.OUT_HOST, None))
        np.testing.assert_array_equal(out_l.cpu().numpy(), want_l)
        np.testing.assert_array_equal(host_l, want_l)
        assert out_f.cpu().numpy().tolist() == want_f.tolist()
    with pytest.raises(_lib.EngineError):
        _lib.check(L.plg_lik_score_trees_shard(la._h, m._h, n, ni + 1, descs.ctypes.data, brs.ctypes.data, None, 3, 3,
                                               host_l.ctypes.data, _lib.OUT_HOST, None))
    # the helper with a single rank: same numbers, result on the device
    ctx = D.Context(0, 1, dev)
    np.testing.assert_array_equal(D.score_trees(ctx, la, m, descs, brs).cpu().numpy(), want_l)
    assert D.score_trees(ctx, fa, None, descs).cpu().numpy().tolist() == want_f.tolist()
    fa.close(); la.close(); m.close()


@pytest.mark.parametrize("move", ["spr", "nni", "tbr"])
def test_neighbourhood_blocks_into_one_device_tensor(move):
    import torch
    from pylogeny_b200 import _lib, distributed as D, neighbourhood as nbh
    fa, la, m, descs, brs = _setup(n=15 if move == "tbr" else 19)
    mt = {"spr": nbh.MOVE_SPR, "nni": nbh.MOVE_NNI, "tbr": nbh.MOVE_TBR}[move]
    d, b = descs[0], brs[0]
    mv, want_l = nbh.score_likelihood(la, m, d, b, move_type=mt)
    _, want_f = nbh.score_parsimony(fa, d, None, mt)
    M = len(want_l)
    dev = torch.device("cuda", 0)
    L = _lib.lib()
    for world in (1, 4):
        out_l = torch.zeros(M, dtype=torch.float64, device=dev)
        out_f = torch.zeros(M, dtype=torch.int32, device=dev)
        torch.cuda.synchronize()
        nm = C.c_int64()
        for r in range(world):
            _lib.check(L.plg_lik_score_neighbourhood_out(la._h, m._h, mt, d.shape[0] + 1, d.ctypes.data, b.ctypes.data, None, r,
                                                         world, C.byref(nm), None, C.c_void_p(out_l.data_ptr()),
                                                         _lib.OUT_DEVICE, None))
            _lib.check(L.plg_fitch_score_neighbourhood_out(fa._h, mt, d.shape[0] + 1, d.ctypes.data, None, r, world,
                                                           C.byref(nm), None, C.c_void_p(out_f.data_ptr()), _lib.OUT_DEVICE,
                                                           None))
        assert nm.value == M
        np.testing.assert_array_equal(out_l.cpu().numpy(), want_l)
        assert out_f.cpu().numpy().tolist() == want_f.tolist()
    ctx = D.Context(0, 1, dev)
    mv2, got = D.score_neighbourhood(ctx, la, m, d, b, move_type=mt, want_moves=True)
    np.testing.assert_array_equal(got.cpu().numpy(), want_l)
    np.testing.assert_array_equal(mv2, mv)
    _, gotf = D.score_neighbourhood(ctx, fa, None, d, move_type=mt)
    assert gotf.cpu().numpy().tolist() == want_f.tolist()
    assert D.argbest(got, True)[0] == int(np.argmax(want_l))
    fa.close(); la.close(); m.close()
