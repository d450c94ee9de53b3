"""world_size-2 gloo test of the product's multi-GPU helper (pylogeny_b200/distributed.py): the shard
rule, the padded / in-place all-gather, tree batches (BASELINE configs[3]) and one neighbourhood split
across ranks.  No GPU here, so the per-block scoring function is the CPU oracle (the helper's
`scorer=` hook); everything else -- init(), shard_range, score_trees, score_neighbourhood,
all_gather_blocks, the host-side neighbourhood enumeration -- is the product code the GPU path runs."""
import os
import socket
import sys

import numpy as np
import pytest
import torch
import torch.multiprocessing as mp

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _case(n_trees, seed=5):
    from util import random_desc
    rng = np.random.default_rng(seed)
    n, P = 7, 40
    codes = (1 << rng.integers(0, 4, (n, P))).astype(np.uint8)
    w = rng.integers(1, 4, P).astype(float)
    exch, freqs = rng.uniform(0.5, 2, 6), rng.dirichlet(np.ones(4) * 5)
    descs = np.stack([random_desc(rng, n) for _ in range(n_trees)])
    brs = rng.exponential(0.1, descs.shape)
    return codes, w, exch, freqs, descs, brs


def _worker(rank, world, port, n_trees, out_dir):
    import oracle
    from pylogeny_b200 import distributed as D, neighbourhood as nbh
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world), LOCAL_RANK=str(rank))
    ctx = D.init(backend="gloo")
    assert (ctx.rank, ctx.world) == (rank, world) and not ctx.on_gpu
    codes, w, exch, freqs, descs, brs = _case(n_trees)
    calls = []

    def lnl_block(lo, hi):
        calls.append((lo, hi))
        return [oracle.lnl(codes, w, exch, freqs, 0.8, descs[i], brs[i]) for i in range(lo, hi)]

    got = D.score_trees(ctx, None, object(), descs, brs, scorer=lnl_block)
    assert calls == ([D.shard_range(n_trees, rank, world)] if n_trees * (rank + 1) // world > n_trees * rank // world else [])
    # Fitch flavour (int32 scores, model None) with a deterministic stand-in per tree
    fit = D.score_trees(ctx, None, None, descs, scorer=lambda lo, hi: [1000 + 7 * i for i in range(lo, hi)])
    # one neighbourhood split across the ranks: every rank scores its block of the move list
    d0, b0 = descs[0], brs[0]
    M = nbh.neighbourhood_size(d0)

    def nb_block(lo, hi):
        dd, bb, _ = nbh.neighbour_trees(d0, b0, np.arange(lo, hi))
        return [oracle.lnl(codes, w, exch, freqs, 0.8, dd[i], bb[i]) for i in range(hi - lo)]

    _, nb = D.score_neighbourhood(ctx, None, object(), d0, b0, scorer=nb_block)
    best = D.argbest(nb, maximise=True)
    np.savez(os.path.join(out_dir, "r%d.npz" % rank), lnl=got.numpy(), fit=fit.numpy(), nb=nb.numpy(), best=np.array(best))
    torch.distributed.destroy_process_group()


@pytest.mark.parametrize("n_trees", [12, 11, 1])
def test_sharded_scoring_over_gloo(tmp_path, n_trees):
    import oracle
    from pylogeny_b200 import neighbourhood as nbh
    world, port = 2, _free_port()
    mp.spawn(_worker, args=(world, port, n_trees, str(tmp_path)), nprocs=world, join=True)
    codes, w, exch, freqs, descs, brs = _case(n_trees)
    want = np.array([oracle.lnl(codes, w, exch, freqs, 0.8, descs[i], brs[i]) for i in range(n_trees)])
    dd, bb, _ = nbh.neighbour_trees(descs[0], brs[0])
    want_nb = np.array([oracle.lnl(codes, w, exch, freqs, 0.8, dd[i], bb[i]) for i in range(len(dd))])
    for r in range(world):
        got = np.load(tmp_path / ("r%d.npz" % r))
        assert np.array_equal(got["lnl"], want)                     # every rank holds every block, bit for bit
        assert got["fit"].tolist() == [1000 + 7 * i for i in range(n_trees)] and got["fit"].dtype == np.int32
        assert np.array_equal(got["nb"], want_nb)
        assert int(got["best"][0]) == int(np.argmax(want_nb)) and got["best"][1] == want_nb.max()


def test_shard_ranges_partition():
    from pylogeny_b200 import distributed as D
    for M in (0, 1, 5, 90, 255530):
        for c in (1, 2, 3, 8):
            blocks = [D.shard_range(M, r, c) for r in range(c)]
            assert blocks[0][0] == 0 and blocks[-1][1] == M
            assert all(a[1] == b[0] for a, b in zip(blocks, blocks[1:])) and all(lo <= hi for lo, hi in blocks)
    with pytest.raises(ValueError):
        D.shard_range(10, 2, 2)


def test_engine_path_fails_loudly_without_a_gpu():
    """No CPU fallback: without `scorer=` the helper calls the engine, which needs a CUDA device."""
    from pylogeny_b200 import _lib, distributed as D
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    ctx = D.Context(0, 1, torch.device("cpu"))
    _, _, _, _, descs, brs = _case(3)
    with pytest.raises(_lib.EngineError):
        D.score_trees(ctx, None, object(), descs, brs)
