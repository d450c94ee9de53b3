"""world_size-2 gloo test of the N>1 host logic: shard ranges tile the move list and the
all-gather layout bench.py uses reassembles the full score vector (no GPU: scores are stand-ins)."""
import os
import socket
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, M, out_dir):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    full = np.arange(M, dtype=np.float64) * 0.5 - 3.0            # what a single GPU would produce
    lo, hi = M * rank // world, M * (rank + 1) // world          # include/pylogeny_b200.h shard rule
    mine = torch.zeros(M // world + 1, dtype=torch.float64)
    mine[:hi - lo] = torch.from_numpy(full[lo:hi])
    gathered = torch.empty(world * (M // world + 1), dtype=torch.float64)
    dist.all_gather_into_tensor(gathered, mine)
    parts = [gathered[r * (M // world + 1):][:M * (r + 1) // world - M * r // world] for r in range(world)]
    merged = torch.cat(parts).numpy()
    # per-rank device time -> max over ranks, as bench.py reports it
    t = torch.tensor([10.0 + rank])
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    np.save(os.path.join(out_dir, "r%d.npy" % rank), np.concatenate([merged, t.numpy()]))
    dist.destroy_process_group()


@pytest.mark.parametrize("M", [90, 14762, 7])
def test_shards_tile_and_allgather_reassembles(tmp_path, M):
    world, port = 2, _free_port()
    mp.spawn(_worker, args=(world, port, M, str(tmp_path)), nprocs=world, join=True)
    want = np.arange(M, dtype=np.float64) * 0.5 - 3.0
    for r in range(world):
        got = np.load(tmp_path / ("r%d.npy" % r))
        assert np.array_equal(got[:-1], want)
        assert got[-1] == 11.0


def test_shard_ranges_partition():
    for M in (0, 1, 5, 90, 255530):
        for c in (1, 2, 3, 8):
            edges = [M * i // c for i in range(c + 1)]
            assert edges[0] == 0 and edges[-1] == M and all(a <= b for a, b in zip(edges, edges[1:]))
