"""Multi-GPU scoring: one process per GPU, every process holds the full pattern-compressed
alignment, the trees (or the neighbours of one base tree) are split into contiguous blocks, and
only the per-tree scalars travel -- one small all-gather (BASELINE.json north_star (4)).

    torchrun --nproc-per-node 8 ...            # one rank per GPU
    ctx = distributed.init()                   # NCCL over NVLink on GPUs ("gloo" on CPU-only hosts: tests)
    lnl = distributed.score_trees(ctx, la, model, descs, brlens)        # BASELINE configs[3]
    moves, lnl = distributed.score_neighbourhood(ctx, la, model, desc, brlen)

There is no data-path collective: a rank reads only its block of the descriptor batch
(`plg_*_score_trees_shard`) or plans and walks only the searches that hold its block of neighbours
(`shard_index / shard_count` of `plg_*_score_neighbourhood_out`).  Scores are written by the
library straight into a device tensor (PLG_OUT_DEVICE) and that tensor is the collective's input:
no device -> host -> device bounce.  Block r of n items is [n*r/c, n*(r+1)/c) (include/pylogeny_b200.h).

PyTorch is plumbing here (device buffers, the process group); all scoring is the C-ABI library.
`scorer=` lets a caller (the gloo CPU tests) substitute the per-block scoring function; the default
is the engine and raises without a CUDA device.
"""
import ctypes as C

import numpy as np

from . import _lib


class Context(object):
    """Rank, world size and device of this process, and the torch.distributed group scores travel over."""

    def __init__(self, rank=0, world=1, device=None, group=None):
        self.rank, self.world, self.device, self.group = rank, world, device, group

    @property
    def on_gpu(self):
        return self.device is not None and self.device.type == "cuda"


def shard_range(n, rank, world):
    """Block `rank` of `world` of n items: the rule every sharded C-ABI entry point uses."""
    if world < 1 or not 0 <= rank < world:
        raise ValueError("bad shard %d of %d" % (rank, world))
    return n * rank // world, n * (rank + 1) // world


def init(backend=None, local_rank=None):
    """Join the process group torchrun set up (RANK / WORLD_SIZE / LOCAL_RANK / MASTER_*), select this
    rank's GPU in torch AND in the engine, and return the Context.  Single process: no group."""
    import os
    import torch
    import torch.distributed as dist
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0")) if local_rank is None else local_rank
    cuda = torch.cuda.is_available()
    device = torch.device("cuda", local) if cuda else torch.device("cpu")
    if cuda:
        torch.cuda.set_device(local)
        _lib.check(_lib.lib().plg_set_device(local))
    group = None
    if world > 1:
        if not dist.is_initialized():
            kw = {"device_id": device} if cuda else {}
            dist.init_process_group(backend or ("nccl" if cuda else "gloo"), rank=rank, world_size=world, **kw)
        group = dist.group.WORLD
    return Context(rank, world, device, group)


def all_gather_blocks(ctx, full, n):
    """`full` is a length-n tensor whose block shard_range(n, rank, world) this rank has filled; after
    the call every rank holds every block.  One all-gather of equal-sized (padded) blocks."""
    if ctx.world == 1:
        return full
    import torch
    import torch.distributed as dist
    lo, hi = shard_range(n, ctx.rank, ctx.world)
    if n % ctx.world == 0:  # equal blocks: gather in place
        dist.all_gather_into_tensor(full, full[lo:hi].clone(), group=ctx.group)
        return full
    width = -(-n // ctx.world)  # ceil: no block is larger
    mine = torch.zeros(width, dtype=full.dtype, device=full.device)
    mine[:hi - lo] = full[lo:hi]
    gathered = torch.empty(ctx.world * width, dtype=full.dtype, device=full.device)
    dist.all_gather_into_tensor(gathered, mine, group=ctx.group)
    for r in range(ctx.world):
        a, b = shard_range(n, r, ctx.world)
        full[a:b] = gathered[r * width:r * width + (b - a)]
    return full


def _out_tensor(ctx, n, dtype):
    import torch
    return torch.zeros(max(n, 1), dtype=dtype, device=ctx.device)


def _stream_ptr(ctx):
    import torch
    return C.c_void_p(torch.cuda.current_stream(ctx.device).cuda_stream) if ctx.on_gpu else None


def score_trees(ctx, aln, model, descs, brlens=None, tip_rows=None, scorer=None):
    """Scores of a descriptor batch sharded over the ranks: GTR/WAG lnL when `model` is given
    (aln: likelihood.LikAlignment, brlens required), Fitch lengths when model is None
    (aln: fitch.FitchAlignment).  Returns a torch tensor [n_trees] on ctx.device, complete on every rank.
    scorer(lo, hi) -> array of the block's scores replaces the engine call (tests)."""
    import torch
    descs = np.ascontiguousarray(descs, dtype=np.int32)
    if descs.ndim != 3 or descs.shape[2] != 2:
        raise ValueError("descs must be int32 [n_trees, n_tips-1, 2]")
    n, ni = descs.shape[0], descs.shape[1]
    lik = model is not None
    out = _out_tensor(ctx, n, torch.float64 if lik else torch.int32)
    lo, hi = shard_range(n, ctx.rank, ctx.world)
    if scorer is not None:
        if hi > lo:
            out[lo:hi] = torch.as_tensor(np.asarray(scorer(lo, hi)), dtype=out.dtype).to(out.device)
    elif n:
        if not ctx.on_gpu:
            raise _lib.EngineError(_lib.ERR_CUDA, "no CUDA device: pylogeny_b200 has no CPU fallback")
        rows = np.ascontiguousarray(tip_rows, dtype=np.int32) if tip_rows is not None else None
        rp = rows.ctypes.data if rows is not None else None
        torch.cuda.current_stream(ctx.device).synchronize()  # `out` was zeroed on torch's stream
        if lik:
            b = np.ascontiguousarray(brlens, dtype=np.float64)
            if b.shape != descs.shape:
                raise ValueError("brlens must have the shape of descs")
            _lib.check(_lib.lib().plg_lik_score_trees_shard(aln._h, model._h, n, ni + 1, descs.ctypes.data, b.ctypes.data, rp,
                                                            ctx.rank, ctx.world, C.c_void_p(out.data_ptr()),
                                                            _lib.OUT_DEVICE, _stream_ptr(ctx)))
        else:
            _lib.check(_lib.lib().plg_fitch_score_trees_shard(aln._h, n, ni + 1, descs.ctypes.data, rp, ctx.rank, ctx.world,
                                                              C.c_void_p(out.data_ptr()), _lib.OUT_DEVICE,
                                                              _stream_ptr(ctx)))
    return all_gather_blocks(ctx, out, n)[:n]


def score_neighbourhood(ctx, aln, model, desc, brlen=None, tip_rows=None, move_type=_lib.MOVE_SPR, want_moves=False,
                        scorer=None):
    """Scores of every distinct neighbour of one base tree, the neighbour list split across the ranks
    (each plans and walks only the searches that hold its block), combined by one all-gather.
    model None: Fitch lengths.  Returns (moves int32 [M, 4] numpy or None, scores tensor [M] on ctx.device)."""
    import torch
    from . import neighbourhood as nbh
    d = np.ascontiguousarray(desc, dtype=np.int32)
    M = nbh.neighbourhood_size(d, move_type)
    lik = model is not None
    out = _out_tensor(ctx, M, torch.float64 if lik else torch.int32)
    moves = np.empty((M, 4), dtype=np.int32) if want_moves else None
    lo, hi = shard_range(M, ctx.rank, ctx.world)
    if scorer is not None:
        if hi > lo:
            out[lo:hi] = torch.as_tensor(np.asarray(scorer(lo, hi)), dtype=out.dtype).to(out.device)
    elif M:
        if not ctx.on_gpu:
            raise _lib.EngineError(_lib.ERR_CUDA, "no CUDA device: pylogeny_b200 has no CPU fallback")
        rows = np.ascontiguousarray(tip_rows, dtype=np.int32) if tip_rows is not None else None
        rp = rows.ctypes.data if rows is not None else None
        n = C.c_int64()
        mp = moves.ctypes.data if want_moves else None
        torch.cuda.current_stream(ctx.device).synchronize()
        if lik:
            b = np.ascontiguousarray(brlen, dtype=np.float64)
            _lib.check(_lib.lib().plg_lik_score_neighbourhood_out(
                aln._h, model._h, move_type, d.shape[0] + 1, d.ctypes.data, b.ctypes.data, rp, ctx.rank, ctx.world,
                C.byref(n), mp, C.c_void_p(out.data_ptr()), _lib.OUT_DEVICE, _stream_ptr(ctx)))
        else:
            _lib.check(_lib.lib().plg_fitch_score_neighbourhood_out(
                aln._h, move_type, d.shape[0] + 1, d.ctypes.data, rp, ctx.rank, ctx.world, C.byref(n), mp,
                C.c_void_p(out.data_ptr()), _lib.OUT_DEVICE, _stream_ptr(ctx)))
    return moves, all_gather_blocks(ctx, out, M)[:M]


def argbest(scores, maximise):
    """Index and value of the best score (first one among ties, like a stable sort's [0])."""
    import torch
    idx = int(torch.argmax(scores) if maximise else torch.argmin(scores))
    return idx, scores[idx].item()
