"""Index-range sharding of an MSM over torch.distributed ranks (one process per GPU).

Each rank reduces its slice to one projective partial sum with `Context.msm_device`, the ranks exchange
the raw partials (192 bytes for G1, 384 for G2) with one all-gather -- point addition is not a
reduction operator NCCL knows -- and every rank folds them with `combine_partials_device`."""
from __future__ import annotations

PARTIAL_WORDS = {"g1": 48, "g2": 96}
AFFINE_WORDS = {"g1": 25, "g2": 49}


def shard_range(n: int, rank: int, world: int):
    """[lo, hi) of rank `rank`: the same split the single-process multi-device context uses."""
    return n * rank // world, n * (rank + 1) // world


def msm_sharded(ctx, group: str, bases, scalars_ptr: int, n_local: int, partial, gathered, out, world: int,
                always_gather: bool = False):
    """partial / gathered / out: int32 torch tensors on the rank's device (48 / 48*world / 25 words for G1).

    Stream ordering is part of this function: the context is switched to torch's CURRENT stream first
    (`Context.set_stream`, it stays there afterwards), because `all_gather_into_tensor` orders itself against that
    stream only -- the MSM that writes `partial`, the collective that reads it and the fold that reads `gathered`
    are then one in-order sequence.  A context left on its private stream would let NCCL read `partial` early.
    always_gather: run the all-gather + fold even when world == 1 (tests)."""
    import torch
    import torch.distributed as dist
    if partial.is_cuda:   # (the gloo / host-emulation tests pass CPU tensors: synchronous, nothing to order)
        ctx.set_stream(torch.cuda.current_stream(partial.device).cuda_stream)
    if world == 1 and not always_gather:
        ctx.msm_device(group, bases, scalars_ptr, n_local, out.data_ptr(), 0)
        return
    ctx.msm_device(group, bases, scalars_ptr, n_local, 0, partial.data_ptr())
    dist.all_gather_into_tensor(gathered, partial)
    ctx.combine_partials_device(group, gathered.data_ptr(), world, out.data_ptr())
