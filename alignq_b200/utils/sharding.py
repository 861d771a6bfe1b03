"""Data-parallel sharding of the hot path over the GPUs of one box (one process per GPU,
``torch.distributed``; NCCL over NVLink on the B200 box, gloo in the CPU tests).

What shards how:
  * activation quantizer  -- batch-sharded, fixed N(0,1) statistics (QA:97): NO collective;
  * weight quantizer      -- weights are replicated, per-tensor statistics of identical data: NO
                             collective (``combine_moments`` is the (sum, sum-of-squares) hook for the
                             case that a weight tensor is itself sharded);
  * gradients             -- ONE all-reduce of the flat gradient buffer per step (``allreduce_mean_``);
  * ADMM Gram             -- G is [B, B] ACROSS samples, so it does not shard over the batch.  Default:
                             replicas only (each rank its own [b, b] Gram and ADMM(b), exactly the
                             reference at train_batch_size = b; config 5 is quoted that way).  For
                             global-batch parity the features are sharded instead: each rank computes
                             the partial Gram of its feature slice over ALL rows and the partials are
                             summed with one all-reduce (``allreduce_gram_partial_``).
The reference has no distributed code at all (``--gpus`` only ever uses ``gpus[0]``, QA:12).
"""
from __future__ import annotations

import torch
import torch.distributed as dist


def shard_batch(x: torch.Tensor, rank: int, world: int) -> torch.Tensor:
    """Contiguous slice of the global batch owned by ``rank`` (ragged tails go to the first ranks)."""
    n = x.shape[0]
    base, extra = divmod(n, world)
    start = rank * base + min(rank, extra)
    return x[start: start + base + (1 if rank < extra else 0)]


def feature_slice(F: int, rank: int, world: int, align: int = 32):
    """[begin, end) of the feature columns owned by ``rank``; boundaries aligned to the kernels' 32-column tiles."""
    tiles = (F + align - 1) // align
    base, extra = divmod(tiles, world)
    t0 = rank * base + min(rank, extra)
    t1 = t0 + base + (1 if rank < extra else 0)
    return min(t0 * align, F), min(t1 * align, F)


def allreduce_mean_(flat: torch.Tensor, group=None, world: int | None = None) -> torch.Tensor:
    """In-place mean over ranks of a flat buffer (the gradient exchange: one collective per step)."""
    world = dist.get_world_size(group) if world is None else world
    if world > 1:
        dist.all_reduce(flat, op=dist.ReduceOp.SUM, group=group)
        flat.mul_(1.0 / world)
    return flat


def allreduce_sum_(flat: torch.Tensor, group=None) -> torch.Tensor:
    """In-place sum over ranks (the caller folds 1/world into its next kernel)."""
    if dist.is_initialized() and dist.get_world_size(group) > 1:
        dist.all_reduce(flat, op=dist.ReduceOp.SUM, group=group)
    return flat


def combine_moments(s: torch.Tensor, ss: torch.Tensor, n: torch.Tensor, group=None):
    """Global mean and unbiased std from per-rank (sum, sum of squares, count): the (sum, sumsq)
    all-reduce that makes sharded statistics equal the single-device ``torch.mean`` / ``torch.std``."""
    buf = torch.stack([s.double().reshape(()), ss.double().reshape(()), n.double().reshape(())])
    if dist.is_initialized() and dist.get_world_size(group) > 1:
        dist.all_reduce(buf, op=dist.ReduceOp.SUM, group=group)
    S, SS, N = buf[0], buf[1], buf[2]
    mean = S / N
    var = (SS - S * mean) / (N - 1)
    return mean.float(), var.clamp_min(0).sqrt().float()


def allreduce_gram_partial_(G_local: torch.Tensor, F_local: int, F_total: int, group=None) -> torch.Tensor:
    """Feature-sharded Gram: ``G_local`` = corr of this rank's feature slice (already divided by
    ``F_local``); returns the full-feature Gram on every rank (in place)."""
    G_local.mul_(float(F_local) / float(F_total))
    if dist.is_initialized() and dist.get_world_size(group) > 1:
        dist.all_reduce(G_local, op=dist.ReduceOp.SUM, group=group)
    return G_local
