"""Multi-GPU plumbing for the two places the path touches more than one rank.

Training: the batch shards by image under the reference's own DDP wrapper (train_util.py:174-175);
every loss kernel is rank-local, so there is nothing to add on the data path.

Evaluation: the reference validates on rank 0 only (train_util.py:354,371-390).  Here every rank
accumulates its own shard of batches in a ``MetricAccumulator`` and the int64 histograms are summed
with ONE all-reduce; integers are associative, so the result is bit-identical for any rank count.
For the reference's float mIoU the per-label first-appearance batch index (which fixes the dict
insertion order, Q11) is reduced with MIN over GLOBAL batch indices in the same call sequence, and the
class filter of the final mean (the last batch's ground truth) is taken from the rank that owns the
globally last batch (``global_last_batch_mask``).  ``validate_model(all_reduce=True)`` does all of it.
"""
from __future__ import annotations

from typing import Optional

import torch
import torch.distributed as dist

from .evaluation import MetricAccumulator


def shard_batches(n_batches: int, rank: int, world_size: int):
    """Round-robin assignment of validation batches to ranks; yields GLOBAL batch indices."""
    return range(rank, n_batches, world_size)


def all_reduce_metrics(acc: MetricAccumulator, group: Optional[dist.ProcessGroup] = None) -> MetricAccumulator:
    """Sum the integer state across ranks in place (one SUM all-reduce of 4*C+3 int64 words, one MIN
    all-reduce of C int32 first-seen indices).  Works with NCCL (CUDA tensors) and gloo (CPU tensors)."""
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
        return acc
    flat = torch.cat([acc.acc.reshape(-1), acc.counters.reshape(-1)])
    dist.all_reduce(flat, op=dist.ReduceOp.SUM, group=group)
    acc.acc.copy_(flat[: acc.acc.numel()].view_as(acc.acc))
    acc.counters.copy_(flat[acc.acc.numel():])
    dist.all_reduce(acc.first_seen, op=dist.ReduceOp.MIN, group=group)
    return acc


def reduce_state_tensors(acc_tensor: torch.Tensor, counters: torch.Tensor, first_seen: torch.Tensor,
                         group: Optional[dist.ProcessGroup] = None) -> None:
    """Same reduction on bare tensors (used by the gloo CPU test, which has no CUDA kernels)."""
    flat = torch.cat([acc_tensor.reshape(-1), counters.reshape(-1)])
    dist.all_reduce(flat, op=dist.ReduceOp.SUM, group=group)
    acc_tensor.copy_(flat[: acc_tensor.numel()].view_as(acc_tensor))
    counters.copy_(flat[acc_tensor.numel():])
    dist.all_reduce(first_seen, op=dist.ReduceOp.MIN, group=group)


def global_last_batch_mask(local_mask: torch.Tensor, last_global_index: int,
                           group: Optional[dist.ProcessGroup] = None) -> torch.Tensor:
    """The class filter of the final mIoU mean (validate.py:206-210 iterates the labels of the LAST batch's ground truth,
    quirk Q11) for sharded validation: ``local_mask`` (uint8 [C]) describes this rank's last batch, ``last_global_index``
    its global batch index (-1: the rank saw no batch).  Returns the mask of the rank that owns the globally last batch
    (one MAX all-reduce of an index, one SUM all-reduce of C bytes)."""
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
        return local_mask
    top = torch.tensor([last_global_index], device=local_mask.device, dtype=torch.int64)
    dist.all_reduce(top, op=dist.ReduceOp.MAX, group=group)
    mine = local_mask.to(torch.int32) if int(top) == last_global_index and last_global_index >= 0 else torch.zeros_like(local_mask, dtype=torch.int32)
    dist.all_reduce(mine, op=dist.ReduceOp.SUM, group=group)
    return (mine > 0).to(torch.uint8)
