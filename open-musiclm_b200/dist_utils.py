"""Host-side pieces of the data-parallel path (device-agnostic so that they can be exercised under gloo on CPU).

The hot path shards by batch: every rank runs the same step on its own sequences, gradients are summed over
ranks with ONE all-reduce of the flat fp32 arena per optimiser step, and the DDP-mean 1/world factor is folded into
the clip/AdamW kernel (trainer.py:154-155,439 of the reference -> HotPathTrainer.train_step).
"""
import torch
import torch.distributed as dist


def world_info(group=None):
    if dist.is_available() and dist.is_initialized():
        return dist.get_world_size(group), dist.get_rank(group)
    return 1, 0


def allreduce_sum_(flat: torch.Tensor, group=None) -> torch.Tensor:
    """In-place sum of the flat gradient arena over all ranks (no-op for a single process)."""
    world, _ = world_info(group)
    if world > 1:
        dist.all_reduce(flat, op=dist.ReduceOp.SUM, group=group)
    return flat


def grad_prescale(group=None) -> float:
    """Factor that turns the all-reduced SUM into DistributedDataParallel's mean."""
    world, _ = world_info(group)
    return 1.0 / world


def rank_seed(seed: int, rank: int) -> int:
    """Per-rank seed of the dropout / forgetful-mask streams (weights use the SAME seed on every rank)."""
    return seed * 1000003 + rank * 7919 + 1
