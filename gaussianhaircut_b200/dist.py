"""Multi-GPU plumbing for the rasterizer path: views shard one per GPU, ONE all-reduce per step.

The reference is single-process / single-GPU (src/utils/general_utils.py:143 pins cuda:0; no
collective call sites anywhere), so this is new functionality whose parity
gate is  allreduce(grads) == sum over ranks of the single-GPU gradient of that rank's view.

A training step consumes one camera (src/train_gaussians.py:103-110) and cameras are independent
given the same Gaussians, so the path shards by view with no collective inside the op.  The only
exchange is the sum of the per-Gaussian gradients, which the backward already writes into one flat
float32 arena (`_C.rasterize_gaussians_backward_arena`): the arena IS the NCCL buffer, no packing
copy.  torch.distributed (NCCL over NVLink/NVSwitch on the B200 box, gloo in CPU tests) is plumbing.
"""
from __future__ import annotations

from typing import List, Optional

import torch
import torch.distributed as dist


def views_for_step(step: int, rank: int, world_size: int, num_cameras: int) -> int:
    """Camera index rendered by `rank` at optimiser step `step`: consecutive blocks of `world_size`
    cameras per step, wrapping around the camera list (one view per GPU per step)."""
    if not (0 <= rank < world_size):
        raise ValueError("rank out of range")
    return (step * world_size + rank) % num_cameras


def epoch_schedule(world_size: int, num_cameras: int) -> List[List[int]]:
    """All steps of one pass over the cameras: ceil(num_cameras / world_size) steps, each a list of
    `world_size` camera ids (the last step wraps)."""
    steps = (num_cameras + world_size - 1) // world_size
    return [[views_for_step(s, r, world_size, num_cameras) for r in range(world_size)] for s in range(steps)]


def trainable_slice(flat: torch.Tensor, num_gaussians: int, call_shape: str = "native") -> torch.Tensor:
    """The contiguous prefix of the gradient arena that the optimizer consumes.

    `native` (the op computes covariances from scales/rotations): rotations, colors, opacity,
    means2D, means3D, scales = the first 24 floats per Gaussian; the conic / cov3D segments behind
    it stay zero and need not travel.  Any other call shape: the whole arena.
    """
    from . import _C
    if call_shape == "native":
        return flat[: num_gaussians * _C.GRAD_FLOATS_TRAINABLE_NATIVE]
    return flat


def allreduce_gradient_arena(flat: torch.Tensor, group: Optional[dist.ProcessGroup] = None,
                             average: bool = False, async_op: bool = False):
    """Sum (or mean) the flat gradient arena (or a `trainable_slice` of it) over all ranks with a
    single collective.

    No-op when torch.distributed is not initialised or the world has one rank.  With
    `async_op=True` returns the work handle so the caller can overlap the next view's forward.
    """
    if not dist.is_available() or not dist.is_initialized():
        return None
    world = dist.get_world_size(group)
    if world == 1:
        return None
    work = dist.all_reduce(flat, op=dist.ReduceOp.SUM, group=group, async_op=async_op)
    if average:
        if async_op:
            work.wait()
        flat.div_(world)
        return None
    return work


def allreduce_densification_stats(grad_accum: torch.Tensor, denom: torch.Tensor, max_radii2D: torch.Tensor,
                                  group: Optional[dist.ProcessGroup] = None) -> None:
    """Keep the densification statistics replica-consistent (src/train_gaussians.py:161-164,
    src/scene/gaussian_model.py:739-741): SUM for the accumulated view-space gradient norm and its
    denominator, MAX for the largest screen radius."""
    if not dist.is_available() or not dist.is_initialized() or dist.get_world_size(group) == 1:
        return
    dist.all_reduce(grad_accum, op=dist.ReduceOp.SUM, group=group)
    dist.all_reduce(denom, op=dist.ReduceOp.SUM, group=group)
    dist.all_reduce(max_radii2D, op=dist.ReduceOp.MAX, group=group)
