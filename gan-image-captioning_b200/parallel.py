"""Data-parallel host logic.  The reference is single-GPU (``--device-ids`` is parsed and
ignored, src/args.py:213-216,276); the path shards by batch rows with ONE exchange per step: an all-reduce(sum) of the
flat G and D gradient buffers BEFORE the clip, averaged inside the clip+Adam kernel (grad_scale = 1/world), so the
clip coefficient is computed on the global mean gradient exactly as ``clip_grad_norm_`` does on one GPU
(src/training.py:198).  Losses are means over B*R logits, so with equal shards the mean of the shard means is exact.

Backend-agnostic: NCCL over NVLink on the GPUs, gloo in the CPU tests."""
from __future__ import annotations

import torch
import torch.distributed as dist


def world_size() -> int:
    return dist.get_world_size() if dist.is_available() and dist.is_initialized() else 1


def rank() -> int:
    return dist.get_rank() if dist.is_available() and dist.is_initialized() else 0


def shard_rows(global_rows: int, rank_: int, world: int) -> slice:
    """Rows [rank*B/world, (rank+1)*B/world) of the global batch; equal shards are required for exact means."""
    if global_rows % world:
        raise ValueError("global batch %d is not divisible by world size %d" % (global_rows, world))
    per = global_rows // world
    return slice(rank_ * per, (rank_ + 1) * per)


def shard_uniforms(u: torch.Tensor, rank_: int, world: int, batch_dim: int = 1) -> torch.Tensor:
    """Slice caller-supplied draws (u[L,B,V], u[L,B], keep[k,B*R,F] with batch_dim rows grouped per caption) by batch row
    so that results do not depend on the number of GPUs."""
    n = u.shape[batch_dim]
    if n % world:
        raise ValueError("dimension %d of size %d is not divisible by world size %d" % (batch_dim, n, world))
    per = n // world
    return u.narrow(batch_dim, rank_ * per, per)


def allreduce_sum_(flat: torch.Tensor) -> torch.Tensor:
    """In-place sum over ranks of a flat gradient buffer (no-op for world 1)."""
    if world_size() > 1:
        dist.all_reduce(flat, op=dist.ReduceOp.SUM)
    return flat


def grad_scale() -> float:
    """Factor applied to the summed gradients inside the optimizer kernel."""
    return 1.0 / world_size()


def clip_coef(sqnorm_sum: float, max_norm: float, scale: float) -> float:
    """clip_grad_norm_ coefficient on the averaged gradient: min(1, max_norm / (||g_sum|| * scale + 1e-6)) * scale."""
    nrm = (sqnorm_sum ** 0.5) * scale
    return min(1.0, max_norm / (nrm + 1e-6)) * scale
