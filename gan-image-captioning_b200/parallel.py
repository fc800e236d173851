"""Data-parallel host logic.  The reference is single-GPU (``--device-ids`` is parsed and
ignored, src/args.py:213-216,276); the path shards by batch rows with ONE exchange per step: an all-reduce(sum) of the
flat G and D gradient buffers BEFORE the clip, averaged inside the clip+Adam kernel (grad_scale = 1/world), so the
clip coefficient is computed on the global mean gradient exactly as ``clip_grad_norm_`` does on one GPU
(src/training.py:198).  Losses are means over B*R logits, so with equal shards the mean of the shard means is exact.

Two transports.  PeerComm: the library's own one-kernel all-reduce over NVLink / NVSwitch peer memory
(gic_allreduce, csrc/allreduce.cu) -- the gradient buffers live in a symmetric allocation that every rank maps through
CUDA IPC, sums are formed in rank order (bit-identical replicas) and the square norm clip_grad_norm_ needs comes out of
the same pass.  torch.distributed (NCCL on the GPUs, gloo in the CPU tests) does the plumbing -- rendezvous, exchange of
the IPC handles, the [2, E] statistics of the synchronised BatchNorm -- and is the fallback transport
(GIC_ALLREDUCE=nccl, or when the peers cannot be mapped)."""
from __future__ import annotations

import ctypes as C
import os

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


class _RawCuda:
    """A device allocation owned by the C library, as an object torch.as_tensor can wrap without copying."""

    def __init__(self, ptr: int, numel: int):
        self.__cuda_array_interface__ = {"shape": (int(numel),), "typestr": "<f4", "data": (int(ptr), False), "version": 2}


class PeerComm:
    """Symmetric gradient buffer + one-kernel all-reduce over peer memory (include/gic_b200.h, gic_comm_* / gic_allreduce).

    One process per GPU of one node.  ``alloc(numel)`` hands out fp32 tensors inside the symmetric buffer (the flat G / D
    gradient buffers); ``allreduce_(t, channel, sqnorm)`` sums such a tensor over the ranks in place on the current stream
    and adds the square norm of the result to ``sqnorm`` (a device scalar) when given."""

    def __init__(self, nbytes: int, device):
        from . import _lib
        self._lib = _lib
        lib = _lib.lib()
        if not (dist.is_available() and dist.is_initialized()):
            raise RuntimeError("PeerComm needs an initialised torch.distributed process group (handle exchange)")
        self.rank, self.world = dist.get_rank(), dist.get_world_size()
        self.device = torch.device(device)
        with torch.cuda.device(self.device):
            self.handle = lib.gic_comm_create(self.rank, self.world, int(nbytes))
            if not self.handle:
                raise _lib.GicError("gic_comm_create: " + lib.gic_last_error().decode(errors="replace"))
            hb = int(lib.gic_comm_handle_bytes())
            mine = C.create_string_buffer(hb)
            _lib.check(lib.gic_comm_ipc_handle(self.handle, mine), "gic_comm_ipc_handle")
            gathered = [None] * self.world
            dist.all_gather_object(gathered, mine.raw)
            _lib.check(lib.gic_comm_open(self.handle, b"".join(gathered)), "gic_comm_open")
        dist.barrier()
        self.base = int(lib.gic_comm_buffer(self.handle))
        self.nbytes = int(lib.gic_comm_buffer_bytes(self.handle))
        self._used = 0

    def alloc(self, numel: int) -> torch.Tensor:
        nbytes = (int(numel) * 4 + 255) & ~255
        if self._used + nbytes > self.nbytes:
            raise MemoryError("PeerComm: symmetric buffer exhausted")
        t = torch.as_tensor(_RawCuda(self.base + self._used, numel), device=self.device)
        t._gic_comm = self                    # the allocation lives as long as its views
        self._used += nbytes
        t.zero_()
        return t

    def owns(self, t: torch.Tensor) -> bool:
        p = t.data_ptr()
        return self.base <= p and p + t.numel() * 4 <= self.base + self.nbytes

    def allreduce_(self, t: torch.Tensor, channel: int, sqnorm=None) -> torch.Tensor:
        lib = self._lib.lib()
        self._lib.check(lib.gic_allreduce(t.data_ptr(), t.numel(), self.handle, int(channel),
                                          None if sqnorm is None else sqnorm.data_ptr(), self._lib.stream()), "gic_allreduce")
        return t

    def error(self) -> bool:
        return bool(self._lib.lib().gic_comm_error(self.handle))

    def close(self):
        if getattr(self, "handle", None):
            self._lib.lib().gic_comm_destroy(self.handle)
            self.handle = None


def peer_transport_enabled() -> bool:
    return os.environ.get("GIC_ALLREDUCE", "peer").lower() != "nccl"
