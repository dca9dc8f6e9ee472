"""Data-parallel plumbing of the NRMS train step: one process per GPU,
`torch.distributed` (NCCL over NVLink/NVSwitch on the GPU box, gloo in the CPU tests).

The path shards by impression: rank r takes impressions [r*B, (r+1)*B) of the global batch and
holds a full replica of the parameters (the 84 MB table included).  The only exchange step of
an iteration is the SUM-allreduce of the two gradient buffers — the flat dense block (662,600
floats) and the dense table gradient — whose entries already carry the 1/B_global factor of
`CrossEntropyLoss`'s mean over the GLOBAL batch (train_eval.py:181,194-195), so the reduced
buffers are exactly the single-process gradients and every rank applies the same Adam update.
The reference has no working multi-GPU path (dead `data_parallel` branch, model/__init__.py:33-38).

Nothing here touches the kernels: tensors may live on any device, which is what lets the
world_size-2 gloo tests run this file on CPU.
"""
from __future__ import annotations

import os
from typing import Dict, Optional, Sequence

import torch
import torch.distributed as dist


def world_info(group=None):
    """(rank, world_size) of the default / given process group; (0, 1) when not initialised."""
    if dist.is_available() and dist.is_initialized():
        return dist.get_rank(group), dist.get_world_size(group)
    return 0, 1


def init_from_env(backend: Optional[str] = None, device: Optional[torch.device] = None):
    """Initialise the default process group from torchrun's environment (RANK, WORLD_SIZE,
    MASTER_ADDR, MASTER_PORT).  No-op for a single process.  Returns (rank, world, local_rank)."""
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if world > 1 and not dist.is_initialized():
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        if backend is None:
            backend = "nccl" if (device is not None and device.type == "cuda") else "gloo"
        kwargs = {"device_id": device} if backend == "nccl" and device is not None else {}
        dist.init_process_group(backend, rank=rank, world_size=world, **kwargs)
    return rank, world, local_rank


def shard_range(n: int, rank: int, world: int):
    """Impressions [lo, hi) of a global batch of n owned by `rank` (even split; the reference's
    loaders use drop_last=False, so a trailing remainder goes to the lowest ranks)."""
    base, rem = divmod(n, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def shard_batch(batch: Dict[str, torch.Tensor], rank: int, world: int) -> Dict[str, torch.Tensor]:
    """This rank's slice of a collated MyDataset batch (data_handler.py:236-250): every tensor
    is sliced along the impression axis."""
    n = next(iter(batch.values())).shape[0]
    lo, hi = shard_range(n, rank, world)
    return {k: v[lo:hi] for k, v in batch.items()}


class GradientExchange:
    """The exchange step of a data-parallel iteration: SUM-allreduce of the gradient buffers.

    `buffers` are reduced in place, largest first so the big transfer starts as early as
    possible; with `async_op` the handles are returned so the caller can overlap the dense
    block's Adam with the table's allreduce."""

    def __init__(self, group=None):
        self.group = group
        self.rank, self.world = world_info(group)

    def allreduce(self, buffers: Sequence[torch.Tensor], async_op: bool = False):
        if self.world == 1:
            return []
        handles = []
        for b in sorted(buffers, key=lambda t: -t.numel()):
            h = dist.all_reduce(b, op=dist.ReduceOp.SUM, group=self.group, async_op=async_op)
            if async_op:
                handles.append(h)
        return handles

    # -- sharded form of the table exchange: reduce-scatter -> Adam on this rank's rows -> all-gather -----
    def _backend(self) -> str:
        return dist.get_backend(self.group) if self.world > 1 else "none"

    def reduce_scatter(self, full: torch.Tensor, out: torch.Tensor) -> None:
        """out [n] = this rank's slice of the SUM over ranks of full [world*n] (rank r owns
        [r*n, (r+1)*n)).  NCCL: one reduce-scatter; gloo (CPU tests) has none: all-reduce + slice."""
        if full.numel() != self.world * out.numel():
            raise ValueError("reduce_scatter: full must hold world * out.numel() elements")
        if self.world == 1:
            out.copy_(full.view(-1))
            return
        if self._backend() == "nccl":
            dist.reduce_scatter_tensor(out.view(-1), full.view(-1), op=dist.ReduceOp.SUM, group=self.group)
        else:
            tmp = full.view(-1).clone()
            dist.all_reduce(tmp, op=dist.ReduceOp.SUM, group=self.group)
            n = out.numel()
            out.view(-1).copy_(tmp[self.rank * n:(self.rank + 1) * n])

    def all_gather(self, full: torch.Tensor, shard: torch.Tensor) -> None:
        """full [world*n] <- every rank's shard [n], in rank order.  `shard` may be the rank's own slice
        of `full` (the in-place form NCCL supports)."""
        if full.numel() != self.world * shard.numel():
            raise ValueError("all_gather: full must hold world * shard.numel() elements")
        if self.world == 1:
            if full.data_ptr() != shard.data_ptr():
                full.view(-1).copy_(shard.view(-1))
            return
        if self._backend() == "nccl":
            dist.all_gather_into_tensor(full.view(-1), shard.view(-1), group=self.group)
        else:
            n = shard.numel()
            parts = [torch.empty(n, dtype=shard.dtype, device=shard.device) for _ in range(self.world)]
            dist.all_gather(parts, shard.reshape(-1).clone(), group=self.group)
            for r, part in enumerate(parts):
                full.view(-1)[r * n:(r + 1) * n].copy_(part)

    def max_over_ranks(self, value: float, device) -> float:
        """Timing helper: the slowest rank defines the step time."""
        if self.world == 1:
            return float(value)
        t = torch.tensor([value], dtype=torch.float64, device=device)
        dist.all_reduce(t, op=dist.ReduceOp.MAX, group=self.group)
        return float(t.item())

    def sum_over_ranks(self, t: torch.Tensor) -> torch.Tensor:
        """Metric sums of a sharded evaluation (4 metric sums + count: SURVEY.md §8e)."""
        if self.world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.SUM, group=self.group)
        return t
