"""Multi-GPU plumbing: one process per GPU (torch.distributed, NCCL over NVLink / NVSwitch).

The projection path shards by instances and needs NO collective: `shard_rows` gives each rank a
contiguous row range.  The only exchange in the training configuration (BASELINE.json configs[3]) is the
data-parallel gradient average of the small actor / critic / dual networks, done by `FlatGradBucket` as ONE
all-reduce over a pre-flattened contiguous bucket per optimiser step (latency-bound at 0.35-1.3 MB; NCCL picks
its NVLS / low-latency protocol) — the reference has no distributed code at all.
"""
from __future__ import annotations

from typing import Iterable, List, Tuple

import torch
import torch.distributed as dist


def shard_rows(n: int, rank: int, world: int) -> Tuple[int, int]:
    """Contiguous row range [lo, hi) of rank `rank` out of `world`; sizes differ by at most one row."""
    base, rem = divmod(n, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def global_loop_test(flags: torch.Tensor, group=None) -> torch.Tensor:
    """Combine the per-step loop-test flags of all shards (logical OR) so that every rank takes the step count the
    reference would take on the concatenated batch (rpo_ddpg.py:271-272).  Optional: without it each shard is its own
    batch (the default of the sharded sweep)."""
    if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
        dist.all_reduce(flags, op=dist.ReduceOp.MAX, group=group)
    return flags


class FlatGradBucket(object):
    """Keeps the gradients of `params` as views into one contiguous buffer and averages it with one all-reduce."""

    def __init__(self, params: Iterable[torch.nn.Parameter]):
        self.params: List[torch.nn.Parameter] = [p for p in params if p.requires_grad]
        if not self.params:
            raise ValueError("no trainable parameters")
        dev, dt = self.params[0].device, self.params[0].dtype
        n = sum(p.numel() for p in self.params)
        self.flat = torch.zeros(n, dtype=dt, device=dev)
        off = 0
        for p in self.params:
            p.grad = self.flat[off:off + p.numel()].view_as(p)
            off += p.numel()

    def zero_(self):
        self.flat.zero_()

    def rebind(self):
        """Re-attach views if an optimiser's zero_grad(set_to_none=True) dropped them."""
        off = 0
        for p in self.params:
            if p.grad is None or p.grad.data_ptr() != self.flat[off:off + p.numel()].data_ptr():
                g = p.grad
                p.grad = self.flat[off:off + p.numel()].view_as(p)
                if g is not None:
                    p.grad.copy_(g)
            off += p.numel()

    def all_reduce_mean(self, group=None, async_op: bool = False):
        if not (dist.is_available() and dist.is_initialized()):
            return None
        world = dist.get_world_size(group)
        if world == 1:
            return None
        self.rebind()
        work = dist.all_reduce(self.flat, op=dist.ReduceOp.SUM, group=group, async_op=async_op)
        if async_op:
            return work
        self.flat.div_(world)
        return None

    def finish(self, work, group=None):
        if work is not None:
            work.wait()
            self.flat.div_(dist.get_world_size(group))


def broadcast_parameters(modules: Iterable[torch.nn.Module], src: int = 0, group=None):
    """All ranks start from rank `src`'s weights (one flat broadcast per module)."""
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
        return
    for m in modules:
        ps = [p.data for p in m.parameters()] + [b.data for b in m.buffers()]
        if not ps:
            continue
        flat = torch.cat([p.reshape(-1) for p in ps])
        dist.broadcast(flat, src=src, group=group)
        off = 0
        for p in ps:
            p.copy_(flat[off:off + p.numel()].view_as(p))
            off += p.numel()
