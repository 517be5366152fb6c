"""One-view-per-GPU data parallelism: a flat gradient bucket and its single all-reduce.

The reference renders one view per optimiser step on one GPU (s2_registration.py:241-251,
s3_appearance.py:107-124) and has no distributed code.  BASELINE.json's multi-GPU configs shard
views one per rank; the only exchange on the path is the sum of the per-view parameter gradients.

Design: all parameter gradients live in ONE contiguous fp32 buffer.  The rasterizer's
backward kernel writes its outputs straight into views of that buffer (see `rasterizer.GradSink`),
autograd's AccumulateGrad adopts those views as `.grad` without a copy, and a single
`all_reduce(AVG)` over the flat buffer is the step's only collective -- no per-tensor launches, no
staging copies.  NVSwitch gives every rank uniform bandwidth, so no topology tuning is needed.
"""

from __future__ import annotations

from typing import List, Sequence

import torch
import torch.distributed as dist


class GradBucket:
    """Flat fp32 gradient storage for `params` + its all-reduce."""

    def __init__(self, params: Sequence[torch.Tensor], world_size: int = 1, register: bool = True):
        self.params: List[torch.Tensor] = list(params)
        self.world = int(world_size)
        dev = self.params[0].device
        offs, total = [], 0
        for p in self.params:
            offs.append(total)
            total += (p.numel() + 63) // 64 * 64          # keep every view 256-byte aligned
        self.offsets = offs
        self.flat = torch.zeros(total, dtype=torch.float32, device=dev)
        self.numel = total
        if register:
            self.register()

    # -- zero-copy hand-off to the rasterizer's backward ------------------------------------
    def register(self):
        from .rasterizer import GradSink
        for p, o in zip(self.params, self.offsets):
            p._gg_sink = GradSink(self.flat, o, p.shape)      # on the tensor object: cannot alias a recycled address

    def unregister(self):
        for p in self.params:
            if getattr(p, "_gg_sink", None) is not None and p._gg_sink.flat is self.flat:
                del p._gg_sink

    def __del__(self):
        try:
            self.unregister()
        except Exception:
            pass

    def view(self, i: int) -> torch.Tensor:
        p, o = self.params[i], self.offsets[i]
        return self.flat[o:o + p.numel()].view(p.shape)

    def zero(self):
        """Drop stale .grad references; the backward kernel overwrites every element of its sink, so
        no memset is needed for sink-backed parameters."""
        for p in self.params:
            p.grad = None
            ent = getattr(p, "_gg_sink", None)
            if ent is not None:
                ent.armed = True       # one zero-copy hand-off per zero(); later backwards accumulate into it

    def adopt(self):
        """Make sure every param's .grad aliases the bucket (copies in the rare case autograd cloned)."""
        for i, p in enumerate(self.params):
            v = self.view(i)
            if p.grad is None:
                v.zero_()
                p.grad = v
            elif p.grad.data_ptr() != v.data_ptr():
                v.copy_(p.grad)
                p.grad = v

    def all_reduce(self):
        """The step's single collective: mean of the flat bucket over ranks."""
        if self.world <= 1 or not dist.is_initialized():
            return
        self.adopt()
        if dist.get_backend() == "nccl":
            dist.all_reduce(self.flat, op=dist.ReduceOp.AVG)
        else:  # gloo (CPU tests): no AVG
            dist.all_reduce(self.flat, op=dist.ReduceOp.SUM)
            self.flat.div_(self.world)


def shard_views(num_views: int, rank: int, world: int) -> List[int]:
    """View indices rendered by `rank`: r, r+G, r+2G, ...."""
    return list(range(rank, num_views, world))
