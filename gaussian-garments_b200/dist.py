"""One-view-per-GPU data parallelism: a flat gradient bucket and its per-step average over the ranks.

The reference renders one view per optimiser step on one GPU (s2_registration.py:241-251,
s3_appearance.py:107-124) and has no distributed code.  BASELINE.json's multi-GPU configs shard
views one per rank; the only exchange on the path is the sum of the per-view parameter gradients.

Design
  * all parameter gradients live in ONE contiguous fp32 buffer.  The rasterizer's backward kernel writes its outputs
    straight into views of that buffer (`rasterizer.GradSink`), autograd's AccumulateGrad adopts those views as
    `.grad` without a copy -- no per-tensor launches, no staging copies.
  * on an NVSwitch box the buffer is SYMMETRIC memory mapped behind one multicast address and the average is ONE
    hand-written kernel per rank (csrc/allreduce.cu behind gg_nvls_allreduce_f32): multimem.ld_reduce lets the switch
    add the ranks' copies, multimem.st lets it write the result back into every rank's bucket, 1/world fused in
    between.  torch.distributed._symmetric_memory is used for allocation / rendezvous only.  Anything else (gloo on
    CPU, no multicast support) falls back to the process group's all_reduce on the same ranges.
  * the exchange is split where the next step's data dependence is: the geometry block (means / scales / rotations /
    opacities, 13 MB at cfg2) is reduced on the compute stream; the `deferred` block (the SH coefficients, 57.6 MB)
    is reduced on a high-priority side stream and only the next forward's SH->RGB kernel waits for it
    (`rasterizer.COLOR_GATE`) -- projection, tile scan and instance emission of the next view do not read SH and
    overlap with it.  Consumers of the deferred gradients call `wait()`.
"""

from __future__ import annotations

import ctypes as C
import os
from typing import List, Optional, Sequence

import torch
import torch.distributed as dist


class GradBucket:
    """Flat fp32 gradient storage for `params` + its per-step average over `world_size` ranks."""

    def __init__(self, params: Sequence[torch.Tensor], world_size: int = 1, register: bool = True,
                 deferred: Sequence[int] = (), symmetric: Optional[bool] = None, group=None):
        self.params: List[torch.Tensor] = list(params)
        self.world = int(world_size)
        self.group = group
        dev = self.params[0].device
        self.device = dev
        # layout: immediate block first, deferred block last; every view 256-byte aligned
        self.deferred = tuple(sorted(set(int(i) for i in deferred)))
        order = [i for i in range(len(self.params)) if i not in self.deferred] + list(self.deferred)
        offs, total = [0] * len(self.params), 0
        self.split = None
        for i in order:
            if i in self.deferred and self.split is None:
                self.split = total
            offs[i] = total
            total += (self.params[i].numel() + 63) // 64 * 64
        if self.split is None:
            self.split = total
        self.offsets = offs
        self.numel = total
        self.impl = "none" if self.world <= 1 else "process_group"
        self.nvls_error = None
        self.flat = None
        self._hdl = None
        want_symm = symmetric if symmetric is not None else (self.world > 1 and dev.type == "cuda")
        if want_symm and self.world > 1 and dist.is_initialized() and dist.get_backend(group) == "nccl":
            try:
                self._setup_symmetric(total, dev)
            except Exception as e:           # no multicast support / symmetric memory unavailable: process-group path
                self.nvls_error = f"{type(e).__name__}: {e}"
                self.flat = None
        if self.flat is None:
            self.flat = torch.zeros(total, dtype=torch.float32, device=dev)
        self.comm_stream = None
        self.late_done = None
        self.device_gate = None
        if dev.type == "cuda" and self.world > 1 and self.split < total:
            self.comm_stream = torch.cuda.Stream(device=dev, priority=-1)
            self.late_done = torch.cuda.Event()
        if register:
            self.register()

    # -- symmetric memory / multicast set-up (plumbing) -------------------------------------------------
    def _setup_symmetric(self, total, dev):
        import torch.distributed._symmetric_memory as symm_mem
        flat = symm_mem.empty(total, dtype=torch.float32, device=dev)
        flat.zero_()
        grp = self.group if self.group is not None else dist.group.WORLD
        hdl = symm_mem.rendezvous(flat, grp)
        mc = int(getattr(hdl, "multicast_ptr", 0) or 0)
        if mc == 0:
            raise RuntimeError("symmetric memory has no multicast (NVLS) mapping on this system")
        self._hdl = hdl
        self._mc_ptr = mc
        self._pads_dev = int(hdl.signal_pad_ptrs_dev)
        pad_words = int(hdl.signal_pad_size) // 4
        blocks = max(2, min(72, pad_words // self.world))      # one pad word per (block, peer)
        self._blocks = (max(1, blocks // 3), max(1, blocks - blocks // 3))      # (immediate, deferred)
        env = os.environ.get("GG_AR_BLOCKS")                   # "a,b": experiment knob, a + b <= pad_words // world
        if env:
            a, b = (int(x) for x in env.split(","))
            if a >= 1 and b >= 1 and (a + b) * self.world <= pad_words and max(a, b) <= 148:
                self._blocks = (a, b)
        self._slot0 = (0, self._blocks[0] * self.world)
        self.flat = flat
        self.impl = "nvls_multimem"
        torch.cuda.synchronize(dev)
        dist.barrier(group=self.group)

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

    # -- the exchange -------------------------------------------------------------------------
    def _reduce_range(self, a: int, b: int, which: int):
        """Average flat[a:b] over the ranks on the CURRENT stream."""
        if b <= a:
            return
        if self.impl == "nvls_multimem":
            from . import _capi
            lib = _capi.load()
            di = self.device.index if self.device.index is not None else torch.cuda.current_device()
            sp = torch.cuda.current_stream(self.device).cuda_stream
            _capi.check(lib.gg_nvls_allreduce_f32(self._mc_ptr, self._pads_dev, dist.get_rank(self.group), self.world, a,
                                                  b - a, 1.0 / self.world, self._slot0[which], self._blocks[which], di, sp),
                        "gg_nvls_allreduce_f32")
            return
        seg = self.flat[a:b]
        if dist.get_backend(self.group) == "nccl":
            dist.all_reduce(seg, op=dist.ReduceOp.AVG, group=self.group)
        else:  # gloo (CPU tests): no AVG
            dist.all_reduce(seg, op=dist.ReduceOp.SUM, group=self.group)
            seg.div_(self.world)

    def all_reduce(self, overlap: bool = True):
        """The step's exchange: mean of the flat bucket over ranks.  The immediate block is reduced on the current
        stream; the deferred block on the side stream (its completion gates the next forward's colour kernel)."""
        if self.world <= 1 or not dist.is_initialized():
            return
        self.adopt()
        self.exchange_immediate()
        if self.split >= self.numel:
            return
        if self.comm_stream is None or not overlap:
            self._reduce_range(self.split, self.numel, 1)
            return
        self.exchange_deferred_async()

    def exchange_immediate(self):
        """Average the immediate (geometry) block on the current stream."""
        if self.world > 1 and dist.is_initialized():
            self._reduce_range(0, self.split, 0)

    def exchange_deferred_async(self):
        """Average the deferred (SH) block on the side stream, ordered after everything already enqueued on the current
        stream; the next forward's SH -> RGB kernel (and `wait()`) join it.  Inside a CUDA-graph capture this is called
        at the START of the captured step (it exchanges what the previous replay's backward left in the bucket), so that
        the fork / join stays inside one graph."""
        if self.world <= 1 or not dist.is_initialized() or self.split >= self.numel or self.comm_stream is None:
            return
        cur = torch.cuda.current_stream(self.device)
        ready = torch.cuda.Event()
        ready.record(cur)
        self.comm_stream.wait_event(ready)
        from . import rasterizer, _capi
        di = self.device.index if self.device.index is not None else torch.cuda.current_device()
        with torch.cuda.stream(self.comm_stream):
            self._reduce_range(self.split, self.numel, 1)
            if self.device_gate is not None:                # graph replays: advance X behind the exchange
                _capi.check(_capi.load().gg_gate_signal(self.device_gate.words.data_ptr(), di,
                                                        self.comm_stream.cuda_stream), "gg_gate_signal")
            self.late_done.record(self.comm_stream)
        if self.device_gate is None:
            rasterizer.COLOR_GATE[di] = self.late_done      # next forward: SH -> RGB waits, projection/binning do not

    def use_device_gate(self, on: bool = True):
        """Switch the colour gate to its device-side form (needed when the step is replayed as a CUDA graph while the
        exchanges are issued eagerly around it): every forward then takes a ticket, every deferred exchange advances
        the counter -- they must alternate strictly.  `on=False` returns to stream events."""
        from . import rasterizer
        di = self.device.index if self.device.index is not None else torch.cuda.current_device()
        torch.cuda.synchronize(self.device)
        if on:
            if self.device_gate is None:
                self.device_gate = rasterizer.DeviceGate(self.device)
            self.device_gate.reset()
            rasterizer.COLOR_GATE[di] = self.device_gate
        else:
            self.device_gate = None
            rasterizer.COLOR_GATE.pop(di, None)

    def wait(self):
        """Make the current stream wait for the deferred block (call before consuming those gradients)."""
        if self.late_done is not None and self.comm_stream is not None:
            torch.cuda.current_stream(self.device).wait_event(self.late_done)

    def check_against_gather(self, local_copy: torch.Tensor) -> dict:
        """Diagnostics (outside any timed region): all-gather every rank's LOCAL bucket (`local_copy`, taken before
        all_reduce) and compare its mean with what the exchange left in the bucket."""
        self.wait()
        torch.cuda.synchronize(self.device) if self.device.type == "cuda" else None
        parts = [torch.empty_like(local_copy) for _ in range(self.world)]
        dist.all_gather(parts, local_copy.contiguous(), group=self.group)
        mean = torch.stack(parts).double().mean(0)
        diff = (self.flat.double() - mean).abs().max()
        scale = mean.abs().max().clamp_min(1e-30)
        return {"impl": self.impl, "max_abs_err": float(diff), "rel_err": float(diff / scale), "elements": int(self.numel),
                "ok": bool(float(diff / scale) <= 1e-6)}


def shard_views(num_views: int, rank: int, world: int) -> List[int]:
    """View indices rendered by `rank`: r, r+G, r+2G, ...."""
    return list(range(rank, num_views, world))
