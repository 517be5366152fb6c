"""Host-side mirror of the rasterizer interface the reference binds to.

Same names, argument meaning and error behaviour as the third-party package the reference
imports at /root/reference/gaussian_renderer/__init__.py:16
(`from diff_gaussian_rasterization_depth_alpha import GaussianRasterizationSettings, GaussianRasterizer`):

  GaussianRasterizationSettings   12-field NamedTuple built by keyword at
                                  gaussian_renderer/__init__.py:39-52 and :142-155
  GaussianRasterizer(raster_settings=...)(means3D=, means2D=, shs=, colors_precomp=, opacities=,
      scales=, rotations=, cov3D_precomp=) -> (color[3,H,W], radii[N] int32, depth[1,H,W], alpha[1,H,W])
                                  gaussian_renderer/__init__.py:54,103-111 and :208-216

PyTorch is plumbing only (device memory, streams, autograd glue); the arithmetic is the
hand-written sm_100a CUDA behind the C ABI of include/gg_raster.h, reached through ctypes.
There is no CPU path: non-CUDA tensors or a missing library raise.
"""

from __future__ import annotations

import ctypes as C
import threading
import time
import weakref
from typing import NamedTuple, Optional

import torch
from torch import nn

from . import _capi
from ._capi import GGInputs, GGView


class GaussianRasterizationSettings(NamedTuple):
    image_height: int
    image_width: int
    tanfovx: float
    tanfovy: float
    bg: torch.Tensor
    scale_modifier: float
    viewmatrix: torch.Tensor
    projmatrix: torch.Tensor
    sh_degree: int
    campos: torch.Tensor
    prefiltered: bool
    debug: bool


# one pinned word per (thread, device) for the num_rendered read-back
_tls = threading.local()

LAST_NUM_RENDERED = 0   # K of the most recent forward (bench/diagnostics)
STATS = {"forwards": 0, "hinted": 0, "overflow_retries": 0, "k_wait_s": 0.0}
DEBUG_CAPTURE = None    # tests set this to a dict: the next forward leaves its binning workspaces in it
# device index -> CUDA event the SH -> RGB kernel must wait for (dist.GradBucket: the deferred SH-gradient exchange of the
# previous step runs on a side stream while this forward's projection / tile scan are already executing)
COLOR_GATE = {}


class DeviceGate:
    """Graph-replay form of the colour gate: three device words {G, X, timeout} (see gg_forward_render_late_color)."""

    def __init__(self, device):
        self.words = torch.zeros(4, dtype=torch.int32, device=device)

    def reset(self):
        self.words.zero_()

    def timed_out(self) -> bool:
        return bool(int(self.words[2]) != 0)


class GradSink:
    """Destination of one leaf tensor's gradient inside a flat bucket (dist.GradBucket): when a leaf carries
    one (attribute `_gg_sink` on the tensor OBJECT -- never keyed by address, so a recycled allocation cannot
    alias it), backward writes that input's gradient straight into the bucket and autograd adopts the view as
    `.grad` without a copy.  A sink is used at most ONCE per arming (GradBucket.zero() arms it) and only while
    `leaf.grad is None`: every further backward before the next zero() returns a fresh tensor, which autograd
    accumulates into the adopted view -- so gradient accumulation over several backward() calls stays exact."""

    __slots__ = ("flat", "offset", "shape", "armed")

    def __init__(self, flat, offset, shape):
        self.flat, self.offset, self.shape, self.armed = flat, int(offset), tuple(shape), False


def _sink_ref(t):
    """weakref to an input that carries a usable sink (None otherwise)."""
    if t is None or not torch.is_tensor(t) or t.numel() == 0:
        return None
    ent = getattr(t, "_gg_sink", None)
    if ent is None or tuple(t.shape) != ent.shape or not t.is_contiguous() or not t.is_leaf:
        return None
    return weakref.ref(t)


# Instance-capacity hints: (device, N, W, H) -> [capacity, largest per-tile count].  With a hint the forward is
# enqueued WITHOUT waiting for num_rendered (upstream blocks on a D2H copy in the middle of every forward,
# SURVEY.md 3.1); K is read once everything is queued and the rare overflow re-runs the instance stages.
_hints = {}


def _capacity_for(K: int) -> int:
    return int(K * 1.25) + 4096


# CUDA-graph capture: nothing may synchronise, so the forward runs entirely from the hint (a previous eager call with
# the same (device, N, W, H) must have happened) with extra head-room, and gg_forward_overflow_check leaves the verdict
# in two sticky device words per device: [overflowed?, largest K seen].  `graph_overflow(device)` reads them.
GRAPH_HEADROOM = 1.5
_graph_flags = {}


def _graph_flag(dev) -> torch.Tensor:
    di = dev.index if dev.index is not None else torch.cuda.current_device()
    f = _graph_flags.get(di)
    if f is None:
        f = _graph_flags[di] = torch.zeros(2, dtype=torch.int32, device=dev)
    return f


def graph_overflow(device=None):
    """(overflowed: bool, largest num_rendered seen) of the sync-free forwards replayed on `device` so far (this
    call synchronises).  After an overflow: run one eager forward (refreshes the capacity hint) and re-capture."""
    dev = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)
    f = _graph_flag(dev).tolist()
    return bool(f[0]), int(f[1]) & 0xFFFFFFFF


def _pinned_word(device_index: int) -> torch.Tensor:
    cache = getattr(_tls, "words", None)
    if cache is None:
        cache = _tls.words = {}
    w = cache.get(device_index)
    if w is None:
        w = cache[device_index] = torch.zeros(2, dtype=torch.int32).pin_memory()
    return w


def _ptr(t: Optional[torch.Tensor]):
    return None if t is None else t.data_ptr()


def _prep(t: Optional[torch.Tensor], name: str, device) -> Optional[torch.Tensor]:
    """None / empty -> None; otherwise fp32, contiguous, on `device`."""
    if t is None or t.numel() == 0:
        return None
    if not t.is_cuda:
        raise RuntimeError(f"gaussian-garments_b200: `{name}` must be a CUDA tensor (there is no CPU path)")
    if t.device != device:
        raise RuntimeError(f"gaussian-garments_b200: `{name}` is on {t.device}, expected {device}")
    if t.dtype != torch.float32:
        t = t.float()
    t = t.contiguous()
    if t.data_ptr() % 16 != 0:          # a view at an odd offset: vector loads need 16-byte alignment
        t = t.clone()
    return t


def _view_struct(s: GaussianRasterizationSettings, N: int, M: int) -> GGView:
    return GGView(int(N), int(M), int(s.sh_degree), int(s.image_width), int(s.image_height),
                  float(s.tanfovx), float(s.tanfovy), float(s.scale_modifier),
                  int(bool(s.prefiltered)), int(bool(s.debug)))


def _ws(nbytes: int, device, zero: bool = False) -> torch.Tensor:
    f = torch.zeros if zero else torch.empty
    return f(max(int(nbytes), 16), dtype=torch.uint8, device=device)


class _RasterizeGaussians(torch.autograd.Function):
    """Replaces upstream's `_RasterizeGaussians` autograd Function."""

    @staticmethod
    def forward(ctx, means3D, means2D, sh, colors_precomp, opacities, scales, rotations, cov3Ds_precomp,
                raster_settings: GaussianRasterizationSettings):
        lib = _capi.load()
        s = raster_settings
        if not means3D.is_cuda:
            raise RuntimeError("gaussian-garments_b200: `means3D` must be a CUDA tensor (there is no CPU path)")
        if means3D.dim() != 2 or means3D.shape[-1] != 3:
            raise RuntimeError("means3D must have dimensions (num_points, 3)")
        dev = means3D.device
        di = dev.index if dev.index is not None else torch.cuda.current_device()
        N = means3D.shape[0]
        H, W = int(s.image_height), int(s.image_width)

        m3 = _prep(means3D, "means3D", dev)
        shs = _prep(sh, "shs", dev)
        col = _prep(colors_precomp, "colors_precomp", dev)
        op = _prep(opacities, "opacities", dev)
        sc = _prep(scales, "scales", dev)
        ro = _prep(rotations, "rotations", dev)
        cv = _prep(cov3Ds_precomp, "cov3D_precomp", dev)
        bg = _prep(s.bg, "bg", dev)
        vm = _prep(s.viewmatrix, "viewmatrix", dev)
        pm = _prep(s.projmatrix, "projmatrix", dev)
        cp = _prep(s.campos, "campos", dev)
        M = int(shs.shape[1]) if shs is not None else 0
        view = _view_struct(s, N, M)
        inputs = GGInputs(_ptr(m3), _ptr(shs), _ptr(col), _ptr(op), _ptr(sc), _ptr(ro), _ptr(cv),
                          _ptr(bg), _ptr(vm), _ptr(pm), _ptr(cp))

        color = torch.empty(3, H, W, dtype=torch.float32, device=dev)
        depth = torch.empty(1, H, W, dtype=torch.float32, device=dev)
        alpha = torch.empty(1, H, W, dtype=torch.float32, device=dev)
        radii = torch.empty(N, dtype=torch.int32, device=dev)

        try:
            with torch.cuda.device(dev):
                stream = torch.cuda.current_stream(dev)
                sp = stream.cuda_stream
                gb, tb, ib = C.c_size_t(), C.c_size_t(), C.c_size_t()
                _capi.check(lib.gg_forward_workspace_bytes(C.byref(view), C.byref(gb), C.byref(tb), C.byref(ib)),
                            "gg_forward_workspace_bytes")
                geom_ws, tile_ws, image_ws = _ws(gb.value, dev), _ws(tb.value, dev), _ws(ib.value, dev)
                K, max_tile, capacity = 0, 0, 0
                key_ws = record_ws = None

                late = {"on": False, "gate": None}

                def render(cap, mt):
                    kb, rb = C.c_size_t(), C.c_size_t()
                    _capi.check(lib.gg_instance_workspace_bytes(cap, C.byref(kb), C.byref(rb)),
                                "gg_instance_workspace_bytes")
                    kw, rw = _ws(kb.value, dev), _ws(rb.value, dev)
                    if late["on"]:       # colours after the sort, behind the gate (the SH exchange of the previous step)
                        _capi.check(lib.gg_forward_render_late_color(
                            C.byref(view), C.byref(inputs), geom_ws.data_ptr(), tile_ws.data_ptr(), kw.data_ptr(),
                            rw.data_ptr(), cap, mt, image_ws.data_ptr(), _ptr(radii), color.data_ptr(), depth.data_ptr(),
                            alpha.data_ptr(),
                            late["gate"].cuda_event if isinstance(late["gate"], torch.cuda.Event) else None,
                            late["gate"].words.data_ptr() if isinstance(late["gate"], DeviceGate) else None, di, sp),
                            "gg_forward_render_late_color")
                        late["gate"] = None          # a retry after an overflow must not wait again (already passed)
                    else:
                        _capi.check(lib.gg_forward_render(C.byref(view), C.byref(inputs), geom_ws.data_ptr(),
                                                          tile_ws.data_ptr(), kw.data_ptr(), rw.data_ptr(), cap, mt,
                                                          image_ws.data_ptr(), _ptr(radii), color.data_ptr(),
                                                          depth.data_ptr(), alpha.data_ptr(), di, sp), "gg_forward_render")
                    return kw, rw

                if N > 0:
                    STATS["forwards"] += 1
                    capturing = torch.cuda.is_current_stream_capturing()
                    word = _pinned_word(di)
                    if not capturing:
                        _graph_flag(dev)         # the sticky overflow words must exist before any capture starts
                    hkey = (di, N, W, H)
                    hint = None if s.debug else _hints.get(hkey)
                    gate = COLOR_GATE.get(di)
                    if gate is not None and hint is None and not isinstance(gate, DeviceGate):
                        # unhinted call (first of its shape): no late-colour order; the colour kernel may be forked onto
                        # the library's side stream right behind the projection, so the whole forward waits here
                        stream.wait_event(gate)
                    _capi.check(lib.gg_forward_project(C.byref(view), C.byref(inputs), geom_ws.data_ptr(),
                                                       tile_ws.data_ptr(), radii.data_ptr(),
                                                       None if capturing else word.data_ptr(), di, sp),
                                "gg_forward_project")
                    if not capturing:
                        k_ready = torch.cuda.Event()
                        k_ready.record(stream)
                    # A pending SH-gradient exchange (dist.GradBucket) gates only the colour stage.  With a capacity
                    # hint the colour stage moves BEHIND emission and sorting (late-colour call), so the exchange
                    # overlaps with projection + emit + sort; without a hint it simply waits here.
                    late["on"] = gate is not None and hint is not None
                    late["gate"] = gate if late["on"] else None
                    if isinstance(gate, DeviceGate) and not late["on"]:
                        raise RuntimeError("gaussian-garments_b200: a device colour gate needs the hinted (late-colour) "
                                           "forward: run one eager forward of this shape first")
                    if not late["on"]:
                        _capi.check(lib.gg_forward_color(C.byref(view), C.byref(inputs), geom_ws.data_ptr(),
                                                         radii.data_ptr(), di, sp), "gg_forward_color")
                    if capturing:
                        if hint is None:
                            raise RuntimeError("gaussian-garments_b200: CUDA-graph capture needs one eager forward with "
                                               "the same (N, width, height) first (it sizes the instance workspaces)")
                        capacity = int(hint[0] * GRAPH_HEADROOM)
                        key_ws, record_ws = render(capacity, hint[1])
                        _capi.check(lib.gg_forward_overflow_check(C.byref(view), tile_ws.data_ptr(), capacity,
                                                                  _graph_flag(dev).data_ptr(), di, sp),
                                    "gg_forward_overflow_check")
                        K, max_tile = -1, hint[1]
                    elif hint is not None:
                        # everything is enqueued before the host looks at K: the GPU never waits for the host
                        STATS["hinted"] += 1
                        capacity = hint[0]
                        key_ws, record_ws = render(capacity, hint[1])
                    if not capturing:
                        _t0 = time.perf_counter()
                        k_ready.synchronize()        # scan + 8-byte copy only; later kernels keep running
                        STATS["k_wait_s"] += time.perf_counter() - _t0
                        K, max_tile = (int(v) & 0xFFFFFFFF for v in word.tolist())
                        if hint is None or K > capacity:
                            if hint is not None:
                                STATS["overflow_retries"] += 1
                            capacity = K
                            key_ws, record_ws = render(capacity, max_tile)
                        _hints[hkey] = [max(_capacity_for(K), int(0.9 * (hint[0] if hint else 0))), max_tile]
                else:
                    key_ws, record_ws = render(0, 0)
        except Exception:
            if s.debug:
                torch.save(dict(means3D=means3D, sh=sh, colors_precomp=colors_precomp, opacities=opacities,
                                scales=scales, rotations=rotations, cov3Ds_precomp=cov3Ds_precomp,
                                raster_settings=tuple(s)), "snapshot_fw.dump")
                print("\nAn error occured in forward. Please forward snapshot_fw.dump for debugging.")
            raise

        global LAST_NUM_RENDERED
        LAST_NUM_RENDERED = K
        if DEBUG_CAPTURE is not None:
            DEBUG_CAPTURE.update(view=view, tile_ws=tile_ws, record_ws=record_ws, capacity=capacity, K=K,
                                 max_tile=max_tile, device=di)
        ctx.raster_settings = s
        ctx.num_rendered = capacity          # record_ws was laid out for this many instances
        ctx.set_materialize_grads(False)     # unused outputs (depth / alpha in the training loops) arrive as None
        ctx.sinks = tuple(_sink_ref(t) for t in (means3D, sh, colors_precomp, opacities, scales, rotations,
                                                 cov3Ds_precomp))
        ctx.shapes = (N, M)
        ctx.has = (shs is not None, col is not None, sc is not None, cv is not None)
        # NOTE: `color` is deliberately NOT saved: the reference's ssim() multiplies it in place before
        # backward (utils/loss_utils.py:45); backward needs only final_T / n_contrib from image_ws.
        saved = [t if t is not None else torch.empty(0, device=dev) for t in
                 (m3, shs, col, op, sc, ro, cv, bg, vm, pm, cp)]
        ctx.save_for_backward(*saved, radii, tile_ws, record_ws, image_ws)
        ctx.mark_non_differentiable(radii)
        return color, radii, depth, alpha

    @staticmethod
    def backward(ctx, grad_color, grad_radii, grad_depth, grad_alpha):
        lib = _capi.load()
        s = ctx.raster_settings
        N, M = ctx.shapes
        has_sh, has_col, has_sr, has_cov = ctx.has
        (m3, shs, col, op, sc, ro, cv, bg, vm, pm, cp, radii, tile_ws, record_ws, image_ws) = ctx.saved_tensors
        dev = m3.device
        di = dev.index if dev.index is not None else torch.cuda.current_device()
        need = ctx.needs_input_grad  # means3D, means2D, sh, colors, opac, scales, rots, cov, settings

        def opt(t):
            return t if t.numel() > 0 else None

        shs, col, sc, ro, cv = opt(shs), opt(col), opt(sc), opt(ro), opt(cv)
        if N == 0:
            z = lambda t: None if t is None else torch.zeros_like(t)
            return (torch.zeros_like(m3), torch.zeros(0, 3, device=dev), z(shs), z(col), torch.zeros_like(op),
                    z(sc), z(ro), z(cv), None)

        def gprep(g):
            if g is None:
                return None
            g = g.float() if g.dtype != torch.float32 else g
            return g.contiguous()

        gc, gd, ga = gprep(grad_color), gprep(grad_depth), gprep(grad_alpha)
        if gc is not None and gc.data_ptr() % 16:
            gc = gc.clone()
        view = _view_struct(s, N, M)
        inputs = GGInputs(_ptr(m3), _ptr(shs), _ptr(col), _ptr(op), _ptr(sc), _ptr(ro), _ptr(cv),
                          _ptr(bg), _ptr(vm), _ptr(pm), _ptr(cp))
        sinks = ctx.sinks

        def E(slot, *shape):
            leaf = sinks[slot]() if (slot is not None and sinks[slot] is not None) else None
            ent = getattr(leaf, "_gg_sink", None) if leaf is not None else None
            if ent is not None and ent.armed and leaf.grad is None and ent.shape == tuple(shape):
                ent.armed = False                            # at most one zero-copy hand-off per arming
                n = 1
                for d in shape:
                    n *= d
                return ent.flat[ent.offset:ent.offset + n].view(*shape)    # fresh view: autograd adopts it as .grad
            return torch.empty(*shape, dtype=torch.float32, device=dev)

        g_m3 = E(0, N, 3) if need[0] else None
        g_m2 = E(None, N, 3) if need[1] else None
        g_sh = E(1, N, M, 3) if (has_sh and need[2]) else None
        g_col = E(2, N, 3) if (has_col and need[3]) else None
        g_op = E(3, N, 1) if need[4] else None
        g_sc = E(4, N, 3) if (has_sr and need[5]) else None
        g_ro = E(5, N, 4) if (has_sr and need[6]) else None
        g_cv = E(6, N, 6) if (has_cov and need[7]) else None
        try:
            with torch.cuda.device(dev):
                sp = torch.cuda.current_stream(dev).cuda_stream
                ab = C.c_size_t()
                _capi.check(lib.gg_backward_workspace_bytes(C.byref(view), C.byref(ab)), "gg_backward_workspace_bytes")
                accum_ws = _ws(ab.value, dev)         # zero-filled by gg_backward itself
                _capi.check(lib.gg_backward(C.byref(view), C.byref(inputs), tile_ws.data_ptr(), record_ws.data_ptr(),
                                            ctx.num_rendered, image_ws.data_ptr(), radii.data_ptr(), accum_ws.data_ptr(),
                                            _ptr(gc), _ptr(gd), _ptr(ga), _ptr(g_m3), _ptr(g_m2), _ptr(g_sh),
                                            _ptr(g_col), _ptr(g_op), _ptr(g_sc), _ptr(g_ro), _ptr(g_cv), di, sp),
                            "gg_backward")
        except Exception:
            if s.debug:
                torch.save(dict(grad_color=grad_color, grad_depth=grad_depth, grad_alpha=grad_alpha,
                                raster_settings=tuple(s)), "snapshot_bw.dump")
                print("\nAn error occured in backward. Writing snapshot_bw.dump for debugging.\n")
            raise
        return (g_m3, g_m2, g_sh, g_col, g_op, g_sc, g_ro, g_cv, None)


def rasterize_gaussians(means3D, means2D, sh, colors_precomp, opacities, scales, rotations, cov3Ds_precomp,
                        raster_settings):
    return _RasterizeGaussians.apply(means3D, means2D, sh, colors_precomp, opacities, scales, rotations,
                                     cov3Ds_precomp, raster_settings)


def mark_visible(positions: torch.Tensor, viewmatrix: torch.Tensor, projmatrix: torch.Tensor) -> torch.Tensor:
    lib = _capi.load()
    if not positions.is_cuda:
        raise RuntimeError("gaussian-garments_b200: `positions` must be a CUDA tensor (there is no CPU path)")
    dev = positions.device
    di = dev.index if dev.index is not None else torch.cuda.current_device()
    p = _prep(positions, "positions", dev)
    vm = _prep(viewmatrix, "viewmatrix", dev)
    pm = _prep(projmatrix, "projmatrix", dev)
    N = positions.shape[0]
    vis = torch.zeros(N, dtype=torch.uint8, device=dev)
    with torch.cuda.device(dev):
        sp = torch.cuda.current_stream(dev).cuda_stream
        _capi.check(lib.gg_mark_visible(N, _ptr(p), _ptr(vm), _ptr(pm), vis.data_ptr(), di, sp), "gg_mark_visible")
    return vis.bool()


class GaussianRasterizer(nn.Module):
    def __init__(self, raster_settings: GaussianRasterizationSettings):
        super().__init__()
        self.raster_settings = raster_settings

    def markVisible(self, positions):
        with torch.no_grad():
            rs = self.raster_settings
            return mark_visible(positions, rs.viewmatrix, rs.projmatrix)

    def forward(self, means3D, means2D, opacities, shs=None, colors_precomp=None, scales=None, rotations=None,
                cov3D_precomp=None):
        rs = self.raster_settings
        if (shs is None and colors_precomp is None) or (shs is not None and colors_precomp is not None):
            raise Exception("Please provide exactly one of either SHs or precomputed colors!")
        if ((scales is None or rotations is None) and cov3D_precomp is None) or \
                ((scales is not None or rotations is not None) and cov3D_precomp is not None):
            raise Exception("Please provide exactly one of either scale/rotation pair or precomputed 3D covariance!")
        empty = torch.Tensor([])
        shs = empty if shs is None else shs
        colors_precomp = empty if colors_precomp is None else colors_precomp
        scales = empty if scales is None else scales
        rotations = empty if rotations is None else rotations
        cov3D_precomp = empty if cov3D_precomp is None else cov3D_precomp
        return rasterize_gaussians(means3D, means2D, shs, colors_precomp, opacities, scales, rotations,
                                   cov3D_precomp, rs)
