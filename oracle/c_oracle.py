"""TEST INFRASTRUCTURE -- ctypes front end of oracle/gg_oracle.c (see its header; parity unpinned).

`rasterize_forward` / `Context.backward` take and return CPU torch tensors so tests can compare
them one-to-one with the CUDA path and with oracle/torch_oracle.py.
"""

from __future__ import annotations

import ctypes as C
import os
import subprocess
from typing import Optional

import numpy as np
import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None


class _Params(C.Structure):
    _fields_ = [("N", C.c_int32), ("M", C.c_int32), ("D", C.c_int32), ("W", C.c_int32), ("H", C.c_int32),
                ("tanfovx", C.c_float), ("tanfovy", C.c_float), ("scale_modifier", C.c_float),
                ("bg", C.c_float * 3), ("viewmatrix", C.c_float * 16), ("projmatrix", C.c_float * 16),
                ("campos", C.c_float * 3), ("prefiltered", C.c_int32)]


def build(force: bool = False) -> str:
    so = os.path.join(_HERE, "libgg_oracle.so")
    src = os.path.join(_HERE, "gg_oracle.c")
    srcs = [src, os.path.join(_HERE, "raycast_oracle.c"), os.path.join(_HERE, "Makefile")]
    if force or not os.path.exists(so) or os.path.getmtime(so) < max(os.path.getmtime(p) for p in srcs):
        subprocess.check_call(["make", "-C", _HERE, "-B", "libgg_oracle.so"], stdout=subprocess.DEVNULL)
    return so


def lib():
    global _LIB
    if _LIB is None:
        _LIB = C.CDLL(build())
        _LIB.ggo_forward.restype = C.c_int
        _LIB.ggo_backward.restype = C.c_int
        _LIB.ggo_num_rendered.restype = C.c_int64
        _LIB.ggo_num_rendered.argtypes = [C.c_void_p]
        _LIB.ggo_free.argtypes = [C.c_void_p]
        _LIB.ggo_num_threads.restype = C.c_int
        # every entry point gets explicit argtypes: without them ctypes passes Python ints as 32-bit C ints,
        # which truncates / sign-extends a state handle as soon as the heap grows past 2 GB
        vp = C.c_void_p
        _LIB.ggo_forward.argtypes = [C.POINTER(_Params)] + [vp] * 8 + [vp] * 5 + [C.c_float, C.POINTER(vp)]
        _LIB.ggo_backward.argtypes = [vp] * 12
        _LIB.ggo_backward_forward_order.argtypes = [vp] * 12
        _LIB.ggo_backward_forward_order.restype = C.c_int
        _LIB.ggo_get_geom.argtypes = [vp] * 6
        _LIB.ggo_get_geom.restype = C.c_int
        _LIB.ggo_get_binning.argtypes = [vp] * 5
        _LIB.ggo_get_binning.restype = C.c_int
        _LIB.ggo_set_num_threads.argtypes = [C.c_int]
        _LIB.ggo_cast_rays_from_point.argtypes = [C.c_int32] * 3 + [vp] * 6
        _LIB.ggo_cast_rays_from_point.restype = C.c_int
    return _LIB


def _p(t: Optional[torch.Tensor]):
    return C.c_void_p(0) if t is None else C.c_void_p(t.data_ptr())


def _f32(t):
    if t is None:
        return None
    t = t.detach().to("cpu", torch.float32).contiguous()
    return t if t.numel() > 0 else None


def make_params(settings, N, M) -> _Params:
    (H, W, tanfovx, tanfovy, bg, mod, view, proj, deg, campos, prefiltered, _debug) = settings
    p = _Params()
    p.N, p.M, p.D, p.W, p.H = int(N), int(M), int(deg), int(W), int(H)
    p.tanfovx, p.tanfovy, p.scale_modifier = float(tanfovx), float(tanfovy), float(mod)
    p.bg[:] = [float(v) for v in bg.detach().cpu().reshape(-1)]
    p.viewmatrix[:] = [float(v) for v in view.detach().cpu().reshape(-1)]
    p.projmatrix[:] = [float(v) for v in proj.detach().cpu().reshape(-1)]
    p.campos[:] = [float(v) for v in campos.detach().cpu().reshape(-1)]
    p.prefiltered = int(bool(prefiltered))
    return p


class Context:
    """Holds the C-side state between forward and backward."""

    def __init__(self, handle, N, M, W, H, has_colors, has_cov):
        self.h, self.N, self.M, self.W, self.H = handle, N, M, W, H
        self.has_colors, self.has_cov = has_colors, has_cov

    @property
    def num_rendered(self) -> int:
        return int(lib().ggo_num_rendered(self.h))

    def geom(self):
        N = self.N
        xy = torch.zeros(N, 2); depth = torch.zeros(N); conic_o = torch.zeros(N, 4)
        rgb = torch.zeros(N, 3); rect = torch.zeros(N, 4, dtype=torch.int32)
        lib().ggo_get_geom(C.c_void_p(self.h), _p(xy), _p(depth), _p(conic_o), _p(rgb), _p(rect))
        return dict(xy=xy, depth=depth, conic_opacity=conic_o, rgb=rgb, rect=rect)

    def binning(self):
        T = ((self.W + 15) // 16) * ((self.H + 15) // 16)
        K = self.num_rendered
        off = torch.zeros(T + 1, dtype=torch.int64)
        inst = torch.zeros(max(K, 1), dtype=torch.int32)
        nc = torch.zeros(self.H, self.W, dtype=torch.int32)
        ft = torch.zeros(self.H, self.W)
        lib().ggo_get_binning(C.c_void_p(self.h), _p(off), _p(inst), _p(nc), _p(ft))
        return dict(tile_off=off, inst=inst[:K], n_contrib=nc, final_T=ft)

    def backward(self, dL_dcolor, dL_ddepth=None, dL_dalpha=None, forward_order=False):
        N, M = self.N, self.M
        gc, gd, ga = _f32(dL_dcolor), _f32(dL_ddepth), _f32(dL_dalpha)
        out = dict(
            means3D=torch.zeros(N, 3), means2D=torch.zeros(N, 3),
            shs=None if self.has_colors else torch.zeros(N, M, 3),
            colors_precomp=torch.zeros(N, 3) if self.has_colors else None,
            opacities=torch.zeros(N, 1),
            scales=None if self.has_cov else torch.zeros(N, 3),
            rotations=None if self.has_cov else torch.zeros(N, 4),
            cov3D_precomp=torch.zeros(N, 6) if self.has_cov else None)
        fn = lib().ggo_backward_forward_order if forward_order else lib().ggo_backward
        rc = fn(C.c_void_p(self.h), _p(gc), _p(gd), _p(ga), _p(out["means3D"]), _p(out["means2D"]),
                                _p(out["shs"]), _p(out["colors_precomp"]), _p(out["opacities"]),
                                _p(out["scales"]), _p(out["rotations"]), _p(out["cov3D_precomp"]))
        if rc != 0:
            raise RuntimeError(f"ggo_backward failed: {rc}")
        return out

    def close(self):
        if self.h:
            lib().ggo_free(C.c_void_p(self.h))
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def rasterize_forward(settings, means3D, shs=None, colors_precomp=None, opacities=None, scales=None,
                      rotations=None, cov3D_precomp=None, means2D=None, fragile_eps: float = 0.0):
    """-> (color[3,H,W], radii[N], depth[1,H,W], alpha[1,H,W], ctx, fragile[H,W] or None)"""
    H, W = int(settings[0]), int(settings[1])
    m3 = _f32(means3D)
    N = 0 if m3 is None else m3.shape[0]
    sh, col = _f32(shs), _f32(colors_precomp)
    M = sh.shape[1] if sh is not None else 0
    op, sc, ro, cv, m2 = _f32(opacities), _f32(scales), _f32(rotations), _f32(cov3D_precomp), _f32(means2D)
    p = make_params(settings, N, M)
    color = torch.zeros(3, H, W); depth = torch.zeros(1, H, W); alpha = torch.zeros(1, H, W)
    radii = torch.zeros(max(N, 1), dtype=torch.int32)
    frag = torch.zeros(H, W, dtype=torch.uint8) if fragile_eps > 0 else None
    handle = C.c_void_p()
    rc = lib().ggo_forward(C.byref(p), _p(m3), _p(sh), _p(col), _p(op), _p(sc), _p(ro), _p(cv), _p(m2),
                           _p(color), _p(depth), _p(alpha), _p(radii), _p(frag), C.c_float(fragile_eps),
                           C.byref(handle))
    if rc != 0:
        raise RuntimeError(f"ggo_forward failed: {rc}")
    ctx = Context(handle.value, N, M, W, H, col is not None, cv is not None)
    return color, radii[:N], depth, alpha, ctx, (frag.bool() if frag is not None else None)


def num_threads() -> int:
    return int(lib().ggo_num_threads())


def set_num_threads(n: int):
    lib().ggo_set_num_threads(int(n))


def cast_rays_from_point(verts, faces, targets, origin):
    """Row N3's oracle (oracle/raycast_oracle.c): first triangle hit by each ray origin -> target.
    -> (primitive_ids int32 [N] with -1 = no hit, t_hit float32 [N])"""
    v = verts.detach().to("cpu", torch.float32).contiguous()
    f = faces.detach().to("cpu", torch.int32).contiguous()
    t = targets.detach().to("cpu", torch.float32).contiguous()
    o = origin.detach().to("cpu", torch.float32).contiguous()
    N = t.shape[0]
    prim = torch.zeros(max(N, 1), dtype=torch.int32)
    th = torch.zeros(max(N, 1))
    rc = lib().ggo_cast_rays_from_point(v.shape[0], f.shape[0], N, _p(v), _p(f), _p(t), _p(o), _p(prim), _p(th))
    if rc != 0:
        raise RuntimeError("ggo_cast_rays_from_point failed")
    return prim[:N], th[:N]
