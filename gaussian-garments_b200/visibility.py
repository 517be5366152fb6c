"""On-device visibility masks ("next" row N3 of SURVEY.md 8f) -- host side.

Mirrors the two `get_visible_mask` methods of the reference, which build the mask that `render(..., vis_mask=)` /
`doll_render(..., vis_mask=)` use to drop occluded Gaussians (gaussian_renderer/__init__.py:92-100):

    AvatarGaussianModel.get_visible_mask(camera)                    scene/avatar_gaussian_model.py:227-263
        rays from the camera centre to every Gaussian's barycentric point on the garment mesh; visible iff the first
        triangle hit is the one the Gaussian is bound to                                 -> visible_mask(...)
    Simulation.get_visible_mask(camera, xyz, binding)               inference.py:285-316
        several garment meshes in one scene; visible iff the first hit belongs to the Gaussian's own garment, or
        nothing is hit                                                                   -> visible_mask_multi(...)

The reference copies mesh and points to the host, runs open3d's RaycastingScene (CPU Embree) and copies the mask
back -- a GPU->CPU->GPU round trip before every s3 / inference render.  Here the cast is the hand-written CUDA of
csrc/visibility.cu behind gg_cast_rays_from_point; nothing leaves the device and nothing synchronises.
"""

from __future__ import annotations

import ctypes as C
from typing import Optional, Sequence, Tuple

import torch

from . import _capi


def _f32(t: torch.Tensor, name: str) -> torch.Tensor:
    if not t.is_cuda:
        raise RuntimeError(f"gaussian-garments_b200: `{name}` must be a CUDA tensor (there is no CPU path)")
    t = t.detach()
    t = t.float() if t.dtype != torch.float32 else t
    return t.contiguous()


def cast_rays_from_point(verts: torch.Tensor, faces: torch.Tensor, targets: torch.Tensor, origin: torch.Tensor,
                         look_at: Optional[torch.Tensor] = None, force_bruteforce: bool = False,
                         return_t: bool = False):
    """First triangle hit by each ray origin -> targets[i] (what `scene.cast_rays(rays)['primitive_ids']` holds in the
    reference, with -1 for "no hit").  verts [V,3], faces [F,3] (any int dtype), targets [N,3], origin [3]: CUDA."""
    lib = _capi.load()
    v, t, o = _f32(verts, "verts"), _f32(targets, "targets"), _f32(origin.reshape(-1), "origin")
    dev = v.device
    f = faces.detach()
    f = (f if f.dtype == torch.int32 else f.to(torch.int32)).contiguous().to(dev)
    if v.dim() != 2 or v.shape[1] != 3 or t.dim() != 2 or t.shape[1] != 3 or o.numel() != 3 or f.dim() != 2 or f.shape[1] != 3:
        raise RuntimeError("cast_rays_from_point expects verts [V,3], faces [F,3], targets [N,3], origin [3]")
    la = _f32((v.mean(0) if look_at is None else look_at).reshape(-1), "look_at")
    di = dev.index if dev.index is not None else torch.cuda.current_device()
    V, F, N = v.shape[0], f.shape[0], t.shape[0]
    wb, cap = C.c_size_t(), C.c_int64()
    _capi.check(lib.gg_cast_rays_workspace_bytes(V, F, C.byref(wb), C.byref(cap)), "gg_cast_rays_workspace_bytes")
    ws = torch.empty(max(wb.value, 16), dtype=torch.uint8, device=dev)
    prim = torch.empty(N, dtype=torch.int32, device=dev)
    th = torch.empty(N, dtype=torch.float32, device=dev) if return_t else None
    with torch.cuda.device(dev):
        sp = torch.cuda.current_stream(dev).cuda_stream
        _capi.check(lib.gg_cast_rays_from_point(V, F, N, v.data_ptr(), f.data_ptr(), t.data_ptr(), o.data_ptr(),
                                                la.data_ptr(), ws.data_ptr(), cap.value, 1 if force_bruteforce else 0,
                                                prim.data_ptr(), None if th is None else th.data_ptr(), di, sp),
                    "gg_cast_rays_from_point")
    return (prim, th) if return_t else prim


def visible_mask(camera_center: torch.Tensor, verts: torch.Tensor, faces: torch.Tensor, points: torch.Tensor,
                 binding: torch.Tensor) -> torch.Tensor:
    """AvatarGaussianModel.get_visible_mask (scene/avatar_gaussian_model.py:227-263): `points` are the Gaussians'
    barycentric anchors (get_barycentric_3d), `binding` their faces; True where the first hit is the bound face."""
    prim = cast_rays_from_point(verts, faces, points, camera_center)
    return prim == binding.to(prim.device).to(torch.int32)


def visible_mask_multi(camera_center: torch.Tensor, meshes: Sequence[Tuple[torch.Tensor, torch.Tensor]],
                       points: torch.Tensor, geometry_ids: torch.Tensor) -> torch.Tensor:
    """Simulation.get_visible_mask (inference.py:285-316): `meshes` = [(verts_g, faces_g)] per garment, all in one
    scene; geometry_ids[i] = the garment Gaussian i belongs to.  True where the first hit lies on the Gaussian's own
    garment or nothing is hit (`ans['geometry_ids'] >= num_gs` in the reference = INVALID_ID)."""
    dev = points.device
    vs, fs, gids, voff = [], [], [], 0
    for g, (v, f) in enumerate(meshes):
        vs.append(v.detach().float().to(dev))
        fs.append(f.detach().to(dev).long() + voff)
        gids.append(torch.full((f.shape[0],), g, dtype=torch.int32, device=dev))
        voff += v.shape[0]
    verts, faces, tri_geom = torch.cat(vs), torch.cat(fs), torch.cat(gids)
    prim = cast_rays_from_point(verts, faces, points, camera_center)
    hit = prim >= 0
    geom = tri_geom[prim.clamp_min(0).long()]
    return (geom == geometry_ids.to(dev).to(torch.int32)) | ~hit
