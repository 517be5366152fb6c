"""Fused mesh-binding transform ("next" row N1 of SURVEY.md 8f) -- host side.

Replaces, behind one autograd op, the per-iteration torch chain of the reference's `MeshGaussianModel`
  update_face_coor()                scene/mesh_gaussian_model.py:90-95  (+ utils/graphics_utils.py:118-137)
  get_xyz / get_scaling / get_rotation   scene/mesh_gaussian_model.py:105-128
i.e. the step immediately before the rasterizer call (gaussian_renderer/__init__.py:56,72-73):

    xyz, scaling, rotation = bind_to_mesh(mesh_v, mesh_f, binding, _xyz, _scaling, _rotation)

Gradients flow to `mesh_v` (what stage 2 optimises and all-reduces, scene/mesh_gaussian_model.py:366-371) and to
the three local parameter tensors.  The arithmetic is the hand-written CUDA of csrc/mesh_binding.cu behind the
C ABI (gg_mesh_bind_forward / gg_mesh_bind_backward); there is no CPU path.
"""

from __future__ import annotations

import ctypes as C

import torch

from . import _capi


def _i32(t: torch.Tensor) -> torch.Tensor:
    return t if t.dtype == torch.int32 else t.to(torch.int32)


def _f32c(t: torch.Tensor, name: str) -> torch.Tensor:
    if not t.is_cuda:
        raise RuntimeError(f"gaussian-garments_b200: `{name}` must be a CUDA tensor (there is no CPU path)")
    t = t.float() if t.dtype != torch.float32 else t
    t = t.contiguous()
    return t.clone() if t.data_ptr() % 16 else t


def _index_tensor(t: torch.Tensor, name: str, device, upper: int, count=None) -> torch.Tensor:
    """int32, contiguous, on `device`; the reference's torch indexing raises IndexError on a bad index, a raw
    device pointer would silently read/atomically write out of bounds -- so the range is validated here (one tiny
    reduction per *new* index tensor: FusedMeshBinding caches the converted tensors)."""
    t = _i32(t).contiguous().to(device)
    if count is not None and t.shape[0] != count:
        raise IndexError(f"gaussian-garments_b200: `{name}` has {t.shape[0]} entries, expected {count}")
    if t.numel() > 0:
        lo, hi = int(t.min()), int(t.max())
        if lo < 0 or hi >= upper:
            raise IndexError(f"gaussian-garments_b200: `{name}` holds index {lo if lo < 0 else hi}, valid range is [0, {upper})")
    return t


class _MeshBind(torch.autograd.Function):
    @staticmethod
    def forward(ctx, mesh_v, faces_i32, binding_i32, local_xyz, local_scaling, local_rotation, barycentric=None,
                face_scaling_remembered=None):
        lib = _capi.load()
        v = _f32c(mesh_v, "mesh_v")
        lx, ls, lr = _f32c(local_xyz, "_xyz"), _f32c(local_scaling, "_scaling"), _f32c(local_rotation, "_rotation")
        bc = None if barycentric is None else _f32c(barycentric.detach(), "barycentric")
        sr = None if face_scaling_remembered is None else _f32c(face_scaling_remembered.detach().reshape(-1), "face_scaling_remembered")
        dev = v.device
        di = dev.index if dev.index is not None else torch.cuda.current_device()
        V, F, N = v.shape[0], faces_i32.shape[0], lx.shape[0]
        for t, name in ((faces_i32, "mesh.f"), (binding_i32, "binding")):
            if t.device != dev or t.dtype != torch.int32:
                raise RuntimeError(f"gaussian-garments_b200: `{name}` must be an int32 tensor on {dev}")
        if binding_i32.shape[0] != N or ls.shape[0] != N or lr.shape[0] != N:
            raise RuntimeError("gaussian-garments_b200: binding / _xyz / _scaling / _rotation disagree on the number of Gaussians")
        if bc is not None and tuple(bc.shape) != (N, 3):
            raise RuntimeError("gaussian-garments_b200: barycentric must be [N,3]")
        if sr is not None and sr.shape[0] != F:
            raise RuntimeError("gaussian-garments_b200: face_scaling_remembered must have one entry per face")
        fb = C.c_size_t()
        _capi.check(lib.gg_mesh_bind_workspace_bytes(F, C.byref(fb)), "gg_mesh_bind_workspace_bytes")
        frames = torch.empty(fb.value, dtype=torch.uint8, device=dev)
        xyz = torch.empty(N, 3, device=dev)
        scaling = torch.empty(N, 3, device=dev)
        rotation = torch.empty(N, 4, device=dev)
        p = lambda t: None if t is None else t.data_ptr()
        with torch.cuda.device(dev):
            sp = torch.cuda.current_stream(dev).cuda_stream
            _capi.check(lib.gg_mesh_bind_forward_ex(V, F, N, v.data_ptr(), faces_i32.data_ptr(), binding_i32.data_ptr(),
                                                    lx.data_ptr(), ls.data_ptr(), lr.data_ptr(), p(bc), p(sr),
                                                    frames.data_ptr(), xyz.data_ptr(), scaling.data_ptr(),
                                                    rotation.data_ptr(), di, sp), "gg_mesh_bind_forward_ex")
        E0 = torch.empty(0, device=dev)
        ctx.save_for_backward(v, faces_i32, binding_i32, lx, ls, lr, frames, bc if bc is not None else E0,
                              sr if sr is not None else E0)
        return xyz, scaling, rotation

    @staticmethod
    def backward(ctx, g_xyz, g_scaling, g_rotation):
        lib = _capi.load()
        v, faces_i32, binding_i32, lx, ls, lr, frames, bc, sr = ctx.saved_tensors
        bc = bc if bc.numel() else None
        sr = sr if sr.numel() else None
        dev = v.device
        di = dev.index if dev.index is not None else torch.cuda.current_device()
        V, F, N = v.shape[0], faces_i32.shape[0], lx.shape[0]
        need = ctx.needs_input_grad

        def prep(g):
            if g is None:
                return None
            g = g.float() if g.dtype != torch.float32 else g
            g = g.contiguous()
            return g.clone() if g.data_ptr() % 16 else g

        gx, gs, gr = prep(g_xyz), prep(g_scaling), prep(g_rotation)
        E = lambda *s: torch.empty(*s, dtype=torch.float32, device=dev)
        g_v = E(V, 3) if need[0] else None
        g_lx = E(N, 3) if need[3] else None
        g_ls = E(N, 3) if need[4] else None
        g_lr = E(N, 4) if need[5] else None
        gF = torch.empty(frames.numel(), dtype=torch.uint8, device=dev) if need[0] else None
        p = lambda t: None if t is None else t.data_ptr()
        with torch.cuda.device(dev):
            sp = torch.cuda.current_stream(dev).cuda_stream
            _capi.check(lib.gg_mesh_bind_backward_ex(V, F, N, v.data_ptr(), faces_i32.data_ptr(), binding_i32.data_ptr(),
                                                     lx.data_ptr(), ls.data_ptr(), lr.data_ptr(), p(bc), p(sr),
                                                     frames.data_ptr(), p(gF), p(gx), p(gs), p(gr), p(g_v), p(g_lx),
                                                     p(g_ls), p(g_lr), di, sp), "gg_mesh_bind_backward_ex")
        return g_v, None, None, g_lx, g_ls, g_lr, None, None


def bind_to_mesh(mesh_v, mesh_f, binding, local_xyz, local_scaling, local_rotation, barycentric=None,
                 face_scaling_remembered=None):
    """(world xyz [N,3], world scaling [N,3], world rotation [N,4] wxyz) for Gaussians bound to mesh faces.

    mesh_f [F,3] and binding [N] are index tensors (any integer dtype; pass int32 to avoid a conversion per call).
    local_scaling is the *pre-activation* (log) scale `_scaling`, local_rotation the raw `_rotation`.
    barycentric [N,3] (or the reference's list of three [N] tensors `gs_bc`) switches to AvatarGaussianModel's
    anchor (scene/avatar_gaussian_model.py:140-159); face_scaling_remembered [F] or [F,1] to the frozen face scale
    of remember_scaling() (scene/mesh_gaussian_model.py:98-110)."""
    dev = mesh_v.device
    f32 = _index_tensor(mesh_f, "mesh.f", dev, mesh_v.shape[0])
    b32 = _index_tensor(binding, "binding", dev, f32.shape[0], count=local_xyz.shape[0])
    if isinstance(barycentric, (list, tuple)):
        barycentric = torch.stack([b.reshape(-1) for b in barycentric], dim=1)
    return _MeshBind.apply(mesh_v, f32, b32, local_xyz, local_scaling, local_rotation, barycentric, face_scaling_remembered)


class FusedMeshBinding:
    """Drop-in provider of the world-space properties for an object with MeshGaussianModel's / AvatarGaussianModel's
    attributes (`mesh.v` / `mesh.f` or `mesh_v` / `mesh_f`, `binding`, `_xyz`, `_scaling`, `_rotation`; optionally
    `gs_bc`, `local_xyz`, `face_scaling_remembered`).  world() = (get_xyz, get_scaling, get_rotation);
    world(final=True) uses `local_xyz` like get_final_xyz (scene/avatar_gaussian_model.py:146-148)."""

    def __init__(self, model):
        self.model = model
        self._cache = {}

    def _mesh(self):
        m = self.model
        if hasattr(m, "mesh"):
            return m.mesh.v, m.mesh.f
        return m.mesh_v, m.mesh_f

    def _cached_index(self, key, src, device, upper, count=None):
        """int32 copy of an index tensor, re-made whenever the SOURCE changes: identity, in-place version, storage
        and shape are all part of the tag (the reference re-assigns `binding` on prune / densify -- often to a tensor
        of the same length, scene/mesh_gaussian_model.py:130-208)."""
        tag = (id(src), src._version, src.data_ptr(), tuple(src.shape), str(device))
        hit = self._cache.get(key)
        if hit is None or hit[0] != tag:
            hit = (tag, _index_tensor(src, key, device, upper, count), src)     # src kept alive: id() stays unique
            self._cache[key] = hit
        return hit[1]

    def world(self, final: bool = False):
        m = self.model
        v, f = self._mesh()
        f32 = self._cached_index("mesh.f", f, v.device, v.shape[0])
        b32 = self._cached_index("binding", m.binding, v.device, f32.shape[0], count=m._xyz.shape[0])
        bc = getattr(m, "gs_bc", None)
        if isinstance(bc, (list, tuple)):
            tag = tuple(id(b) for b in bc)
            hit = self._cache.get("gs_bc")
            if hit is None or hit[0] != tag:
                hit = (tag, torch.stack([b.reshape(-1).float() for b in bc], dim=1).contiguous().to(v.device), bc)
                self._cache["gs_bc"] = hit
            bc = hit[1]
        lx = m.local_xyz if final else m._xyz
        if final and lx is None:
            raise RuntimeError("gaussian-garments_b200: world(final=True) needs `local_xyz` (scene/avatar_net.py:82)")
        rem = getattr(m, "face_scaling_remembered", None)
        return _MeshBind.apply(v, f32, b32, lx, m._scaling, m._rotation, bc, rem)
