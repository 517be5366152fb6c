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


class _MeshBind(torch.autograd.Function):
    @staticmethod
    def forward(ctx, mesh_v, faces_i32, binding_i32, local_xyz, local_scaling, local_rotation):
        lib = _capi.load()
        v = _f32c(mesh_v, "mesh_v")
        lx, ls, lr = _f32c(local_xyz, "_xyz"), _f32c(local_scaling, "_scaling"), _f32c(local_rotation, "_rotation")
        dev = v.device
        di = dev.index if dev.index is not None else torch.cuda.current_device()
        V, F, N = v.shape[0], faces_i32.shape[0], lx.shape[0]
        fb = C.c_size_t()
        _capi.check(lib.gg_mesh_bind_workspace_bytes(F, C.byref(fb)), "gg_mesh_bind_workspace_bytes")
        frames = torch.empty(fb.value, dtype=torch.uint8, device=dev)
        xyz = torch.empty(N, 3, device=dev)
        scaling = torch.empty(N, 3, device=dev)
        rotation = torch.empty(N, 4, device=dev)
        with torch.cuda.device(dev):
            sp = torch.cuda.current_stream(dev).cuda_stream
            _capi.check(lib.gg_mesh_bind_forward(V, F, N, v.data_ptr(), faces_i32.data_ptr(), binding_i32.data_ptr(),
                                                 lx.data_ptr(), ls.data_ptr(), lr.data_ptr(), frames.data_ptr(),
                                                 xyz.data_ptr(), scaling.data_ptr(), rotation.data_ptr(), di, sp),
                        "gg_mesh_bind_forward")
        ctx.save_for_backward(v, faces_i32, binding_i32, lx, ls, lr, frames)
        return xyz, scaling, rotation

    @staticmethod
    def backward(ctx, g_xyz, g_scaling, g_rotation):
        lib = _capi.load()
        v, faces_i32, binding_i32, lx, ls, lr, frames = ctx.saved_tensors
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
            _capi.check(lib.gg_mesh_bind_backward(V, F, N, v.data_ptr(), faces_i32.data_ptr(), binding_i32.data_ptr(),
                                                  lx.data_ptr(), ls.data_ptr(), lr.data_ptr(), frames.data_ptr(), p(gF),
                                                  p(gx), p(gs), p(gr), p(g_v), p(g_lx), p(g_ls), p(g_lr), di, sp),
                        "gg_mesh_bind_backward")
        return g_v, None, None, g_lx, g_ls, g_lr


def bind_to_mesh(mesh_v, mesh_f, binding, local_xyz, local_scaling, local_rotation):
    """(world xyz [N,3], world scaling [N,3], world rotation [N,4] wxyz) for Gaussians bound to mesh faces.

    mesh_f [F,3] and binding [N] are index tensors (any integer dtype; pass int32 to avoid a conversion per call).
    local_scaling is the *pre-activation* (log) scale `_scaling`, local_rotation the raw `_rotation`."""
    return _MeshBind.apply(mesh_v, _i32(mesh_f).contiguous(), _i32(binding).contiguous(), local_xyz, local_scaling,
                           local_rotation)


class FusedMeshBinding:
    """Drop-in provider of the three world-space properties for an object with MeshGaussianModel's attributes
    (`mesh.v` / `mesh.f` or `mesh_v` / `mesh_f`, `binding`, `_xyz`, `_scaling`, `_rotation`)."""

    def __init__(self, model):
        self.model = model
        self._f32 = None
        self._b32 = None

    def _mesh(self):
        m = self.model
        if hasattr(m, "mesh"):
            return m.mesh.v, m.mesh.f
        return m.mesh_v, m.mesh_f

    def world(self):
        m = self.model
        v, f = self._mesh()
        if self._f32 is None or self._f32.shape[0] != f.shape[0] or self._f32.device != v.device:
            self._f32 = _i32(f).contiguous().to(v.device)
        if self._b32 is None or self._b32.shape[0] != m.binding.shape[0] or self._b32.device != v.device:
            self._b32 = _i32(m.binding).contiguous().to(v.device)
        return _MeshBind.apply(v, self._f32, self._b32, m._xyz, m._scaling, m._rotation)
