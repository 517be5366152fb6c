"""Synthetic Gaussian states for the BASELINE.json configs.

Nothing here reads /root/reference at run time.  The mesh-bound state restates
the arithmetic of the reference's `MeshGaussianModel`:
  scene/mesh_gaussian_model.py:90-95    update_face_coor  (face centre, frame, scale, quat)
  scene/mesh_gaussian_model.py:105-116  get_scaling   = exp(_scaling) * face_scaling[binding]
  scene/mesh_gaussian_model.py:118-122  get_rotation  = normalize(q_face (x) normalize(_rotation))
  scene/mesh_gaussian_model.py:124-128  get_xyz       = R_face[binding] @ _xyz * s_face + c_face
  utils/graphics_utils.py:118-137       compute_face_orientation
  scene/gaussian_model.py:33-41,107-116 activations, get_features, get_opacity
`roma` (rotmat_to_unitquat / quat_product, xyzw) is absent from the image, so the
two quaternion helpers are restated here from their published definitions.
"""

from __future__ import annotations

import math
from dataclasses import dataclass
from typing import Optional

import torch

from .cameras import PinholeCamera, make_camera, ring_cameras, fov2focal

SEED = 31359  # the reference's own seed, s3_appearance.py:89-90


# --------------------------------------------------------------------------- #
# small math helpers (restated, device agnostic)
# --------------------------------------------------------------------------- #
def _safe_normalize(x, eps=1e-20):
    # utils/graphics_utils.py:100-104
    return x / torch.sqrt(torch.clamp((x * x).sum(-1, keepdim=True), min=eps))


def face_orientation(verts: torch.Tensor, faces: torch.Tensor):
    """utils/graphics_utils.py:118-137 (return_scale=True)."""
    v0, v1, v2 = verts[faces[:, 0]], verts[faces[:, 1]], verts[faces[:, 2]]
    a0 = _safe_normalize(v1 - v0)
    a1 = _safe_normalize(torch.cross(a0, v2 - v0, dim=-1))
    a2 = -_safe_normalize(torch.cross(a1, a0, dim=-1))
    orientation = torch.stack([a0, a1, a2], dim=-1)          # columns a0 | a1 | a2
    s0 = torch.sqrt(torch.clamp(((v1 - v0) ** 2).sum(-1, keepdim=True), min=1e-20))
    s1 = (a2 * (v2 - v0)).sum(-1, keepdim=True).abs()
    return orientation, (s0 + s1) / 2


def rotmat_to_unitquat_xyzw(R: torch.Tensor) -> torch.Tensor:
    """Rotation matrix -> unit quaternion (x,y,z,w); largest-component branch selection (same algorithm as
    roma.rotmat_to_unitquat, written with torch.where so that it never synchronises with the host)."""
    m00, m01, m02 = R[:, 0, 0], R[:, 0, 1], R[:, 0, 2]
    m10, m11, m12 = R[:, 1, 0], R[:, 1, 1], R[:, 1, 2]
    m20, m21, m22 = R[:, 2, 0], R[:, 2, 1], R[:, 2, 2]
    tr = m00 + m11 + m22
    choice = torch.stack([m00, m11, m22, tr], dim=1).argmax(dim=1)
    q3 = torch.stack([m21 - m12, m02 - m20, m10 - m01, 1 + tr], dim=1)                 # trace largest
    q0 = torch.stack([1 - tr + 2 * m00, m10 + m01, m20 + m02, m21 - m12], dim=1)
    q1 = torch.stack([m10 + m01, 1 - tr + 2 * m11, m21 + m12, m02 - m20], dim=1)
    q2 = torch.stack([m20 + m02, m21 + m12, 1 - tr + 2 * m22, m10 - m01], dim=1)
    c = choice[:, None]
    q = torch.where(c == 3, q3, torch.where(c == 0, q0, torch.where(c == 1, q1, q2)))
    return q / q.norm(dim=1, keepdim=True)


def quat_product_xyzw(p: torch.Tensor, q: torch.Tensor) -> torch.Tensor:
    """Hamilton product p*q, both (x,y,z,w)."""
    px, py, pz, pw = p.unbind(-1)
    qx, qy, qz, qw = q.unbind(-1)
    return torch.stack([
        pw * qx + px * qw + py * qz - pz * qy,
        pw * qy - px * qz + py * qw + pz * qx,
        pw * qz + px * qy - py * qx + pz * qw,
        pw * qw - px * qx - py * qy - pz * qz,
    ], dim=-1)


def _xyzw_to_wxyz(q):
    return torch.cat([q[..., 3:4], q[..., 0:3]], dim=-1)


def _wxyz_to_xyzw(q):
    return torch.cat([q[..., 1:4], q[..., 0:1]], dim=-1)


# --------------------------------------------------------------------------- #
# states
# --------------------------------------------------------------------------- #
@dataclass
class GaussianState:
    """World-space, post-activation tensors exactly as the facade hands them to the rasterizer
    (gaussian_renderer/__init__.py:56-87,103-111)."""

    means3D: torch.Tensor      # [N,3]
    scales: torch.Tensor       # [N,3]   post-exp (x face scale)
    rotations: torch.Tensor    # [N,4]   wxyz, normalised
    opacities: torch.Tensor    # [N,1]   post-sigmoid
    shs: torch.Tensor          # [N,M,3] coefficient-major, RGB-minor
    sh_degree: int             # active degree
    bg: torch.Tensor           # [3]

    @property
    def N(self):
        return self.means3D.shape[0]

    def to(self, device):
        return GaussianState(*(t.to(device) if torch.is_tensor(t) else t for t in (
            self.means3D, self.scales, self.rotations, self.opacities, self.shs, self.sh_degree, self.bg)))

    def detach_clone(self, requires_grad=False):
        ts = [self.means3D, self.scales, self.rotations, self.opacities, self.shs]
        ts = [t.detach().clone().requires_grad_(requires_grad) for t in ts]
        return GaussianState(*ts, self.sh_degree, self.bg.detach().clone())

    def nbytes(self):
        return sum(t.numel() * t.element_size() for t in
                   (self.means3D, self.scales, self.rotations, self.opacities, self.shs))


def random_cloud(N=10_000, sh_degree=3, seed=SEED, extent=1.0, log_scale=(-4.5, -2.5),
                 max_sh_degree=3, dtype=torch.float32) -> GaussianState:
    """SURVEY.md 8d cfg1 state (and, scaled, cfg5)."""
    g = torch.Generator().manual_seed(seed)
    M = (max_sh_degree + 1) ** 2
    means = (torch.rand(N, 3, generator=g) * 2 - 1) * extent
    scales = torch.exp(torch.rand(N, 3, generator=g) * (log_scale[1] - log_scale[0]) + log_scale[0])
    rots = torch.nn.functional.normalize(torch.randn(N, 4, generator=g))
    opac = torch.sigmoid(torch.randn(N, 1, generator=g) * 1.5)
    shs = torch.randn(N, M, 3, generator=g) * 0.15
    shs[:, 0, :] = torch.randn(N, 3, generator=g)
    bg = torch.rand(3, generator=g)
    return GaussianState(means.to(dtype), scales.to(dtype), rots.to(dtype), opac.to(dtype),
                         shs.to(dtype), sh_degree, bg.to(dtype))


def cfg1_camera(width=512, height=512, fov_deg=50.0) -> PinholeCamera:
    """SURVEY.md 8d cfg1 camera: R=I, T=(0,0,4), off-centre principal point."""
    import numpy as np
    f = fov2focal(math.radians(fov_deg), width)
    return make_camera(np.eye(3), np.array([0.0, 0.0, 4.0]), f, f,
                       width / 2 + 3.5, height / 2 - 2.25, width, height)


def cylinder_mesh(n_around=250, n_along=100, radius=0.35, height=1.2, wrinkle_amp=0.01):
    """Synthetic template_uv.obj stand-in: open cylinder, n_around*n_along quads -> 2x triangles."""
    i = torch.arange(n_around, dtype=torch.float64)
    j = torch.arange(n_along + 1, dtype=torch.float64)
    theta = (2 * math.pi * i / n_around)[None, :].expand(n_along + 1, n_around)
    y = (height * j / n_along)[:, None].expand(n_along + 1, n_around)
    r = radius + wrinkle_amp * torch.sin(7 * theta) * torch.sin(2 * math.pi * 3 * y / height)
    verts = torch.stack([r * torch.sin(theta), y, r * torch.cos(theta)], dim=-1).reshape(-1, 3)
    jj, ii = torch.meshgrid(torch.arange(n_along), torch.arange(n_around), indexing="ij")
    v00 = jj * n_around + ii
    v01 = jj * n_around + (ii + 1) % n_around
    v10 = (jj + 1) * n_around + ii
    v11 = (jj + 1) * n_around + (ii + 1) % n_around
    faces = torch.cat([torch.stack([v00, v01, v11], -1).reshape(-1, 3),
                       torch.stack([v00, v11, v10], -1).reshape(-1, 3)], dim=0)
    return verts.float(), faces.long()


class MeshBoundGaussians:
    """Local (face-frame) Gaussian parameters + mesh; produces the world-space state.

    Mirrors the part of `MeshGaussianModel` that feeds the rasterizer.  Kept in plain
    torch so gradients chain to `mesh_v` the way stage 2 needs them
    (scene/mesh_gaussian_model.py:366-371)."""

    def __init__(self, n_faces_around=250, n_along=100, per_face=6, sh_degree=3, seed=SEED,
                 max_sh_degree=3):
        g = torch.Generator().manual_seed(seed)
        self.mesh_v, self.mesh_f = cylinder_mesh(n_faces_around, n_along)
        F = self.mesh_f.shape[0]
        self.binding = torch.arange(F).repeat_interleave(per_face)
        N = self.binding.shape[0]
        M = (max_sh_degree + 1) ** 2
        xyz = torch.randn(N, 3, generator=g)
        # face frame columns: a0 (edge), a1 (normal), a2 (in-plane) -> in-plane axes 0 and 2
        self._xyz = xyz * torch.tensor([0.25, 0.02, 0.25])
        self._scaling = torch.log(torch.rand(N, 3, generator=g) * (0.6 - 0.15) + 0.15)
        self._rotation = torch.nn.functional.normalize(
            torch.tensor([1.0, 0, 0, 0]) + 0.1 * torch.randn(N, 4, generator=g))
        op = (0.85 + 0.1 * torch.randn(N, 1, generator=g)).clamp(0.05, 0.995)
        self._opacity = torch.log(op / (1 - op))
        shs = torch.randn(N, M, 3, generator=g) * 0.15
        shs[:, 0, :] = torch.randn(N, 3, generator=g)
        self._features = shs
        self.active_sh_degree = sh_degree
        self.max_sh_degree = max_sh_degree
        self.bg = torch.rand(3, generator=g)

    def to(self, device):
        for k in ("mesh_v", "mesh_f", "binding", "_xyz", "_scaling", "_rotation", "_opacity",
                  "_features", "bg"):
            setattr(self, k, getattr(self, k).to(device))
        return self

    # scene/mesh_gaussian_model.py:90-95
    def update_face_coor(self):
        self.face_center = self.mesh_v[self.mesh_f].mean(1)
        self.face_orien_mat, self.face_scaling = face_orientation(self.mesh_v, self.mesh_f)
        self.face_orien_quat = _xyzw_to_wxyz(rotmat_to_unitquat_xyzw(self.face_orien_mat))

    @property
    def get_scaling(self):
        return torch.exp(self._scaling) * self.face_scaling[self.binding]

    @property
    def get_rotation(self):
        norm = torch.nn.functional.normalize
        rot = norm(self._rotation)
        fq = norm(self.face_orien_quat[self.binding])
        world = _xyzw_to_wxyz(quat_product_xyzw(_wxyz_to_xyzw(fq), _wxyz_to_xyzw(rot)))
        return norm(world)

    @property
    def get_xyz(self):
        xyz = torch.bmm(self.face_orien_mat[self.binding], self._xyz[..., None]).squeeze(-1)
        return xyz * self.face_scaling[self.binding] + self.face_center[self.binding]

    @property
    def get_opacity(self):
        return torch.sigmoid(self._opacity)

    @property
    def get_features(self):
        return self._features

    def world_state(self) -> GaussianState:
        self.update_face_coor()
        return GaussianState(self.get_xyz.contiguous(), self.get_scaling.contiguous(),
                             self.get_rotation.contiguous(), self.get_opacity.contiguous(),
                             self.get_features.contiguous(), self.active_sh_degree, self.bg)


def mesh_bound_state(n_gaussians=300_000, sh_degree=3, seed=SEED, max_sh_degree=3) -> GaussianState:
    """cfg2: 50 000-face cylinder x 6 = 300 000; cfg4: 25 000 faces x 6 = 150 000."""
    per_face = 6
    faces = n_gaussians // per_face
    n_around = 250
    n_along = max(1, faces // (2 * n_around))
    with torch.no_grad():
        model = MeshBoundGaussians(n_around, n_along, per_face, sh_degree, seed, max_sh_degree)
        st = model.world_state()
    return st


def cfg2_cameras(n=8, width=1920, height=1080):
    return ring_cameras(n, radius=3.0, height_y=0.6, target=(0.0, 0.6, 0.0), fovy_deg=30.0,
                        width=width, height=height, seed=SEED)


def stress_cloud(N=2_000_000, seed=SEED, sh_degree=3) -> GaussianState:
    """cfg5: random cloud inside the cylinder's bounding volume, scales x0.5."""
    st = random_cloud(N, sh_degree, seed, extent=1.0, log_scale=(-4.5 + math.log(0.5), -2.5 + math.log(0.5)))
    st.means3D = st.means3D * torch.tensor([0.35, 0.6, 0.35]) + torch.tensor([0.0, 0.6, 0.0])
    return st
