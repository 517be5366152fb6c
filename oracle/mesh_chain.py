"""TEST INFRASTRUCTURE -- CPU/PyTorch restatement of the reference's mesh-binding chain (row N1's oracle).

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg may import this; the product path
(gaussian-garments_b200/csrc/mesh_binding.cu behind gg_mesh_bind_*_ex) never does.

Follows, line by line:
  utils/graphics_utils.py:97-104   length / safe_normalize (clamp 1e-20 under the sqrt)
  utils/graphics_utils.py:118-137  compute_face_orientation(verts, faces, return_scale=True)
  scene/mesh_gaussian_model.py:90-95     update_face_coor (centre, frame, scale, quaternion wxyz)
  scene/mesh_gaussian_model.py:98-110    remember_scaling / get_scaling (face_scaling_remembered branch)
  scene/mesh_gaussian_model.py:118-128   get_rotation, get_xyz (face-centre anchor)
  scene/avatar_gaussian_model.py:140-159 get_xyz / get_final_xyz / get_barycentric_3d (barycentric anchor)
`roma` (rotmat_to_unitquat, quat_product; xyzw) is not in the image: both are restated from their published
definitions.  PINNING: compute_face_orientation is importable here, so tests/golden/mesh.npz holds its outputs
(tests/golden/make_golden.py) and tests/test_mesh_binding_cpu.py checks this file against them; the quaternion
helpers are pinned by the rotation they must reproduce (R(q) == face_orien_mat).  Gradients come from autograd
(any dtype; tests use float64), independent of the hand-derived CUDA backward.
"""
import torch


def length(x, eps=1e-20):
    return torch.sqrt(torch.clamp((x * x).sum(-1, keepdim=True), min=eps))


def safe_normalize(x, eps=1e-20):
    return x / length(x, eps)


def compute_face_orientation(verts, faces):
    i0, i1, i2 = faces[..., 0].long(), faces[..., 1].long(), faces[..., 2].long()
    v0, v1, v2 = verts[..., i0, :], verts[..., i1, :], verts[..., i2, :]
    a0 = safe_normalize(v1 - v0)
    a1 = safe_normalize(torch.cross(a0, v2 - v0, dim=-1))
    a2 = -safe_normalize(torch.cross(a1, a0, dim=-1))
    orientation = torch.cat([a0[..., None], a1[..., None], a2[..., None]], dim=-1)
    s0 = length(v1 - v0)
    s1 = (a2 * (v2 - v0)).sum(-1, keepdim=True).abs()
    return orientation, (s0 + s1) / 2


def rotmat_to_unitquat_xyzw(R):
    """Largest-of-(m00, m11, m22, trace) branch selection, unit xyzw quaternion (roma.rotmat_to_unitquat)."""
    m = lambda i, j: R[:, i, j]
    tr = m(0, 0) + m(1, 1) + m(2, 2)
    choice = torch.stack([m(0, 0), m(1, 1), m(2, 2), tr], dim=1).argmax(dim=1)
    cand = [torch.stack([1 - tr + 2 * m(0, 0), m(1, 0) + m(0, 1), m(2, 0) + m(0, 2), m(2, 1) - m(1, 2)], 1),
            torch.stack([m(1, 0) + m(0, 1), 1 - tr + 2 * m(1, 1), m(2, 1) + m(1, 2), m(0, 2) - m(2, 0)], 1),
            torch.stack([m(2, 0) + m(0, 2), m(2, 1) + m(1, 2), 1 - tr + 2 * m(2, 2), m(1, 0) - m(0, 1)], 1),
            torch.stack([m(2, 1) - m(1, 2), m(0, 2) - m(2, 0), m(1, 0) - m(0, 1), 1 + tr], 1)]
    q = cand[3]
    for k in range(3):
        q = torch.where((choice == k)[:, None], cand[k], q)
    return q / q.norm(dim=1, keepdim=True)


def quat_product_xyzw(p, q):
    px, py, pz, pw = p.unbind(-1)
    qx, qy, qz, qw = q.unbind(-1)
    return torch.stack([pw * qx + px * qw + py * qz - pz * qy, pw * qy - px * qz + py * qw + pz * qx,
                        pw * qz + px * qy - py * qx + pz * qw, pw * qw - px * qx - py * qy - pz * qz], dim=-1)


def quat_to_rotmat_wxyz(q):
    """utils/general_utils.py:88-110 build_rotation (normalises first) -- used to pin the quaternion helpers."""
    q = q / q.norm(dim=1, keepdim=True)
    r, x, y, z = q.unbind(-1)
    return torch.stack([1 - 2 * (y * y + z * z), 2 * (x * y - r * z), 2 * (x * z + r * y),
                        2 * (x * y + r * z), 1 - 2 * (x * x + z * z), 2 * (y * z - r * x),
                        2 * (x * z - r * y), 2 * (y * z + r * x), 1 - 2 * (x * x + y * y)], dim=-1).reshape(-1, 3, 3)


_xyzw_to_wxyz = lambda q: torch.cat([q[..., 3:4], q[..., 0:3]], dim=-1)
_wxyz_to_xyzw = lambda q: torch.cat([q[..., 1:4], q[..., 0:1]], dim=-1)


class MeshChain:
    """Attribute-compatible with the part of MeshGaussianModel / AvatarGaussianModel that produces the rasterizer's
    inputs.  All tensors are taken as given (any dtype / requires_grad)."""

    def __init__(self, mesh_v, mesh_f, binding, _xyz, _scaling, _rotation, gs_bc=None, local_xyz=None):
        self.mesh_v, self.mesh_f, self.binding = mesh_v, mesh_f.long(), binding.long()
        self._xyz, self._scaling, self._rotation = _xyz, _scaling, _rotation
        self.gs_bc = gs_bc                      # [N,3] or None
        self.local_xyz = local_xyz
        self.face_scaling_remembered = None

    def update_face_coor(self):                                     # mesh_gaussian_model.py:90-95
        self.face_center = self.mesh_v[self.mesh_f].mean(1)
        self.face_orien_mat, self.face_scaling = compute_face_orientation(self.mesh_v, self.mesh_f)
        self.face_orien_quat = _xyzw_to_wxyz(rotmat_to_unitquat_xyzw(self.face_orien_mat))

    def remember_scaling(self):                                     # :98-103
        self.face_scaling_remembered = self.face_scaling.detach()

    @property
    def get_scaling(self):                                          # :105-116
        fs = self.face_scaling_remembered if self.face_scaling_remembered is not None else self.face_scaling
        return torch.exp(self._scaling) * fs[self.binding]

    @property
    def get_rotation(self):                                         # :118-122
        norm = torch.nn.functional.normalize
        rot = norm(self._rotation)
        fq = norm(self.face_orien_quat[self.binding])
        return norm(_xyzw_to_wxyz(quat_product_xyzw(_wxyz_to_xyzw(fq), _wxyz_to_xyzw(rot))))

    def get_barycentric_3d(self):                                   # avatar_gaussian_model.py:151-154
        tri = self.mesh_v[self.mesh_f][self.binding]
        a, b, c = self.gs_bc.unbind(-1)
        return a[:, None] * tri[:, 0] + b[:, None] * tri[:, 1] + c[:, None] * tri[:, 2]

    def _world(self, local):
        xyz = torch.bmm(self.face_orien_mat[self.binding], local[..., None]).squeeze(-1)
        anchor = self.get_barycentric_3d() if self.gs_bc is not None else self.face_center[self.binding]
        return xyz * self.face_scaling[self.binding] + anchor

    @property
    def get_xyz(self):                                              # mesh :124-128 / avatar :140-143
        return self._world(self._xyz)

    @property
    def get_final_xyz(self):                                        # avatar :145-148
        return self._world(self.local_xyz)
