"""CPU tests of row N3's oracle (oracle/raycast_oracle.c): known answers for the first-hit ray cast that
tests/test_gpu_parity.py compares the CUDA visibility kernels with (bit-equal primitive ids)."""
import math

import torch

import helpers as h

gg = h.gg
co = h.c_oracle


def test_known_answers_two_parallel_quads():
    """Two unit quads facing the origin at z = 2 (faces 0,1) and z = 3 (faces 2,3): rays through the near quad hit it
    at t = distance to z = 2; rays that pass beside it hit the far quad; rays outside both hit nothing."""
    def quad(z, s):
        return torch.tensor([[-s, -s, z], [s, -s, z], [s, s, z], [-s, s, z]], dtype=torch.float32)
    verts = torch.cat([quad(2.0, 0.5), quad(3.0, 2.0)])
    faces = torch.tensor([[0, 1, 2], [0, 2, 3], [4, 5, 6], [4, 6, 7]])
    origin = torch.zeros(3)
    targets = torch.tensor([[0.2, -0.1, 3.0],     # through the near quad (lower-right triangle 0) -> 0
                            [-0.2, 0.1, 3.0],     # near quad, triangle 1
                            [1.5, -0.5, 3.0],     # misses the near quad, lands on the far one (triangle 2: x > y)
                            [-1.5, 0.5, 3.0],     # far quad, triangle 3
                            [9.0, 9.0, 3.0],      # outside everything
                            [0.1, 0.05, -1.0]])   # pointing away
    prim, t = co.cast_rays_from_point(verts, faces, targets, origin)
    assert prim.tolist() == [0, 1, 2, 3, -1, -1]
    d0 = targets[0] / targets[0].norm()
    assert abs(float(t[0]) - 2.0 / float(d0[2])) < 1e-5
    assert math.isinf(float(t[4])) and math.isinf(float(t[5]))


def test_avatar_semantics_on_a_cylinder():
    """get_visible_mask semantics (scene/avatar_gaussian_model.py:227-263): points on the camera-facing half of a
    closed-in-view cylinder see their own face first, points on the far half are occluded by the near wall."""
    verts, faces = gg.scenes.cylinder_mesh(n_around=48, n_along=8, wrinkle_amp=0.0)
    centers = verts[faces].mean(1)
    cam = torch.tensor([0.0, 0.6, 3.0])
    prim, _ = co.cast_rays_from_point(verts, faces, centers, cam)
    vis = prim == torch.arange(faces.shape[0], dtype=torch.int32)
    front = centers[:, 2] > 0.05
    back = centers[:, 2] < -0.05
    assert bool(vis[front].all()) and not bool(vis[back].any())
    assert bool((prim >= 0).all())
