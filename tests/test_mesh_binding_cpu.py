"""CPU tests of the fused mesh-binding transform's math: the product header
gaussian-garments_b200/csrc/mesh_binding_math.h is compiled for the host and compared with autograd through
the restated reference chain (scene/mesh_gaussian_model.py:90-128, utils/graphics_utils.py:118-137)."""
import ctypes as C
import os
import subprocess

import pytest
import torch

import helpers as h

gg = h.gg
HERE = os.path.dirname(os.path.abspath(__file__))


@pytest.fixture(scope="module")
def hostlib(tmp_path_factory):
    out = tmp_path_factory.mktemp("mb") / "libmb_host.so"
    subprocess.check_call(["/usr/bin/g++", "-O2", "-shared", "-fPIC", "-ffp-contract=off", "-o", str(out),
                           os.path.join(HERE, "native", "mesh_math_host.cpp")])
    return C.CDLL(str(out))


from oracle import mesh_chain as mc  # noqa: E402  (row N1's oracle: golden-pinned restatement of the reference chain)


def _model(seed=3, avatar=False, remembered=False, dtype=torch.float32):
    """Local Gaussian parameters on a small perturbed cylinder -> oracle chain object."""
    m = gg.scenes.MeshBoundGaussians(n_faces_around=16, n_along=4, per_face=3, seed=seed)
    g = torch.Generator().manual_seed(seed)
    verts = m.mesh_v + 0.01 * torch.randn(m.mesh_v.shape, generator=g)          # break the cylinder's symmetry
    rot = torch.randn(m._rotation.shape, generator=g)                           # not pre-normalised
    N = m.binding.shape[0]
    bc = None
    if avatar:
        bc = torch.rand(N, 3, generator=g) + 0.05
        bc = bc / bc.sum(1, keepdim=True)
    chain = mc.MeshChain(verts.to(dtype), m.mesh_f, m.binding, m._xyz.to(dtype), m._scaling.to(dtype), rot.to(dtype),
                         gs_bc=None if bc is None else bc.to(dtype))
    if remembered:
        chain.update_face_coor()
        chain.face_scaling_remembered = (chain.face_scaling * (0.8 + 0.4 * torch.rand(chain.face_scaling.shape, generator=g).to(dtype))).detach()
    return chain


def _ptr(t):
    return C.c_void_p(0) if t is None else C.c_void_p(t.data_ptr())


def test_oracle_chain_is_pinned_by_the_reference_compute_face_orientation():
    """tests/golden/mesh.npz was produced by /root/reference/utils/graphics_utils.py:118-137 itself."""
    import numpy as np
    z = np.load(os.path.join(HERE, "golden", "mesh.npz"))
    verts, faces = torch.tensor(z["verts"]), torch.tensor(z["faces"])
    o, s = mc.compute_face_orientation(verts, faces)
    assert torch.equal(o, torch.tensor(z["orientation"])) and torch.equal(s, torch.tensor(z["scale"]))
    o64, s64 = mc.compute_face_orientation(verts.double(), faces)
    assert torch.equal(o64, torch.tensor(z["orientation64"])) and torch.equal(s64, torch.tensor(z["scale64"]))
    # the restated roma helpers are pinned by the rotation they must reproduce (non-degenerate faces)
    q = torch.cat([mc.rotmat_to_unitquat_xyzw(o64)[:, 3:4], mc.rotmat_to_unitquat_xyzw(o64)[:, :3]], dim=1)
    ok = torch.linalg.det(o64).abs() > 0.5
    assert float((mc.quat_to_rotmat_wxyz(q)[ok] - o64[ok]).abs().max()) < 1e-12
    # and the product's synthetic-scene generator agrees with the pinned oracle
    o2, s2 = gg.scenes.face_orientation(verts, faces)
    assert torch.equal(o2, o) and torch.equal(s2, s)


@pytest.mark.parametrize("avatar,remembered", [(False, False), (True, False), (False, True), (True, True)])
def test_forward_matches_reference_chain(hostlib, avatar, remembered):
    m = _model(avatar=avatar, remembered=remembered)
    m.update_face_coor()
    F, N = m.mesh_f.shape[0], m.binding.shape[0]
    xyz, sc, ro = torch.zeros(N, 3), torch.zeros(N, 3), torch.zeros(N, 4)
    frames = torch.zeros(F, 17)
    rem = None if m.face_scaling_remembered is None else m.face_scaling_remembered.reshape(-1).contiguous()
    hostlib.mb_forward(F, _ptr(m.mesh_v), _ptr(m.mesh_f), N, _ptr(m.binding), _ptr(m._xyz), _ptr(m._scaling),
                       _ptr(m._rotation), _ptr(xyz), _ptr(sc), _ptr(ro), _ptr(frames), _ptr(m.gs_bc), _ptr(rem))
    assert torch.allclose(frames[:, :9].reshape(F, 3, 3), m.face_orien_mat, atol=1e-6)
    assert torch.allclose(frames[:, 9:10], m.face_scaling, atol=1e-7)
    assert torch.allclose(frames[:, 10:13], m.face_center, atol=1e-7)
    assert torch.allclose(frames[:, 13:17], m.face_orien_quat, atol=1e-6)
    assert torch.allclose(xyz, m.get_xyz, atol=1e-6)
    assert torch.allclose(sc, m.get_scaling, atol=1e-7)
    assert torch.allclose(ro, m.get_rotation, atol=1e-6)


@pytest.mark.parametrize("avatar,remembered", [(False, False), (True, False), (False, True), (True, True)])
def test_backward_matches_autograd_fp64(hostlib, avatar, remembered):
    m = _model(seed=7, avatar=avatar, remembered=remembered)
    F, N, V = m.mesh_f.shape[0], m.binding.shape[0], m.mesh_v.shape[0]
    g = torch.Generator().manual_seed(1)
    g_xyz, g_sc, g_ro = torch.randn(N, 3, generator=g), torch.randn(N, 3, generator=g), torch.randn(N, 4, generator=g)
    # autograd reference in fp64
    leaves = {k: getattr(m, k).double().clone().requires_grad_(True) for k in ("mesh_v", "_xyz", "_scaling", "_rotation")}
    d = mc.MeshChain(leaves["mesh_v"], m.mesh_f, m.binding, leaves["_xyz"], leaves["_scaling"], leaves["_rotation"],
                     gs_bc=None if m.gs_bc is None else m.gs_bc.double())
    d.update_face_coor()
    if remembered:
        d.face_scaling_remembered = m.face_scaling_remembered.double()
    loss = (d.get_xyz * g_xyz.double()).sum() + (d.get_scaling * g_sc.double()).sum() + (d.get_rotation * g_ro.double()).sum()
    loss.backward()
    gv, gx, gs, gr = torch.zeros(V, 3), torch.zeros(N, 3), torch.zeros(N, 3), torch.zeros(N, 4)
    rem = None if m.face_scaling_remembered is None else m.face_scaling_remembered.reshape(-1).contiguous()
    hostlib.mb_backward(V, F, _ptr(m.mesh_v), _ptr(m.mesh_f), N, _ptr(m.binding), _ptr(m._xyz), _ptr(m._scaling),
                        _ptr(m._rotation), _ptr(g_xyz), _ptr(g_sc), _ptr(g_ro), _ptr(gv), _ptr(gx), _ptr(gs), _ptr(gr),
                        _ptr(m.gs_bc), _ptr(rem))
    for name, got, ref in (("mesh_v", gv, leaves["mesh_v"].grad), ("_xyz", gx, leaves["_xyz"].grad),
                           ("_scaling", gs, leaves["_scaling"].grad), ("_rotation", gr, leaves["_rotation"].grad)):
        err = h.rel_inf(got, ref.float())
        assert err < 2e-4, f"{name}: {err:.3e}"
