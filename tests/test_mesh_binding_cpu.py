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


def _model(seed=3):
    m = gg.scenes.MeshBoundGaussians(n_faces_around=16, n_along=4, per_face=3, seed=seed)
    g = torch.Generator().manual_seed(seed)
    m.mesh_v = m.mesh_v + 0.01 * torch.randn(m.mesh_v.shape, generator=g)       # break the cylinder's symmetry
    m._rotation = torch.randn(m._rotation.shape, generator=g)                   # not pre-normalised
    return m


def _ptr(t):
    return C.c_void_p(t.data_ptr())


def test_forward_matches_reference_chain(hostlib):
    m = _model()
    m.update_face_coor()
    F, N = m.mesh_f.shape[0], m.binding.shape[0]
    xyz, sc, ro = torch.zeros(N, 3), torch.zeros(N, 3), torch.zeros(N, 4)
    frames = torch.zeros(F, 17)
    hostlib.mb_forward(F, _ptr(m.mesh_v), _ptr(m.mesh_f), N, _ptr(m.binding), _ptr(m._xyz), _ptr(m._scaling),
                       _ptr(m._rotation), _ptr(xyz), _ptr(sc), _ptr(ro), _ptr(frames))
    assert torch.allclose(frames[:, :9].reshape(F, 3, 3), m.face_orien_mat, atol=1e-6)
    assert torch.allclose(frames[:, 9:10], m.face_scaling, atol=1e-7)
    assert torch.allclose(frames[:, 10:13], m.face_center, atol=1e-7)
    assert torch.allclose(frames[:, 13:17], m.face_orien_quat, atol=1e-6)
    assert torch.allclose(xyz, m.get_xyz, atol=1e-6)
    assert torch.allclose(sc, m.get_scaling, atol=1e-7)
    assert torch.allclose(ro, m.get_rotation, atol=1e-6)


def test_backward_matches_autograd_fp64(hostlib):
    m = _model(seed=7)
    F, N, V = m.mesh_f.shape[0], m.binding.shape[0], m.mesh_v.shape[0]
    g = torch.Generator().manual_seed(1)
    g_xyz, g_sc, g_ro = torch.randn(N, 3, generator=g), torch.randn(N, 3, generator=g), torch.randn(N, 4, generator=g)
    # autograd reference in fp64
    d = gg.scenes.MeshBoundGaussians.__new__(gg.scenes.MeshBoundGaussians)
    d.mesh_f, d.binding = m.mesh_f, m.binding
    leaves = {k: getattr(m, k).double().clone().requires_grad_(True) for k in ("mesh_v", "_xyz", "_scaling", "_rotation")}
    for k, v in leaves.items():
        setattr(d, k, v)
    d.update_face_coor()
    loss = (d.get_xyz * g_xyz.double()).sum() + (d.get_scaling * g_sc.double()).sum() + (d.get_rotation * g_ro.double()).sum()
    loss.backward()
    gv, gx, gs, gr = torch.zeros(V, 3), torch.zeros(N, 3), torch.zeros(N, 3), torch.zeros(N, 4)
    hostlib.mb_backward(V, F, _ptr(m.mesh_v), _ptr(m.mesh_f), N, _ptr(m.binding), _ptr(m._xyz), _ptr(m._scaling),
                        _ptr(m._rotation), _ptr(g_xyz), _ptr(g_sc), _ptr(g_ro), _ptr(gv), _ptr(gx), _ptr(gs), _ptr(gr))
    for name, got, ref in (("mesh_v", gv, leaves["mesh_v"].grad), ("_xyz", gx, leaves["_xyz"].grad),
                           ("_scaling", gs, leaves["_scaling"].grad), ("_rotation", gr, leaves["_rotation"].grad)):
        err = h.rel_inf(got, ref.float())
        assert err < 2e-4, f"{name}: {err:.3e}"
