"""Runs the reference's OWN densification code -- MeshGaussianModel.densify_and_prune / prune_points / reset_opacity /
add_densification_stats and the Adam-state surgery of /root/reference/scene/gaussian_model.py:257-406 and
/root/reference/scene/mesh_gaussian_model.py:130-203, unmodified -- on a small CPU model and writes the state before
and after to tests/golden/densify.npz (authoring container only; needs /root/reference).

    python tests/golden/make_densify_golden.py

tests/test_densify_cpu.py replays the same steps through gaussian-garments_b200/densify.py and demands identical
tensors (parameters, Adam moments, binding, binding_counter, statistics).

Shims used only here: the model classes import packages this image lacks (trimesh, smplx, plyfile, simple_knn, open3d,
roma, lbs, munch ...) -> stub modules; `device="cuda"` allocations are redirected to the CPU; the model object is made
with __new__ (its __init__ reads files) and given exactly the attributes the methods touch.
"""
import os
import sys
import types

import numpy as np
import torch

REF = "/root/reference"
HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))


def import_reference_models():
    sys.path.insert(0, REF)

    def stub(name, **attrs):
        m = types.ModuleType(name)
        m.__dict__.update(attrs)
        sys.modules[name] = m
        return m

    anyf = lambda *a, **k: None
    stub("trimesh"); stub("smplx"); stub("open3d"); stub("lbs", prepare_lbs=anyf)
    stub("plyfile", PlyData=object, PlyElement=object)
    stub("simple_knn"); stub("simple_knn._C", distCUDA2=anyf)
    stub("roma", rotmat_to_unitquat=anyf, quat_xyzw_to_wxyz=anyf, quat_wxyz_to_xyzw=anyf, quat_product=anyf, unitquat_to_rotmat=anyf)
    stub("munch", munchify=lambda d: types.SimpleNamespace(**d), Munch=dict)
    stub("utils.defaults", DEFAULTS=types.SimpleNamespace(output_root="", stage1="", stage2="", aux_root=""))
    stub("utils.io_utils", read_obj=anyf, fetchPly=anyf, storePly=anyf, write_obj=anyf)
    stub("scene.mesh_model", MeshModel=object)
    scene_pkg = types.ModuleType("scene")
    scene_pkg.__path__ = [os.path.join(REF, "scene")]
    sys.modules["scene"] = scene_pkg
    cams = types.ModuleType("scene.cameras")
    cams.Camera = object
    sys.modules["scene.cameras"] = cams
    for fn in ("zeros", "ones", "zeros_like", "ones_like", "tensor"):
        orig = getattr(torch, fn)

        def cpu(*a, _orig=orig, **k):
            if k.get("device") == "cuda":
                k.pop("device")
            return _orig(*a, **k)

        setattr(torch, fn, cpu)
    torch.cuda.empty_cache = lambda: None
    from scene.mesh_gaussian_model import MeshGaussianModel
    return MeshGaussianModel


def make_model(cls, seed=3):
    sys.path.insert(0, ROOT)
    import importlib
    dgr = importlib.import_module("diff_gaussian_rasterization_depth_alpha")  # registers gaussian_garments_b200
    import gaussian_garments_b200 as gg
    from oracle import mesh_chain as mc
    src = gg.scenes.MeshBoundGaussians(n_faces_around=10, n_along=3, per_face=3, seed=seed)     # 60 faces, 180 Gaussians
    g = torch.Generator().manual_seed(seed)
    m = cls.__new__(cls)
    m.active_sh_degree, m.max_sh_degree = 3, 3
    m.setup_functions()
    P = torch.nn.Parameter
    m._xyz = P(src._xyz.clone())
    m._features_dc = P(src._features[:, :1].clone())
    m._features_rest = P(src._features[:, 1:].clone())
    m._opacity = P((src._opacity + torch.randn(src._opacity.shape, generator=g) * 2.0).clone())
    m._scaling = P((src._scaling + torch.randn(src._scaling.shape, generator=g) * 0.8).clone())
    m._rotation = P(src._rotation.clone())
    m.mesh = types.SimpleNamespace(v=P(src.mesh_v.clone()), f=src.mesh_f.clone(), valid_faces=None)
    m.binding = src.binding.clone()
    m.binding_counter = torch.zeros(src.mesh_f.shape[0], dtype=torch.int32)
    m.binding_counter.scatter_add_(0, m.binding, torch.ones_like(m.binding, dtype=torch.int32))
    chain = mc.MeshChain(m.mesh.v.detach(), m.mesh.f, m.binding, m._xyz, m._scaling, m._rotation)
    chain.update_face_coor()
    m.face_center, m.face_orien_mat, m.face_scaling, m.face_orien_quat = (chain.face_center, chain.face_orien_mat,
                                                                          chain.face_scaling, chain.face_orien_quat)
    m.percent_dense = 0.01
    N = m._xyz.shape[0]
    m.xyz_gradient_accum, m.denom, m.max_radii2D = torch.zeros(N, 1), torch.zeros(N, 1), torch.zeros(N)
    groups = [{"params": [m._xyz], "lr": 1e-3, "name": "xyz"}, {"params": [m._features_dc], "lr": 2e-3, "name": "f_dc"},
              {"params": [m._features_rest], "lr": 1e-4, "name": "f_rest"}, {"params": [m._opacity], "lr": 5e-2, "name": "opacity"},
              {"params": [m._scaling], "lr": 5e-3, "name": "scaling"}, {"params": [m._rotation], "lr": 1e-3, "name": "rotation"},
              {"params": [m.mesh.v], "lr": 1e-4, "name": "vertex"}]
    m.optimizer = torch.optim.Adam(groups, lr=0.0, eps=1e-15)
    # one real Adam step so that every group has non-trivial moments
    for grp in groups:
        p = grp["params"][0]
        p.grad = torch.randn(p.shape, generator=g) * 0.1
    m.optimizer.step()
    m.optimizer.zero_grad(set_to_none=True)
    # quaternion helpers the reference takes from roma (get_rotation is not touched by densification, get_xyz is)
    return m, g


def snapshot(m, tag, out):
    names = {"xyz": "_xyz", "f_dc": "_features_dc", "f_rest": "_features_rest", "opacity": "_opacity",
             "scaling": "_scaling", "rotation": "_rotation"}
    for k, a in names.items():
        p = getattr(m, a)
        out[f"{tag}__{k}"] = p.detach().numpy().copy()
        st = m.optimizer.state.get(p, None)
        if st is not None and "exp_avg" in st:
            out[f"{tag}__{k}__exp_avg"] = st["exp_avg"].numpy().copy()
            out[f"{tag}__{k}__exp_avg_sq"] = st["exp_avg_sq"].numpy().copy()
    out[f"{tag}__binding"] = m.binding.numpy().copy()
    out[f"{tag}__binding_counter"] = m.binding_counter.numpy().copy()
    out[f"{tag}__xyz_gradient_accum"] = m.xyz_gradient_accum.numpy().copy()
    out[f"{tag}__denom"] = m.denom.numpy().copy()
    out[f"{tag}__max_radii2D"] = m.max_radii2D.numpy().copy()
    out[f"{tag}__vertex"] = m.mesh.v.detach().numpy().copy()
    vst = m.optimizer.state.get(m.mesh.v, None)
    out[f"{tag}__vertex__exp_avg"] = vst["exp_avg"].numpy().copy()


def main():
    cls = import_reference_models()
    m, g = make_model(cls)
    out = {}
    out["mesh_f"] = m.mesh.f.numpy().copy()
    out["face_scaling"] = m.face_scaling.numpy().copy()
    out["face_center"] = m.face_center.numpy().copy()
    out["face_orien_mat"] = m.face_orien_mat.numpy().copy()
    out["percent_dense"] = np.float32(m.percent_dense)
    snapshot(m, "s0", out)

    # ---- step 1: statistics from two "views" (scene/gaussian_model.py:410-412)
    N = m._xyz.shape[0]
    for k in range(2):
        vs = torch.zeros(N, 3)
        vs.grad = torch.randn(N, 3, generator=g) * (0.002 if k == 0 else 0.0006)
        filt = torch.rand(N, generator=g) > 0.3
        out[f"view{k}_grad"], out[f"view{k}_filter"] = vs.grad.numpy().copy(), filt.numpy().copy()
        m.add_densification_stats(vs, filt)
    m.max_radii2D = torch.rand(N, generator=g) * 40.0
    out["max_radii2D_in"] = m.max_radii2D.numpy().copy()
    snapshot(m, "s1", out)

    # ---- step 2: densify_and_prune (clone + split + prune), template-stage arguments (arguments/__init__.py:100-110)
    args = dict(max_grad=0.0002, min_opacity=0.05, extent=2.0, max_screen_size=20)
    out["dp_args"] = np.array([args["max_grad"], args["min_opacity"], args["extent"], args["max_screen_size"]], dtype=np.float64)
    out["split_seed"] = np.int64(777)
    torch.manual_seed(777)                       # the split's torch.normal draws from the global CPU generator
    m.densify_and_prune(args["max_grad"], args["min_opacity"], args["extent"], args["max_screen_size"])
    snapshot(m, "s2", out)

    # ---- step 3: prune with a mask that would strip some faces bare (the keep-one-per-face rule)
    N = m._xyz.shape[0]
    mask = torch.rand(N, generator=g) > 0.25
    out["prune_mask"] = mask.numpy().copy()
    m.prune_points(mask)
    snapshot(m, "s3", out)

    # ---- step 4: reset_opacity (scene/gaussian_model.py:211-214)
    m.reset_opacity()
    snapshot(m, "s4", out)

    np.savez_compressed(os.path.join(os.environ.get("GG_GOLDEN_OUT", HERE), "densify.npz"), **out)
    print("wrote densify.npz:", {t: out[f"{t}__xyz"].shape[0] for t in ("s0", "s1", "s2", "s3", "s4")})


if __name__ == "__main__":
    main()
