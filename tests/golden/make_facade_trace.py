"""Runs the reference's OWN facade -- /root/reference/gaussian_renderer/__init__.py, unmodified -- with a
recording rasterizer in place of `diff_gaussian_rasterization_depth_alpha`, and writes what it saw to
tests/golden/facade_trace.json + facade_trace.npz (authoring container only; needs /root/reference).

    python tests/golden/make_facade_trace.py

For every scenario the trace holds: the keyword set `GaussianRasterizationSettings` was built with, the keyword set
/ dtypes / shapes / strides / requires_grad / leaf-ness of the rasterizer call, which inputs were None, the model
state and `vis_mask` the facade started from, and the keys / derivations of what it returned
(`render()` dict, gaussian_renderer/__init__.py:115-122; `doll_render()` tuple, :221).
tests/test_gpu_parity.py::test_reference_facade_trace_replay replays each call through THIS repo's package on the
GPU and checks the same contract; tests/test_facade_trace_cpu.py regenerates the trace whenever /root/reference is
present and compares it with the committed one.

Shims used only here (SURVEY.md 8b "Importing the real callers"): `scene.mesh_gaussian_model` / `scene.gaussian_model`
are stub modules (the facade imports the two classes for type hints only), `torch.zeros_like(..., device="cuda")`
(:29, :132) is redirected to the CPU.
"""

import importlib.util
import json
import math
import os
import sys
import types

import numpy as np
import torch

REF = "/root/reference"
HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
N, W, H = 600, 96, 64


class Recorder:
    """Stands in for the third-party rasterizer package."""

    def __init__(self):
        self.settings_kwargs = None
        self.call_kwargs = None

    def module(self):
        rec = self
        m = types.ModuleType("diff_gaussian_rasterization_depth_alpha")

        def GaussianRasterizationSettings(*args, **kwargs):
            assert not args, "the facade builds the settings by keyword only"
            rec.settings_kwargs = dict(kwargs)
            return ("settings", kwargs)

        class GaussianRasterizer:
            def __init__(self, *args, **kwargs):
                assert not args and list(kwargs) == ["raster_settings"]
                self.rs = kwargs["raster_settings"][1]

            def __call__(self, *args, **kwargs):
                assert not args, "the facade calls the rasterizer by keyword only"
                rec.call_kwargs = dict(kwargs)
                n = kwargs["means3D"].shape[0]
                h, w = self.rs["image_height"], self.rs["image_width"]
                dep = kwargs["means3D"].sum() * 0 + kwargs["means2D"].sum() * 0          # keeps the graph alive
                color = torch.full((3, h, w), 0.25) + dep
                radii = (torch.arange(n) % 3).to(torch.int32)                            # every third one "invisible"
                return color, radii, torch.full((1, h, w), 2.0) + dep, torch.full((1, h, w), 0.5) + dep

        m.GaussianRasterizationSettings = GaussianRasterizationSettings
        m.GaussianRasterizer = GaussianRasterizer
        return m


def import_facade(rec):
    sys.path.insert(0, REF)
    sys.modules["diff_gaussian_rasterization_depth_alpha"] = rec.module()
    scene_pkg = types.ModuleType("scene")
    scene_pkg.__path__ = []
    mgm = types.ModuleType("scene.mesh_gaussian_model")
    mgm.MeshGaussianModel = type("MeshGaussianModel", (), {})
    gm = types.ModuleType("scene.gaussian_model")
    gm.GaussianModel = type("GaussianModel", (), {})
    sys.modules.update({"scene": scene_pkg, "scene.mesh_gaussian_model": mgm, "scene.gaussian_model": gm})
    _zl = torch.zeros_like

    def zeros_like_cpu(*a, **k):
        k.pop("device", None)
        return _zl(*a, **k)

    torch.zeros_like = zeros_like_cpu
    spec = importlib.util.spec_from_file_location("ref_gaussian_renderer", os.path.join(REF, "gaussian_renderer", "__init__.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def build_state(seed=5):
    sys.path.insert(0, ROOT)
    import diff_gaussian_rasterization_depth_alpha  # noqa: F401  (this repo's shim; registers gaussian_garments_b200)
    import gaussian_garments_b200 as gg
    st = gg.scenes.random_cloud(N, seed=seed)
    cam = gg.scenes.cfg1_camera(W, H)
    return gg, st, cam


class PC:
    """Attribute surface of GaussianModel / MeshGaussianModel / AvatarGaussianModel that render() touches
    (scene/gaussian_model.py:95-119, scene/avatar_gaussian_model.py:140-148, scene/avatar_net.py:82-84)."""

    def __init__(self, st, gen_utils, avatar=False):
        leaf = lambda t: t.clone().requires_grad_(True)
        self._xyz = leaf(st.means3D)
        self._scaling = leaf(torch.log(st.scales))
        self._rotation = leaf(st.rotations)
        self._opacity = leaf(torch.logit(st.opacities.clamp(1e-4, 1 - 1e-4)))
        self._features = leaf(st.shs)
        self.active_sh_degree = st.sh_degree
        self.max_sh_degree = 3
        self._gen = gen_utils
        if avatar:
            self.local_xyz = self._xyz + 0.01          # AvatarNet: local_xyz = _xyz + offset (scene/avatar_net.py:82)
            self.shs = self._features * 1.5            # per-frame SHs from the appearance net (:84)

    get_xyz = property(lambda s: s._xyz)
    get_final_xyz = property(lambda s: s.local_xyz * 1.0 + 0.5)
    get_opacity = property(lambda s: torch.sigmoid(s._opacity))
    get_scaling = property(lambda s: torch.exp(s._scaling))
    get_rotation = property(lambda s: torch.nn.functional.normalize(s._rotation))
    get_features = property(lambda s: s._features)

    def get_covariance(self, scaling_modifier=1):
        L = self._gen.build_scaling_rotation(scaling_modifier * self.get_scaling, self._rotation)
        return self._gen.strip_symmetric(L @ L.transpose(1, 2))


class Doll:
    """inference.py's Doll: plain attributes (gaussian_renderer/__init__.py:159-192)."""

    def __init__(self, st, gen_utils):
        self.xyz, self.opacity, self.scaling, self.rotation = st.means3D, st.opacities, st.scales, st.rotations
        self.features = st.shs
        self.active_sh_degree, self.max_sh_degree = st.sh_degree, 3
        self._gen = gen_utils

    def covariance(self, scaling_modifier=1):
        L = self._gen.build_scaling_rotation(scaling_modifier * self.scaling, self.rotation)
        return self._gen.strip_symmetric(L @ L.transpose(1, 2))


def tmeta(t):
    if t is None:
        return None
    return dict(shape=list(t.shape), dtype=str(t.dtype).replace("torch.", ""), stride=list(t.stride()),
                requires_grad=bool(t.requires_grad), is_leaf=bool(t.is_leaf), contiguous=bool(t.is_contiguous()))


def main():
    gg, st, cam = build_state()           # with this repo's shim; the recorder replaces it for the facade import
    rec = Recorder()
    facade = import_facade(rec)
    # utils.general_utils hard-codes device="cuda" in build_rotation/strip_symmetric: CPU redirect for torch.zeros
    _z = torch.zeros

    def zeros_cpu(*a, **k):
        k.pop("device", None)
        return _z(*a, **k)

    torch.zeros = zeros_cpu
    sys.modules.setdefault("open3d", types.ModuleType("open3d"))
    cams_stub = types.ModuleType("scene.cameras")
    cams_stub.Camera = object
    sys.modules["scene.cameras"] = cams_stub
    from utils import general_utils as gen

    g = torch.Generator().manual_seed(11)
    vis_mask = torch.rand(N, generator=g) > 0.35
    override_color = torch.rand(N, 3, generator=g)
    override_shs = torch.randn(N, 16, 3, generator=g) * 0.2
    bg = st.bg.clone()
    mk_pipe = lambda cov=False, sh=False: types.SimpleNamespace(debug=False, compute_cov3D_python=cov, convert_SHs_python=sh)

    scenarios = [
        dict(name="render_default", fn="render", pipe=(False, False)),
        dict(name="render_vis_mask_avatar", fn="render", pipe=(False, False), avatar=True, vis_mask=True),
        dict(name="render_convert_SHs_python", fn="render", pipe=(False, True)),
        dict(name="render_compute_cov3D_python", fn="render", pipe=(True, False), scaling_modifier=1.3),
        dict(name="render_override_color_masked", fn="render", pipe=(False, False), override_color=True, vis_mask=True),
        dict(name="doll_default", fn="doll_render", pipe=(False, False)),
        dict(name="doll_override_shs_masked", fn="doll_render", pipe=(False, False), override_shs=True, vis_mask=True),
        dict(name="doll_cov_python_override_color", fn="doll_render", pipe=(True, False), override_color=True),
    ]
    arrays = dict(means3D=st.means3D.numpy(), scales=st.scales.numpy(), rotations=st.rotations.numpy(),
                  opacities=st.opacities.numpy(), shs=st.shs.numpy(), bg=bg.numpy(), vis_mask=vis_mask.numpy(),
                  override_color=override_color.numpy(), override_shs=override_shs.numpy())
    trace = dict(reference_file="gaussian_renderer/__init__.py", N=N, W=W, H=H, sh_degree=st.sh_degree,
                 camera=dict(width=W, height=H, fov_deg=50.0, note="gg.scenes.cfg1_camera(W, H)"), scenarios=[])

    for sc in scenarios:
        pipe = mk_pipe(*sc["pipe"])
        kw = {}
        if "scaling_modifier" in sc:
            kw["scaling_modifier"] = sc["scaling_modifier"]
        if sc.get("override_color"):
            kw["override_color"] = override_color
        if sc.get("vis_mask"):
            kw["vis_mask"] = vis_mask
        if sc["fn"] == "render":
            pc = PC(st, gen, avatar=sc.get("avatar", False))
            out = facade.render(cam, pc, pipe, bg, **kw)
            ret = dict(type="dict", keys=list(out.keys()),
                       visibility_filter_is_radii_gt_0=bool(torch.equal(out["visibility_filter"], out["radii"] > 0)),
                       viewspace_points=tmeta(out["viewspace_points"]),
                       position_is_call_means3D=bool(out["3dposition"] is rec.call_kwargs["means3D"]),
                       render_is_first_output=True)
            # the side channel: backward through the stub must leave a grad on the retained non-leaf
            out["render"].sum().backward()
            ret["viewspace_points_grad_shape"] = list(out["viewspace_points"].grad.shape)
        else:
            if sc.get("override_shs"):
                kw["override_shs"] = override_shs
            pc = Doll(st, gen)
            with torch.no_grad():                                   # inference.py:462
                out = facade.doll_render(cam, pc, pipe, bg, **kw)
            ret = dict(type="tuple", length=len(out), order=["rendered_image", "depth", "alpha"],
                       shapes=[list(o.shape) for o in out])
        skw = rec.settings_kwargs
        ckw = rec.call_kwargs
        entry = dict(name=sc["name"], fn=sc["fn"], pipe=dict(compute_cov3D_python=sc["pipe"][0], convert_SHs_python=sc["pipe"][1]),
                     options={k: v for k, v in sc.items() if k not in ("name", "fn", "pipe")},
                     settings_keys=list(skw.keys()),
                     settings_scalars={k: (v if not torch.is_tensor(v) else None) for k, v in skw.items()},
                     settings_tensors={k: tmeta(v) for k, v in skw.items() if torch.is_tensor(v)},
                     call_keys=list(ckw.keys()), call=dict((k, tmeta(v)) for k, v in ckw.items()), returns=ret)
        for k, v in ckw.items():
            if v is not None and k in ("colors_precomp", "cov3D_precomp"):     # values the reference's own code computed
                arrays[f"{sc['name']}__{k}"] = v.detach().numpy()
        trace["scenarios"].append(entry)

    out_dir = os.environ.get("GG_FACADE_TRACE_OUT", HERE)
    with open(os.path.join(out_dir, "facade_trace.json"), "w") as f:
        json.dump(trace, f, indent=1, sort_keys=True)
    np.savez_compressed(os.path.join(out_dir, "facade_trace.npz"), **arrays)
    print(f"wrote facade_trace.json / .npz ({len(trace['scenarios'])} scenarios)")
    return trace


if __name__ == "__main__":
    main()
