"""Generates tests/golden/*.npz by IMPORTING THE REFERENCE (run in the authoring container only).

    python tests/golden/make_golden.py        # needs /root/reference; writes small .npz fixtures

The rasterizer itself is not in /root/reference (third-party CUDA extension), so
these vectors pin the in-tree pieces of the path's arithmetic that the oracle and the kernels
restate:
  sh.npz       utils/sh_utils.py:eval_sh  (+0.5 / clamp of gaussian_renderer/__init__.py:84-85)
  cov3d.npz    utils/general_utils.py:build_scaling_rotation + strip_symmetric, composed as
               scene/gaussian_model.py:27-31
  camera.npz   utils/graphics_utils.py:getWorld2View2 / getProjectionMatrix / focal2fov composed
               exactly as scene/cameras.py:53-62, plus projected pixel centres of sample points
  mesh.npz     utils/graphics_utils.py:118-137 compute_face_orientation(return_scale=True) on a random triangle
               soup (incl. one degenerate face) + the face centres of scene/mesh_gaussian_model.py:92
Shims used only here: `open3d` is stubbed (utils/general_utils.py:18 imports it but the two
functions we call do not use it) and torch.zeros(..., device="cuda") is redirected to CPU
(utils/general_utils.py:75,93,112 hard-code the device).
"""

import math
import os
import sys
import types

import numpy as np
import torch

REF = "/root/reference"
OUT = os.path.dirname(os.path.abspath(__file__))


def _import_reference():
    sys.path.insert(0, REF)
    sys.modules.setdefault("open3d", types.ModuleType("open3d"))
    # scene.cameras is imported by utils.general_utils only for type use; provide a light stub
    # so that importing it does not drag scene/__init__.py (which needs many absent packages).
    scene_pkg = types.ModuleType("scene")
    scene_pkg.__path__ = []
    cams = types.ModuleType("scene.cameras")
    cams.Camera = object
    sys.modules["scene"] = scene_pkg
    sys.modules["scene.cameras"] = cams
    _zeros = torch.zeros

    def zeros_cpu(*a, **k):
        k.pop("device", None)
        return _zeros(*a, **k)

    torch.zeros = zeros_cpu
    from utils import sh_utils, graphics_utils, general_utils  # noqa
    return sh_utils, graphics_utils, general_utils


def main():
    sh_utils, gu, gen = _import_reference()
    g = torch.Generator().manual_seed(31359)

    # ---- SH ----
    n = 64
    dirs = torch.nn.functional.normalize(torch.randn(n, 3, generator=g, dtype=torch.float64))
    shs = torch.randn(n, 16, 3, generator=g, dtype=torch.float64)            # rasterizer layout [N,M,3]
    out = {}
    for deg in range(4):
        res = sh_utils.eval_sh(deg, shs.transpose(1, 2), dirs)                 # eval_sh wants [N,3,M]
        out[f"rgb_deg{deg}"] = torch.clamp_min(res + 0.5, 0.0).numpy()
        out[f"raw_deg{deg}"] = res.numpy()
    np.savez(os.path.join(OUT, "sh.npz"), dirs=dirs.numpy(), shs=shs.numpy(), **out)

    # ---- covariance ----
    scales = torch.exp(torch.rand(n, 3, generator=g) * 3 - 4)
    rots = torch.nn.functional.normalize(torch.randn(n, 4, generator=g))
    covs = {}
    for mod in (1.0, 1.7):
        L = gen.build_scaling_rotation(mod * scales, rots)
        covs[f"cov_mod{mod}"] = gen.strip_symmetric(L @ L.transpose(1, 2)).numpy()
    np.savez(os.path.join(OUT, "cov3d.npz"), scales=scales.numpy(), rotations=rots.numpy(), **covs)

    # ---- cameras ----
    rng = np.random.RandomState(7)
    cams = []
    for k, (w, h) in enumerate([(512, 512), (1920, 1080), (940, 1280)]):
        ax = rng.randn(3); ax /= np.linalg.norm(ax)
        ang = 0.3 * (k + 1)
        Kx = np.array([[0, -ax[2], ax[1]], [ax[2], 0, -ax[0]], [-ax[1], ax[0], 0]])
        Rw2c = np.eye(3) + math.sin(ang) * Kx + (1 - math.cos(ang)) * Kx @ Kx
        R = Rw2c.T                                                           # scene/dataloader.py:175
        T = np.array([0.1 * k, -0.2, 3.0 + k])
        fx = 0.9 * w + 10 * k; fy = 0.95 * w
        cx = w / 2 + 3.5 - k; cy = h / 2 - 2.25 + 2 * k
        FoVx, FoVy = gu.focal2fov(fx, w), gu.focal2fov(fy, h)
        wvt = torch.tensor(gu.getWorld2View2(R, T, np.array([0.0, 0.0, 0.0]), 1.0)).transpose(0, 1)
        proj = gu.getProjectionMatrix(znear=0.01, zfar=100.0, fovX=FoVx, fovY=FoVy, fx=fx, fy=fy,
                                      cx=cx, cy=cy, w=w, h=h).transpose(0, 1)
        full = (wvt.unsqueeze(0).bmm(proj.unsqueeze(0))).squeeze(0)
        center = wvt.inverse()[3, :3]
        pts = torch.tensor(rng.uniform(-1, 1, size=(32, 3)).astype(np.float32))
        hom = torch.cat([pts, torch.ones(32, 1)], 1) @ full
        ndc = hom[:, :2] / (hom[:, 3:4] + 1e-7)
        pix = torch.stack([((ndc[:, 0] + 1) * w - 1) * 0.5, ((ndc[:, 1] + 1) * h - 1) * 0.5], 1)
        zview = (torch.cat([pts, torch.ones(32, 1)], 1) @ wvt)[:, 2]
        cams.append(dict(R=R, T=T, fx=fx, fy=fy, cx=cx, cy=cy, w=w, h=h, FoVx=FoVx, FoVy=FoVy,
                         wvt=wvt.numpy(), proj=proj.numpy(), full=full.numpy(), center=center.numpy(),
                         pts=pts.numpy(), pix=pix.numpy(), zview=zview.numpy()))
    flat = {}
    for k, c in enumerate(cams):
        for name, v in c.items():
            flat[f"cam{k}_{name}"] = np.asarray(v)
    np.savez(os.path.join(OUT, "camera.npz"), n_cams=len(cams), **flat)
    # ---- photometric loss (utils/loss_utils.py); the reference's ssim() mutates its inputs in place -> clones
    from utils import loss_utils
    gl = torch.Generator().manual_seed(99)
    out = {}
    for tag, (h, w, use_mask) in {"a": (40, 52, False), "b": (37, 45, True)}.items():
        img = torch.rand(3, h, w, generator=gl)
        gt = (img + 0.2 * torch.randn(3, h, w, generator=gl)).clamp(0, 1)
        mask = (torch.rand(1, h, w, generator=gl) > 0.3).float() if use_mask else None
        x = img.clone().requires_grad_(True)
        l1 = loss_utils.l1_loss(x, gt, mask)
        xs = x * 1.0                                      # non-leaf copy: ssim() multiplies it by the mask in place
        ss = loss_utils.ssim(xs, gt.clone(), mask.clone() if mask is not None else None)
        total = l1 * (1.0 - 0.2) + (1.0 - ss * 0.2)       # s2_registration.py:259-260 with lambda_dssim = 0.2
        total.backward()
        out.update({f"{tag}_img": img.numpy(), f"{tag}_gt": gt.numpy(), f"{tag}_l1": l1.item(), f"{tag}_ssim": ss.item(),
                    f"{tag}_total": total.item(), f"{tag}_grad": x.grad.numpy()})
        if mask is not None:
            out[f"{tag}_mask"] = mask.numpy()
    np.savez(os.path.join(OUT, "loss.npz"), **out)

    # ---- mesh face frames (utils/graphics_utils.py:118-137 compute_face_orientation, return_scale=True) --------
    gm = torch.Generator().manual_seed(4242)
    V, F = 40, 64
    verts = torch.randn(V, 3, generator=gm) * 0.3
    faces = torch.stack([torch.randperm(V, generator=gm)[:3] for _ in range(F)]).long()
    faces[-1] = torch.tensor([0, 1, 1])                       # degenerate face: exercises the 1e-20 clamps
    orient, scale = gu.compute_face_orientation(verts, faces, return_scale=True)
    vd = verts.double()
    orient64, scale64 = gu.compute_face_orientation(vd, faces, return_scale=True)
    np.savez(os.path.join(OUT, "mesh.npz"), verts=verts.numpy(), faces=faces.numpy(), orientation=orient.numpy(),
             scale=scale.numpy(), orientation64=orient64.numpy(), scale64=scale64.numpy(),
             center=verts[faces].mean(1).numpy())
    print("wrote sh.npz cov3d.npz camera.npz loss.npz mesh.npz to", OUT)


if __name__ == "__main__":
    main()
