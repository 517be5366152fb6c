"""Shared helpers of the parity tests (test infrastructure)."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

import diff_gaussian_rasterization_depth_alpha as dgr  # noqa: E402  (registers gaussian_garments_b200)
import gaussian_garments_b200 as gg  # noqa: E402
from oracle import c_oracle, torch_oracle  # noqa: E402

RGB_TOL = 1e-4      # BASELINE.json north_star: RGB within 1e-4 abs
GRAD_TOL = 1e-3     # gradients within 1e-3 rel


def settings_for(cam, state, device=None, sh_degree=None, scale_modifier=1.0, debug=False):
    dev = device if device is not None else state.means3D.device
    return dgr.GaussianRasterizationSettings(
        image_height=int(cam.image_height), image_width=int(cam.image_width),
        tanfovx=cam.tanfovx, tanfovy=cam.tanfovy, bg=state.bg.to(dev), scale_modifier=scale_modifier,
        viewmatrix=cam.world_view_transform.to(dev), projmatrix=cam.full_proj_transform.to(dev),
        sh_degree=state.sh_degree if sh_degree is None else sh_degree, campos=cam.camera_center.to(dev),
        prefiltered=False, debug=debug)


def cpu_settings(s):
    return tuple(v.detach().cpu() if torch.is_tensor(v) else v for v in s)


def run_cuda(settings, st, grads=None, colors_precomp=None, cov3D_precomp=None, use_shs=True):
    """Forward (+ backward with upstream grads (Gc, Gd, Ga)) through the public API on the GPU."""
    dev = torch.device("cuda:0")
    leaf = lambda t: None if t is None else t.detach().to(dev).clone().requires_grad_(True)
    m3, op = leaf(st.means3D), leaf(st.opacities)
    shs = leaf(st.shs) if (use_shs and colors_precomp is None) else None
    col = leaf(colors_precomp)
    sc = leaf(st.scales) if cov3D_precomp is None else None
    ro = leaf(st.rotations) if cov3D_precomp is None else None
    cv = leaf(cov3D_precomp)
    m2 = torch.zeros_like(m3, requires_grad=True)
    rast = dgr.GaussianRasterizer(raster_settings=settings)
    color, radii, depth, alpha = rast(means3D=m3, means2D=m2, shs=shs, colors_precomp=col, opacities=op,
                                      scales=sc, rotations=ro, cov3D_precomp=cv)
    out = dict(color=color.detach().cpu(), radii=radii.cpu(), depth=depth.detach().cpu(), alpha=alpha.detach().cpu())
    if grads is not None:
        Gc, Gd, Ga = (g.to(dev) if g is not None else None for g in grads)
        loss = (color * Gc).sum()
        if Gd is not None:
            loss = loss + (depth * Gd).sum()
        if Ga is not None:
            loss = loss + (alpha * Ga).sum()
        loss.backward()
        g = lambda t: None if t is None else t.grad.detach().cpu()
        out["grads"] = dict(means3D=g(m3), means2D=g(m2), shs=g(shs), colors_precomp=g(col), opacities=g(op),
                            scales=g(sc), rotations=g(ro), cov3D_precomp=g(cv))
    return out


def run_c_oracle(settings, st, grads=None, colors_precomp=None, cov3D_precomp=None, fragile_eps=2e-4):
    S = cpu_settings(settings)
    use_cov = cov3D_precomp is not None
    color, radii, depth, alpha, ctx, frag = c_oracle.rasterize_forward(
        S, st.means3D, None if colors_precomp is not None else st.shs, colors_precomp, st.opacities,
        None if use_cov else st.scales, None if use_cov else st.rotations, cov3D_precomp, fragile_eps=fragile_eps)
    out = dict(color=color, radii=radii, depth=depth, alpha=alpha, fragile=frag, K=ctx.num_rendered, ctx=ctx)
    if grads is not None:
        out["grads"] = ctx.backward(*grads)
    return out


def rel_inf(a, b):
    return float((a.float() - b.float()).abs().max() / b.float().abs().max().clamp_min(1e-12))


def assert_images_close(got, ref, tol=RGB_TOL, max_fragile_frac=1e-2):
    """All pixels within tol except those the oracle flags as sitting on a discrete threshold
    (alpha ~ 1/255 or T ~ 1e-4), which must stay a tiny fraction."""
    frag = ref["fragile"]
    for name, t in (("color", tol), ("alpha", tol), ("depth", None)):
        d = (got[name] - ref[name]).abs()
        if name == "depth":   # un-normalised depth carries the scene scale: relative tolerance
            t = tol * max(1.0, float(ref[name].abs().max()))
        d = d.amax(0)
        bad = d > t
        unexplained = bad & ~frag
        assert int(unexplained.sum()) == 0, (
            f"{name}: {int(unexplained.sum())} pixels differ by more than {t} (max {float(d.max()):.3e}) "
            f"away from discrete thresholds")
    assert float(frag.float().mean()) <= max_fragile_frac or int(frag.sum()) < 64


def assert_grads_close(got, ref, tol=GRAD_TOL):
    for name, g_ref in ref.items():
        if g_ref is None:
            continue
        g = got[name]
        assert g is not None, f"missing gradient {name}"
        assert g.shape == g_ref.shape, (name, g.shape, g_ref.shape)
        assert torch.isfinite(g).all(), f"non-finite gradient {name}"
        err = rel_inf(g, g_ref)
        assert err <= tol, f"grad {name}: rel-inf error {err:.3e} > {tol}"
