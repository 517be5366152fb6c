"""TEST INFRASTRUCTURE -- PyTorch autograd oracle of the Gaussian-splat rasterizer hot path.

**PARITY UNPINNED**: the reference's rasterizer is the third-party, un-vendored,
un-pinned CUDA extension `diff_gaussian_rasterization_depth_alpha`
(github.com/lizhe00/AnimatableGaussians, gaussians/diff_gaussian_rasterization_depth_alpha,
cloned at HEAD by /root/reference/setup.sh:26-29).  Its sources are not in
/root/reference and the reference holds no tests or golden vectors for this path
(SURVEY.md sections 4 and 8c).  This file restates the *published* 3DGS rasterization
algorithm with the depth/alpha outputs, anchored on the reference's call sites
(gaussian_renderer/__init__.py:39-54,103-111) and on the in-tree pieces of the same
math that ARE pinned by tests/golden/ (SH basis utils/sh_utils.py:56-111, covariance
scene/gaussian_model.py:27-31 + utils/general_utils.py:74-120, camera matrices
scene/cameras.py:53-62 + utils/graphics_utils.py:38-81).

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
legs may import this module.  The product path never does.

Gradients come from autograd through the forward pass (so they are independent of the
hand-derived backward in oracle/gg_oracle.c and in the CUDA kernels), with three
deliberate emulations of the rasterizer's conventions:
  * straight-through 0.99 clamp of alpha,
  * `means2D` gradient side channel in (W/2, H/2)-scaled units
    (consumer: scene/gaussian_model.py:410-412),
  * the 1.3*tanfov clamp of t.x/t.z, t.y/t.z gates d/dt.x, d/dt.y and treats the clamped
    value as constant.
Discrete decisions (culling, radius, tile rectangle, depth order, alpha<1/255,
power>0, T<1e-4 stop) are always taken in fp32 from the fp32 inputs, also in fp64 mode.
"""

from __future__ import annotations

import math
from typing import NamedTuple, Optional

import torch

TILE = 16
NEAR_Z = 0.2
ALPHA_MIN = 1.0 / 255.0
ALPHA_MAX = 0.99
T_STOP = 1e-4
BLUR = 0.3

# utils/sh_utils.py:25-42
SH_C0 = 0.28209479177387814
SH_C1 = 0.4886025119029199
SH_C2 = [1.0925484305920792, -1.0925484305920792, 0.31539156525252005,
         -1.0925484305920792, 0.5462742152960396]
SH_C3 = [-0.5900435899266435, 2.890611442640554, -0.4570457994644658, 0.3731763325901154,
         -0.4570457994644658, 1.445305721320277, -0.5900435899266435]


class OracleSettings(NamedTuple):
    """Same 12 fields as GaussianRasterizationSettings (gaussian_renderer/__init__.py:39-52)."""
    image_height: int
    image_width: int
    tanfovx: float
    tanfovy: float
    bg: torch.Tensor
    scale_modifier: float
    viewmatrix: torch.Tensor
    projmatrix: torch.Tensor
    sh_degree: int
    campos: torch.Tensor
    prefiltered: bool
    debug: bool


def eval_sh_rgb(deg: int, shs: torch.Tensor, dirs: torch.Tensor) -> torch.Tensor:
    """shs [N,M,3] (rasterizer layout), dirs [N,3] unit -> [N,3]; basis of utils/sh_utils.py:56-111."""
    res = SH_C0 * shs[:, 0]
    if deg > 0:
        x, y, z = dirs[:, 0:1], dirs[:, 1:2], dirs[:, 2:3]
        res = res - SH_C1 * y * shs[:, 1] + SH_C1 * z * shs[:, 2] - SH_C1 * x * shs[:, 3]
        if deg > 1:
            xx, yy, zz = x * x, y * y, z * z
            xy, yz, xz = x * y, y * z, x * z
            res = (res + SH_C2[0] * xy * shs[:, 4] + SH_C2[1] * yz * shs[:, 5]
                   + SH_C2[2] * (2.0 * zz - xx - yy) * shs[:, 6]
                   + SH_C2[3] * xz * shs[:, 7] + SH_C2[4] * (xx - yy) * shs[:, 8])
            if deg > 2:
                res = (res + SH_C3[0] * y * (3.0 * xx - yy) * shs[:, 9]
                       + SH_C3[1] * xy * z * shs[:, 10]
                       + SH_C3[2] * y * (4.0 * zz - xx - yy) * shs[:, 11]
                       + SH_C3[3] * z * (2.0 * zz - 3.0 * xx - 3.0 * yy) * shs[:, 12]
                       + SH_C3[4] * x * (4.0 * zz - xx - yy) * shs[:, 13]
                       + SH_C3[5] * z * (xx - yy) * shs[:, 14]
                       + SH_C3[6] * x * (xx - 3.0 * yy) * shs[:, 15])
    return res


def rotation_matrix(q: torch.Tensor) -> torch.Tensor:
    """wxyz quaternion -> R, entries of utils/general_utils.py:100-108, WITHOUT renormalisation
    (the kernel consumes the pre-normalised quaternion as is)."""
    r, x, y, z = q[:, 0], q[:, 1], q[:, 2], q[:, 3]
    R = torch.stack([
        1 - 2 * (y * y + z * z), 2 * (x * y - r * z), 2 * (x * z + r * y),
        2 * (x * y + r * z), 1 - 2 * (x * x + z * z), 2 * (y * z - r * x),
        2 * (x * z - r * y), 2 * (y * z + r * x), 1 - 2 * (x * x + y * y)], dim=-1)
    return R.reshape(-1, 3, 3)


def covariance3d(scales, mod, rotations):
    """Sigma = (R S)(R S)^T packed xx,xy,xz,yy,yz,zz (scene/gaussian_model.py:27-31,
    utils/general_utils.py:74-83,110-120)."""
    L = rotation_matrix(rotations) * (mod * scales)[:, None, :]
    S = L @ L.transpose(1, 2)
    return torch.stack([S[:, 0, 0], S[:, 0, 1], S[:, 0, 2], S[:, 1, 1], S[:, 1, 2], S[:, 2, 2]], dim=-1)


def _project(means3D, cov6, s: OracleSettings, dt):
    """Per-Gaussian geometry in dtype dt (differentiable)."""
    W, H = s.image_width, s.image_height
    V = s.viewmatrix.to(dt)          # = W2C^T  (scene/cameras.py:59)
    Pm = s.projmatrix.to(dt)         # = (P W2C)^T
    ones = torch.ones_like(means3D[:, :1])
    hom = torch.cat([means3D, ones], dim=1)
    p_view = hom @ V                  # row-vector convention
    p_hom = hom @ Pm
    p_w = 1.0 / (p_hom[:, 3] + 1e-7)
    ndc = p_hom[:, :2] * p_w[:, None]
    tz = p_view[:, 2]
    limx, limy = 1.3 * s.tanfovx, 1.3 * s.tanfovy
    txtz = p_view[:, 0] / tz
    tytz = p_view[:, 1] / tz
    in_x = (txtz >= -limx) & (txtz <= limx)
    in_y = (tytz >= -limy) & (tytz <= limy)
    tx = torch.where(in_x, p_view[:, 0], (txtz.clamp(-limx, limx) * tz).detach())
    ty = torch.where(in_y, p_view[:, 1], (tytz.clamp(-limy, limy) * tz).detach())
    fx = W / (2.0 * s.tanfovx)
    fy = H / (2.0 * s.tanfovy)
    zero = torch.zeros_like(tz)
    J = torch.stack([fx / tz, zero, -fx * tx / (tz * tz),
                     zero, fy / tz, -fy * ty / (tz * tz)], dim=-1).reshape(-1, 2, 3)
    W3 = V[:3, :3].transpose(0, 1)    # rotation part of W2C
    Tm = J @ W3                       # [N,2,3]
    Sig = torch.stack([cov6[:, 0], cov6[:, 1], cov6[:, 2],
                       cov6[:, 1], cov6[:, 3], cov6[:, 4],
                       cov6[:, 2], cov6[:, 4], cov6[:, 5]], dim=-1).reshape(-1, 3, 3)
    c2 = Tm @ Sig @ Tm.transpose(1, 2)
    a = c2[:, 0, 0] + BLUR
    b = c2[:, 0, 1]
    c = c2[:, 1, 1] + BLUR
    det = a * c - b * b
    conic = torch.stack([c / det, -b / det, a / det], dim=-1)
    pix = torch.stack([((ndc[:, 0] + 1.0) * W - 1.0) * 0.5, ((ndc[:, 1] + 1.0) * H - 1.0) * 0.5], dim=-1)
    return dict(depth=tz, pix=pix, conic=conic, a=a, b=b, c=c, det=det)


def _discrete(geo32, s: OracleSettings):
    """Culling / radius / tile rectangle, fp32, no grad."""
    W, H = s.image_width, s.image_height
    gx, gy = (W + TILE - 1) // TILE, (H + TILE - 1) // TILE
    a, c, det = geo32["a"], geo32["c"], geo32["det"]
    mid = 0.5 * (a + c)
    lam = mid + torch.sqrt(torch.clamp(mid * mid - det, min=0.1))
    radius = torch.ceil(3.0 * torch.sqrt(lam))
    px, py = geo32["pix"][:, 0], geo32["pix"][:, 1]

    def tdiv(v):  # C-style (int)(v / 16.0f)
        return torch.trunc(v / float(TILE)).to(torch.int64)

    x0 = tdiv(px - radius).clamp(0, gx)
    y0 = tdiv(py - radius).clamp(0, gy)
    x1 = tdiv(px + radius + (TILE - 1)).clamp(0, gx)
    y1 = tdiv(py + radius + (TILE - 1)).clamp(0, gy)
    ok = (geo32["depth"] > NEAR_Z) & (det != 0) & ((x1 - x0) * (y1 - y0) > 0)
    ok = ok & torch.isfinite(radius) & torch.isfinite(px) & torch.isfinite(py)
    radii = torch.where(ok, radius, torch.zeros_like(radius)).to(torch.int32)
    return ok, radii, x0, y0, x1, y1, gx, gy


def rasterize(settings, means3D, means2D=None, shs=None, colors_precomp=None, opacities=None,
              scales=None, rotations=None, cov3D_precomp=None, dtype=torch.float32,
              return_aux=False):
    """Differentiable rasterization; returns (color[3,H,W], radii[N] int32, depth[1,H,W], alpha[1,H,W])."""
    s = OracleSettings(*settings)
    dt = dtype
    W, H = int(s.image_width), int(s.image_height)
    N = means3D.shape[0]
    dev = means3D.device
    if N == 0:
        z = torch.zeros
        out = (z(3, H, W, dtype=dt), z(0, dtype=torch.int32), z(1, H, W, dtype=dt), z(1, H, W, dtype=dt))
        return out + ({},) if return_aux else out

    def cast(t):
        return None if t is None else t.to(dt)

    m3 = cast(means3D)
    if cov3D_precomp is not None and cov3D_precomp.numel() > 0:
        cov6 = cast(cov3D_precomp)
    else:
        cov6 = covariance3d(cast(scales), float(s.scale_modifier), cast(rotations))
    geo = _project(m3, cov6, s, dt)
    with torch.no_grad():
        if dt == torch.float32:
            geo32 = {k: v.detach() for k, v in geo.items()}
        else:
            m32 = means3D.detach().float()
            if cov3D_precomp is not None and cov3D_precomp.numel() > 0:
                c32 = cov3D_precomp.detach().float()
            else:
                c32 = covariance3d(scales.detach().float(), float(s.scale_modifier), rotations.detach().float())
            geo32 = _project(m32, c32, s, torch.float32)
        ok, radii, x0, y0, x1, y1, gx, gy = _discrete(geo32, s)

    # colour
    if colors_precomp is not None and colors_precomp.numel() > 0:
        rgb = cast(colors_precomp)
    else:
        d = m3 - s.campos.to(dt)[None, :]
        d = d / d.norm(dim=1, keepdim=True)
        rgb = torch.clamp_min(eval_sh_rgb(int(s.sh_degree), cast(shs), d) + 0.5, 0.0)

    pix = geo["pix"]
    if means2D is not None:
        pix = pix + cast(means2D)[:, :2] * torch.tensor([0.5 * W, 0.5 * H], dtype=dt, device=dev)
    opac = cast(opacities).reshape(-1)
    depth_g = geo["depth"]
    conic = geo["conic"]
    bg = s.bg.to(dt)

    color = torch.zeros(3, H, W, dtype=dt, device=dev) + bg[:, None, None]
    depth_img = torch.zeros(1, H, W, dtype=dt, device=dev)
    alpha_img = torch.zeros(1, H, W, dtype=dt, device=dev)
    ncontrib = torch.zeros(H, W, dtype=torch.int32, device=dev)
    fragile = torch.zeros(H, W, dtype=torch.bool, device=dev)

    # ---- binning (no grad): instance list sorted by (tile, depth fp32, index) ----
    with torch.no_grad():
        idx_ok = torch.nonzero(ok).reshape(-1)
        nx = (x1 - x0)[idx_ok]
        ny = (y1 - y0)[idx_ok]
        cnt = nx * ny
        K = int(cnt.sum())
        rep = torch.repeat_interleave(torch.arange(idx_ok.numel(), device=dev), cnt)
        start = torch.cumsum(cnt, 0) - cnt
        local = torch.arange(K, device=dev) - start[rep]
        g_id = idx_ok[rep]
        tx_ = x0[g_id] + local % nx[rep]
        ty_ = y0[g_id] + local // nx[rep]
        tile_id = ty_ * gx + tx_
        d32 = geo32["depth"][g_id]
        # stable lexicographic sort: by index (already ascending), then depth, then tile
        o1 = torch.sort(d32, stable=True).indices
        o2 = torch.sort(tile_id[o1], stable=True).indices
        order = o1[o2]
        g_sorted = g_id[order]
        t_sorted = tile_id[order]
        tiles, counts = torch.unique_consecutive(t_sorted, return_counts=True)
        offs = torch.cumsum(counts, 0) - counts
        pix32 = geo32["pix"]
        if means2D is not None:
            pix32 = pix32 + means2D.detach().float()[:, :2] * torch.tensor([0.5 * W, 0.5 * H], device=dev)
        con32 = geo32["conic"]
        op32 = opacities.detach().float().reshape(-1)

    lx = torch.arange(TILE, device=dev)
    for t, o, n in zip(tiles.tolist(), offs.tolist(), counts.tolist()):
        tyi, txi = divmod(t, gx)
        ids = g_sorted[o:o + n]
        xs = (txi * TILE + lx)
        ys = (tyi * TILE + lx)
        xs = xs[xs < W]
        ys = ys[ys < H]
        PX = xs[None, :].expand(ys.numel(), xs.numel()).reshape(-1)
        PY = ys[:, None].expand(ys.numel(), xs.numel()).reshape(-1)
        with torch.no_grad():
            dx32 = pix32[ids, 0][None, :] - PX[:, None].float()
            dy32 = pix32[ids, 1][None, :] - PY[:, None].float()
            pw32 = (-0.5 * (con32[ids, 0][None] * dx32 * dx32 + con32[ids, 2][None] * dy32 * dy32)
                    - con32[ids, 1][None] * dx32 * dy32)
            al32 = torch.clamp(op32[ids][None] * torch.exp(pw32), max=ALPHA_MAX)
            valid = (pw32 <= 0) & (al32 >= ALPHA_MIN)
            cum = torch.cumprod(torch.where(valid, 1.0 - al32, torch.ones_like(al32)), dim=1)
            incl = valid & (cum >= T_STOP)
            last = torch.where(incl, torch.arange(1, n + 1, device=dev)[None, :], 0).max(dim=1).values
            frag = (((al32 * 255.0 - 1.0).abs() < 2e-3) & (pw32 <= 0) & (cum >= 0.5 * T_STOP)).any(1) \
                | ((valid & ((cum / T_STOP - 1.0).abs() < 2e-3)).any(1))
        dxx = pix[ids, 0][None, :] - PX[:, None].to(dt)
        dyy = pix[ids, 1][None, :] - PY[:, None].to(dt)
        cn = conic[ids]
        power = -0.5 * (cn[:, 0][None] * dxx * dxx + cn[:, 2][None] * dyy * dyy) - cn[:, 1][None] * dxx * dyy
        power = torch.where(incl, power, torch.zeros_like(power))     # keep exp() finite off-support
        raw = opac[ids][None] * torch.exp(power)
        al = raw + (torch.clamp(raw, max=ALPHA_MAX) - raw).detach()   # straight-through clamp
        al = torch.where(incl, al, torch.zeros_like(al))
        one_m = 1.0 - al
        Tcum = torch.cumprod(one_m, dim=1)
        Texc = torch.cat([torch.ones_like(Tcum[:, :1]), Tcum[:, :-1]], dim=1)
        w = al * Texc
        Tfin = Tcum[:, -1]
        col = w @ rgb[ids] + Tfin[:, None] * bg[None, :]
        dep = w @ depth_g[ids]
        acc = w.sum(1)
        color[:, PY, PX] = col.transpose(0, 1)
        depth_img[0, PY, PX] = dep
        alpha_img[0, PY, PX] = acc
        ncontrib[PY, PX] = last.to(torch.int32)
        fragile[PY, PX] = frag

    out = (color, radii, depth_img, alpha_img)
    if return_aux:
        aux = dict(pix=geo["pix"].detach(), depth=geo["depth"].detach(), conic=geo["conic"].detach(),
                   rgb=rgb.detach(), ok=ok, K=K, n_contrib=ncontrib, fragile=fragile,
                   tile_ids=t_sorted, gauss_ids=g_sorted, rect=(x0, y0, x1, y1))
        return out + (aux,)
    return out
