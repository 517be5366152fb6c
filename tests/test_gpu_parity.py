"""-m gpu parity tests: CUDA path (through the public API -> C ABI) vs the oracles."""
import pytest
import torch

import helpers as h

gg = h.gg
pytestmark = pytest.mark.gpu


def _upstream_grads(H, W, seed=1, depth_alpha=True):
    g = torch.Generator().manual_seed(seed)
    Gc = torch.randn(3, H, W, generator=g)
    if not depth_alpha:
        return (Gc, None, None)
    return (Gc, torch.randn(1, H, W, generator=g) * 0.3, torch.randn(1, H, W, generator=g))


def _parity_masked(S, st, grads, max_fragile_frac=1e-2, tol=h.GRAD_TOL, **kw):
    """Forward + backward parity at the north-star tolerances (RGB 1e-4 abs, grads 1e-3 rel) with the upstream
    gradients zeroed on the oracle's threshold pixels for BOTH implementations (helpers.mask_upstream); the
    unmasked error and the worst Gaussians are recorded in the parity report as diagnostics."""
    ref = h.run_c_oracle(S, st, None, **kw)
    masked = h.mask_upstream(grads, ref["fragile"])
    got = h.run_cuda(S, st, masked, **kw)
    ref["grads"] = ref["ctx"].backward(*masked)
    h.assert_images_close(got, ref, max_fragile_frac=max_fragile_frac)
    h.assert_grads_close(got["grads"], ref["grads"], tol=tol)
    # diagnostics: same comparison with the threshold pixels left in the loss
    got_u = h.run_cuda(S, st, grads, **kw)
    ref_u = ref["ctx"].backward(*grads)
    diag = {}
    for k, v in ref_u.items():
        if v is None:
            continue
        e = h.rel_inf(got_u["grads"][k], v)
        diag[k] = dict(rel_inf_unmasked=e)
        if e > tol:
            diag[k]["worst_gaussians(id, err/max)"] = h.worst_gaussians(got_u["grads"], ref_u, k)
    h._note("unmasked", **diag)
    return got, ref


def test_cfg1_parity_vs_c_oracle():
    """BASELINE.json configs[0]: 10k random Gaussians, 512x512, SH degree 3; fwd + bwd."""
    st = gg.scenes.random_cloud(10_000)
    cam = gg.scenes.cfg1_camera()
    S = h.settings_for(cam, st, device=torch.device("cuda:0"))
    grads = _upstream_grads(cam.image_height, cam.image_width)
    got = h.run_cuda(S, st, grads)
    ref = h.run_c_oracle(S, st, grads)
    assert int((got["radii"] != ref["radii"]).sum()) <= 2
    h.assert_images_close(got, ref)
    h.assert_grads_close(got["grads"], ref["grads"])


def test_cfg1_training_style_grads():
    """Only the colour feeds the loss (as in s2/s3): depth/alpha grads arrive as None."""
    st = gg.scenes.random_cloud(4_000)
    cam = gg.scenes.cfg1_camera(384, 272)     # not a multiple of 16 in height
    S = h.settings_for(cam, st, device=torch.device("cuda:0"))
    grads = _upstream_grads(cam.image_height, cam.image_width, depth_alpha=False)
    got = h.run_cuda(S, st, grads)
    ref = h.run_c_oracle(S, st, grads)
    h.assert_images_close(got, ref)
    h.assert_grads_close(got["grads"], ref["grads"])


# ------------------------------------------------------------------------------------ alt paths
def _sh_colors(st, cam, deg=3):
    d = st.means3D - cam.camera_center[None]
    d = d / d.norm(dim=1, keepdim=True)
    return torch.clamp_min(h.torch_oracle.eval_sh_rgb(deg, st.shs, d) + 0.5, 0.0)


def test_colors_precomp_and_cov3d_precomp_paths():
    """--convert_SHs_python / --compute_cov3D_python facade branches (gaussian_renderer/__init__.py:69-89)."""
    st = gg.scenes.random_cloud(3000, seed=4)
    cam = gg.scenes.cfg1_camera(256, 200)
    S = h.settings_for(cam, st, device=torch.device("cuda:0"))
    grads = _upstream_grads(cam.image_height, cam.image_width)
    col = _sh_colors(st, cam)
    cov = h.torch_oracle.covariance3d(st.scales, 1.0, st.rotations)
    got = h.run_cuda(S, st, grads, colors_precomp=col, cov3D_precomp=cov)
    ref = h.run_c_oracle(S, st, grads, colors_precomp=col, cov3D_precomp=cov)
    h.assert_images_close(got, ref)
    h.assert_grads_close(got["grads"], ref["grads"])
    assert got["grads"]["shs"] is None and got["grads"]["scales"] is None


@pytest.mark.parametrize("deg", [0, 1, 2])
def test_lower_active_sh_degree_with_full_coefficient_tensor(deg):
    st = gg.scenes.random_cloud(2500, seed=6)
    st.sh_degree = deg
    cam = gg.scenes.cfg1_camera(208, 176)
    S = h.settings_for(cam, st, device=torch.device("cuda:0"))
    grads = _upstream_grads(cam.image_height, cam.image_width)
    got = h.run_cuda(S, st, grads)
    ref = h.run_c_oracle(S, st, grads)
    h.assert_images_close(got, ref)
    h.assert_grads_close(got["grads"], ref["grads"])
    nb = (deg + 1) ** 2
    assert float(got["grads"]["shs"][:, nb:].abs().max()) == 0.0        # coefficients above D get zero gradient


def test_registration_style_degree0_single_coefficient():
    """s2_registration.py:158 forces sh_degree 0; shs is [N,1,3] there (generic-M kernel path)."""
    st = gg.scenes.random_cloud(3000, seed=8, max_sh_degree=0, sh_degree=0)
    cam = gg.scenes.cfg1_camera(256, 256)
    S = h.settings_for(cam, st, device=torch.device("cuda:0"))
    grads = _upstream_grads(cam.image_height, cam.image_width, depth_alpha=False)
    got = h.run_cuda(S, st, grads)
    ref = h.run_c_oracle(S, st, grads)
    h.assert_images_close(got, ref)
    h.assert_grads_close(got["grads"], ref["grads"])


def test_scale_modifier():
    st = gg.scenes.random_cloud(1500, seed=9)
    cam = gg.scenes.cfg1_camera(160, 160)
    S = h.settings_for(cam, st, device=torch.device("cuda:0"), scale_modifier=1.7)
    grads = _upstream_grads(cam.image_height, cam.image_width)
    got = h.run_cuda(S, st, grads)
    ref = h.run_c_oracle(S, st, grads)
    h.assert_images_close(got, ref)
    h.assert_grads_close(got["grads"], ref["grads"])


# ------------------------------------------------------------------------------------ edge cases
def test_zero_gaussians_gives_zero_image():
    st = gg.scenes.random_cloud(0)
    cam = gg.scenes.cfg1_camera(64, 48)
    S = h.settings_for(cam, st, device=torch.device("cuda:0"))
    got = h.run_cuda(S, st)
    assert got["color"].shape == (3, 48, 64) and float(got["color"].abs().max()) == 0.0
    assert got["radii"].numel() == 0


def test_all_gaussians_behind_camera():
    st = gg.scenes.random_cloud(500, seed=3)
    st.means3D = st.means3D - torch.tensor([0.0, 0.0, 10.0])
    cam = gg.scenes.cfg1_camera(64, 64)
    S = h.settings_for(cam, st, device=torch.device("cuda:0"))
    grads = _upstream_grads(64, 64)
    got = h.run_cuda(S, st, grads)
    assert int((got["radii"] > 0).sum()) == 0
    assert torch.allclose(got["color"], st.bg[:, None, None].expand(3, 64, 64))
    assert float(got["alpha"].abs().max()) == 0.0
    for k in ("means3D", "shs", "opacities", "scales", "rotations"):
        assert float(got["grads"][k].abs().max()) == 0.0


def test_image_sizes_of_the_reference_data():
    """940x1280 (s3_appearance.py:92): neither side a multiple of 16 -> partial tiles."""
    st = gg.scenes.random_cloud(4000, seed=12)
    cam = gg.scenes.cfg1_camera(470, 330)
    S = h.settings_for(cam, st, device=torch.device("cuda:0"))
    grads = _upstream_grads(cam.image_height, cam.image_width)
    got = h.run_cuda(S, st, grads)
    ref = h.run_c_oracle(S, st, grads)
    h.assert_images_close(got, ref)
    h.assert_grads_close(got["grads"], ref["grads"])


def test_dense_tile_exceeding_shared_memory_sort_capacity():
    """> 4096 instances in one tile: the per-tile sort falls back to its global-memory path."""
    n = 6000
    st = gg.scenes.random_cloud(n, seed=13)
    g = torch.Generator().manual_seed(5)
    st.means3D = torch.cat([torch.randn(n, 2, generator=g) * 0.01, 4.0 + torch.rand(n, 1, generator=g) * 2], dim=1)
    st.means3D[:, 2] -= 4.0          # camera T=(0,0,4): view depth 4..6
    st.scales = torch.full((n, 3), 0.004)
    st.opacities = torch.full((n, 1), 0.02)
    cam = gg.scenes.cfg1_camera(64, 64)
    S = h.settings_for(cam, st, device=torch.device("cuda:0"))
    grads = _upstream_grads(64, 64)
    got = h.run_cuda(S, st, grads)
    ref = h.run_c_oracle(S, st, grads)
    off = ref["ctx"].binning()["tile_off"]
    assert int((off[1:] - off[:-1]).max()) > 4096
    h.assert_images_close(got, ref, max_fragile_frac=0.05)
    # 6000 faint splats (o = 0.02 -> alpha within a hair of 1/255 over most of their footprint): nearly every pixel
    # is a threshold pixel, so this scene is compared with the threshold pixels masked out of the loss -- at 1e-3.
    _parity_masked(S, st, grads, max_fragile_frac=1.0)


def test_depth_ties_keep_index_order():
    """Equal fp32 depths in one tile: order must be ascending Gaussian index (stable-sort semantics)."""
    n = 64
    st = gg.scenes.random_cloud(n, seed=14)
    g = torch.Generator().manual_seed(2)
    st.means3D = torch.cat([torch.rand(n, 2, generator=g) * 0.2 - 0.1, torch.zeros(n, 1)], dim=1)   # same depth 4.0
    st.scales = torch.full((n, 3), 0.05)
    st.opacities = torch.full((n, 1), 0.6)
    cam = gg.cameras.make_camera(__import__("numpy").eye(3), [0, 0, 4.0], 300.0, 300.0, 32.0, 32.0, 64, 64)
    S = h.settings_for(cam, st, device=torch.device("cuda:0"))
    got = h.run_cuda(S, st)
    ref = h.run_c_oracle(S, st)
    assert float((got["color"] - ref["color"]).abs().max()) < 1e-4


def test_kat_single_gaussian_on_gpu():
    import numpy as np
    cam = gg.cameras.make_camera(np.eye(3), np.zeros(3), 80.0, 80.0, 32.5, 32.5, 64, 64)
    st = gg.scenes.random_cloud(1, seed=1)
    st.means3D = torch.tensor([[0.0, 0.0, 2.0]])
    st.scales = torch.full((1, 3), 0.05)
    st.rotations = torch.tensor([[1.0, 0, 0, 0]])
    st.opacities = torch.tensor([[0.5]])
    st.bg = torch.tensor([0.1, 0.2, 0.3])
    col = torch.tensor([[0.2, 0.6, 0.9]])
    S = h.settings_for(cam, st, device=torch.device("cuda:0"))
    got = h.run_cuda(S, st, colors_precomp=col)
    assert abs(float(got["alpha"][0, 32, 32]) - 0.5) < 1e-6
    assert torch.allclose(got["color"][:, 32, 32], 0.5 * col[0] + 0.5 * st.bg, atol=1e-6)
    assert abs(float(got["depth"][0, 32, 32]) - 1.0) < 1e-6


# ------------------------------------------------------------------------------------ API contract
def _api_inputs(n=2000, res=(160, 128), seed=21):
    dev = torch.device("cuda:0")
    st = gg.scenes.random_cloud(n, seed=seed).to(dev)
    cam = gg.scenes.cfg1_camera(*res).to(dev)
    S = h.settings_for(cam, st, device=dev)
    return dev, st, cam, S


def test_inplace_mutation_of_color_before_backward_is_allowed():
    """utils/loss_utils.py:45 multiplies the rendered image in place (ssim mask) before backward()."""
    dev, st, cam, S = _api_inputs()
    m3 = st.means3D.clone().requires_grad_(True)
    m2 = torch.zeros_like(m3, requires_grad=True)
    color, radii, depth, alpha = h.dgr.GaussianRasterizer(raster_settings=S)(
        means3D=m3, means2D=m2, shs=st.shs, colors_precomp=None, opacities=st.opacities, scales=st.scales,
        rotations=st.rotations, cov3D_precomp=None)
    gt = torch.rand_like(color)
    mask = (torch.rand(1, *color.shape[1:], device=dev) > 0.3).float()
    loss1 = torch.abs((color - gt) * mask).mean()          # l1_loss(image, gt, mask)   utils/loss_utils.py:17-21
    color *= mask                                           # ssim(): `img1 *= mask`      utils/loss_utils.py:45
    loss2 = (color * gt).mean()
    (loss1 + loss2).backward()
    assert torch.isfinite(m3.grad).all() and float(m3.grad.abs().max()) > 0


def test_no_grad_and_visibility_filter():
    dev, st, cam, S = _api_inputs()
    with torch.no_grad():
        color, radii, depth, alpha = h.dgr.GaussianRasterizer(raster_settings=S)(
            means3D=st.means3D, means2D=torch.zeros_like(st.means3D), shs=st.shs, colors_precomp=None,
            opacities=st.opacities, scales=st.scales, rotations=st.rotations, cov3D_precomp=None)
    assert not color.requires_grad and radii.dtype == torch.int32 and radii.shape == (st.N,)
    assert color.shape == (3, cam.image_height, cam.image_width) and depth.shape[0] == 1 and alpha.shape[0] == 1
    vis = h.dgr.GaussianRasterizer(raster_settings=S).markVisible(st.means3D)
    zview = (torch.cat([st.means3D, torch.ones(st.N, 1, device=dev)], 1) @ cam.world_view_transform)[:, 2]
    assert torch.equal(vis, zview > 0.2)
    assert bool(((radii > 0) <= vis).all())


def test_means2D_side_channel_and_masked_inputs():
    """viewspace_points.grad[:, :2] is read by densification (scene/gaussian_model.py:410-412); inputs may
    come out of boolean-mask indexing (gaussian_renderer/__init__.py:92-100)."""
    dev, st, cam, S = _api_inputs(3000)
    mask = torch.rand(st.N, device=dev) > 0.3
    leaves = [t.clone().requires_grad_(True) for t in (st.means3D, st.shs, st.opacities, st.scales, st.rotations)]
    screenspace = torch.zeros_like(leaves[0], requires_grad=True) + 0
    screenspace.retain_grad()
    color, radii, depth, alpha = h.dgr.GaussianRasterizer(raster_settings=S)(
        means3D=leaves[0][mask], means2D=screenspace[mask], shs=leaves[1][mask], colors_precomp=None,
        opacities=leaves[2][mask], scales=leaves[3][mask], rotations=leaves[4][mask], cov3D_precomp=None)
    color.sum().backward()
    assert screenspace.grad.shape == (st.N, 3)
    assert float(screenspace.grad[~mask].abs().max()) == 0.0
    assert float(screenspace.grad[:, 2].abs().max()) == 0.0
    assert float(screenspace.grad[mask][:, :2].abs().max()) > 0
    assert radii.shape[0] == int(mask.sum())


def test_forward_is_deterministic_and_backward_is_stable():
    dev, st, cam, S = _api_inputs(4000, (256, 256))
    outs = []
    for _ in range(2):
        m3 = st.means3D.clone().requires_grad_(True)
        color, radii, depth, alpha = h.dgr.GaussianRasterizer(raster_settings=S)(
            means3D=m3, means2D=torch.zeros_like(m3), shs=st.shs, colors_precomp=None, opacities=st.opacities,
            scales=st.scales, rotations=st.rotations, cov3D_precomp=None)
        color.square().sum().backward()
        outs.append((color.detach().clone(), m3.grad.clone()))
    assert torch.equal(outs[0][0], outs[1][0])                       # forward: bit-identical
    assert h.rel_inf(outs[0][1], outs[1][1]) < 1e-4                  # backward: float atomics reorder only


def test_grad_bucket_zero_copy_sink():
    """dist.GradBucket: backward writes straight into the flat bucket and autograd adopts the views."""
    from gaussian_garments_b200.dist import GradBucket
    dev, st, cam, S = _api_inputs(2500)
    params = [t.clone().requires_grad_(True) for t in (st.means3D, st.scales, st.rotations, st.opacities, st.shs)]
    plain = [t.clone().requires_grad_(True) for t in (st.means3D, st.scales, st.rotations, st.opacities, st.shs)]

    def run(ps):
        color, *_ = h.dgr.GaussianRasterizer(raster_settings=S)(
            means3D=ps[0], means2D=torch.zeros_like(ps[0]), shs=ps[4], colors_precomp=None, opacities=ps[3],
            scales=ps[1], rotations=ps[2], cov3D_precomp=None)
        (color - 0.3).abs().mean().backward()

    bucket = GradBucket(params, 1)
    try:
        bucket.zero()
        run(params)
        for i, p in enumerate(params):
            assert p.grad is not None and p.grad.data_ptr() == bucket.view(i).data_ptr(), "grad was copied, not adopted"
    finally:
        bucket.unregister()
    run(plain)
    for a, b in zip(params, plain):
        assert h.rel_inf(a.grad, b.grad) < 1e-4


def test_cfg2_full_size_invariants():
    """BASELINE.json configs[1] at full size: size-independent properties (oracle too slow to loop here)."""
    dev = torch.device("cuda:0")
    st = gg.scenes.mesh_bound_state(300_000).to(dev)
    cam = gg.scenes.cfg2_cameras(8)[3].to(dev)
    S = h.settings_for(cam, st, device=dev)
    leaves = [t.clone().requires_grad_(True) for t in (st.means3D, st.shs, st.opacities, st.scales, st.rotations)]

    def render(bg=None, sh=None):
        s = S if bg is None else S._replace(bg=bg)
        return h.dgr.GaussianRasterizer(raster_settings=s)(
            means3D=leaves[0], means2D=torch.zeros_like(leaves[0]), shs=leaves[1] if sh is None else sh,
            colors_precomp=None, opacities=leaves[2], scales=leaves[3], rotations=leaves[4], cov3D_precomp=None)

    color, radii, depth, alpha = render()
    assert torch.isfinite(color).all() and torch.isfinite(depth).all()
    assert float(alpha.min()) >= 0.0 and float(alpha.max()) <= 1.0 + 1e-5
    # background linearity: color(bg) - color(0) == (1 - alpha) * bg  (alpha = 1 - T_final up to rounding)
    c0, _, _, a0 = render(bg=torch.zeros(3, device=dev))
    assert float((color - c0 - (1 - a0) * st.bg[:, None, None]).abs().max()) < 2e-5
    # colour linearity in the DC coefficients (no clamp active when DC is large and positive)
    (color.mean() + depth.mean() * 0.1).backward()
    for t in leaves:
        assert torch.isfinite(t.grad).all()
    assert float(leaves[1].grad[:, 0].abs().sum()) > 0
    # full-size parity against the C oracle (0.5 s on 8 cores): forward + gradients
    gC = torch.full((3, cam.image_height, cam.image_width), 1.0 / (3 * cam.image_height * cam.image_width))
    gD = torch.full((1, cam.image_height, cam.image_width), 0.1 / (cam.image_height * cam.image_width))
    ref = h.run_c_oracle(S, st.to("cpu"), (gC, gD, None))
    got = dict(color=color.detach().cpu(), depth=depth.detach().cpu(), alpha=alpha.detach().cpu())
    assert int((radii.cpu() != ref["radii"]).sum()) <= 3
    h.assert_images_close(got, ref)
    gg_ = dict(means3D=leaves[0].grad.cpu(), shs=leaves[1].grad.cpu(), opacities=leaves[2].grad.cpu(),
               scales=leaves[3].grad.cpu(), rotations=leaves[4].grad.cpu())
    refg = {k: v for k, v in ref["grads"].items() if k in gg_}
    h.assert_grads_close(gg_, refg)


def test_cfg2_full_size_parity_on_every_camera_of_the_ring():
    """BASELINE.json configs[1] at full size (300k mesh-bound Gaussians, 1920x1080, SH degree 3) on ALL 8 ring cameras --
    the views bench.py cycles through -- forward at 1e-4 and gradients at 1e-3 with random upstream gradients for colour,
    depth and alpha (threshold pixels masked out of the loss for both sides)."""
    st = gg.scenes.mesh_bound_state(300_000)
    dev = torch.device("cuda:0")
    for ci, cam in enumerate(gg.scenes.cfg2_cameras(8)):
        S = h.settings_for(cam, st, device=dev)
        ref = h.run_c_oracle(S, st, None)
        grads = h.mask_upstream(_upstream_grads(cam.image_height, cam.image_width, seed=40 + ci), ref["fragile"])
        got = h.run_cuda(S, st, grads)
        ref["grads"] = ref["ctx"].backward(*grads)
        assert int((got["radii"] != ref["radii"]).sum()) <= 3, ci
        h.assert_images_close(got, ref)
        h.assert_grads_close(got["grads"], ref["grads"])
        ref["ctx"].close()


def test_stage1_geometry_is_bit_identical_to_oracle():
    """project_kernel is built without FMA contraction: xy / depth / conic / radius / tile rectangle of every
    Gaussian equal the C oracle's bit for bit, so all geometry-derived discrete decisions agree by construction."""
    import ctypes as C
    from gaussian_garments_b200 import _capi
    from gaussian_garments_b200.rasterizer import _view_struct
    lib = _capi.load()
    dev = torch.device("cuda:0")
    st_cpu = gg.scenes.random_cloud(20_000, seed=33)
    cam = gg.scenes.cfg1_camera(640, 360)
    S = h.settings_for(cam, st_cpu, device=dev)
    st = st_cpu.to(dev)
    N = st.N
    view = _view_struct(S, N, 16)
    keep = [st.means3D, st.shs, st.opacities, st.scales, st.rotations, S.bg, S.viewmatrix, S.projmatrix, S.campos]
    inp = _capi.GGInputs(st.means3D.data_ptr(), st.shs.data_ptr(), None, st.opacities.data_ptr(), st.scales.data_ptr(),
                         st.rotations.data_ptr(), None, S.bg.data_ptr(), S.viewmatrix.data_ptr(),
                         S.projmatrix.data_ptr(), S.campos.data_ptr())
    gb, tb, ib = C.c_size_t(), C.c_size_t(), C.c_size_t()
    assert lib.gg_forward_workspace_bytes(C.byref(view), C.byref(gb), C.byref(tb), C.byref(ib)) == 0
    geom = torch.empty(gb.value, dtype=torch.uint8, device=dev)
    tile = torch.empty(tb.value, dtype=torch.uint8, device=dev)
    radii = torch.empty(N, dtype=torch.int32, device=dev)
    sp = torch.cuda.current_stream().cuda_stream
    _capi.check(lib.gg_forward_project(C.byref(view), C.byref(inp), geom.data_ptr(), tile.data_ptr(), radii.data_ptr(),
                                       None, 0, sp), "project")
    _capi.check(lib.gg_forward_color(C.byref(view), C.byref(inp), geom.data_ptr(), radii.data_ptr(), 0, sp), "color")
    xy = torch.empty(N, 2, device=dev); depth = torch.empty(N, device=dev); con = torch.empty(N, 4, device=dev)
    rgb = torch.empty(N, 3, device=dev); rect = torch.empty(N, 2, dtype=torch.int32, device=dev)
    _capi.check(lib.gg_debug_read_geom(C.byref(view), geom.data_ptr(), xy.data_ptr(), depth.data_ptr(), con.data_ptr(),
                                       rgb.data_ptr(), rect.data_ptr(), 0, sp), "read_geom")
    torch.cuda.synchronize()
    ref = h.run_c_oracle(S, st_cpu)
    g = ref["ctx"].geom()
    vis = ref["radii"] > 0
    assert torch.equal(radii.cpu(), ref["radii"])
    assert torch.equal(xy.cpu()[vis], g["xy"][vis])
    assert torch.equal(depth.cpu()[vis], g["depth"][vis])
    assert torch.equal(con.cpu()[vis], g["conic_opacity"][vis])
    r = rect.cpu()
    unpacked = torch.stack([r[:, 0] & 0xFFFF, (r[:, 0] >> 16) & 0xFFFF, r[:, 1] & 0xFFFF, (r[:, 1] >> 16) & 0xFFFF], 1)
    # the kernel's rectangle is the upstream 3-sigma rectangle (== oracle) intersected with the bounding box of the
    # alpha >= 1/255 ellipse (exact tile culling): a sub-rectangle, and a strict one for a good share of the splats
    u, o = unpacked[vis].int(), g["rect"][vis]
    nonempty = (u[:, 2] > u[:, 0]) & (u[:, 3] > u[:, 1])
    assert bool((u[nonempty, 0] >= o[nonempty, 0]).all() and (u[nonempty, 1] >= o[nonempty, 1]).all())
    assert bool((u[nonempty, 2] <= o[nonempty, 2]).all() and (u[nonempty, 3] <= o[nonempty, 3]).all())
    area = lambda q: ((q[:, 2] - q[:, 0]) * (q[:, 3] - q[:, 1])).clamp_min(0).sum()
    assert int(area(u)) <= int(area(o))
    assert float((rgb.cpu()[vis] - g["rgb"][vis]).abs().max()) < 2e-6      # SH sum order differs: continuous only
    del keep


# ------------------------------------------------------------------------------------ lazy fused forward path
@pytest.fixture
def lazy_path(monkeypatch):
    monkeypatch.setenv("GG_FWD_PATH", "lazy")
    yield
    monkeypatch.delenv("GG_FWD_PATH", raising=False)


def test_lazy_forward_path_matches_oracle_cfg1(lazy_path):
    st = gg.scenes.random_cloud(10_000)
    cam = gg.scenes.cfg1_camera()
    S = h.settings_for(cam, st, device=torch.device("cuda:0"))
    grads = _upstream_grads(cam.image_height, cam.image_width)
    got = h.run_cuda(S, st, grads)
    ref = h.run_c_oracle(S, st, grads)
    h.assert_images_close(got, ref)
    h.assert_grads_close(got["grads"], ref["grads"])


def _dense_tile_state(n, same_depth=False, opacity=0.02, seed=5):
    st = gg.scenes.random_cloud(n, seed=13)
    g = torch.Generator().manual_seed(seed)
    z = torch.zeros(n, 1) if same_depth else torch.rand(n, 1, generator=g) * 2
    st.means3D = torch.cat([torch.randn(n, 2, generator=g) * 0.01, z], dim=1)      # camera T=(0,0,4): depth 4..6
    st.scales = torch.full((n, 3), 0.004)
    st.opacities = torch.full((n, 1), opacity)
    return st


@pytest.mark.parametrize("same_depth", [False, True])
def test_lazy_forward_path_bucketed_tiles(lazy_path, same_depth):
    """> 2048 instances per tile: depth-bucketed lazy sort; all-equal depths force the degenerate (L2 sort) branch."""
    st = _dense_tile_state(6000, same_depth=same_depth)
    cam = gg.scenes.cfg1_camera(64, 64)
    S = h.settings_for(cam, st, device=torch.device("cuda:0"))
    grads = _upstream_grads(64, 64)
    _parity_masked(S, st, grads, max_fragile_frac=1.0)


def test_lazy_path_equals_tma_path_when_tiles_saturate(monkeypatch):
    """Opaque dense tile: the lazy path stops after the first buckets; result must equal the full-sort path."""
    st = _dense_tile_state(9000, opacity=0.6)
    cam = gg.scenes.cfg1_camera(64, 64)
    S = h.settings_for(cam, st, device=torch.device("cuda:0"))
    grads = _upstream_grads(64, 64)
    monkeypatch.setenv("GG_FWD_PATH", "tma")
    a = h.run_cuda(S, st, grads)
    monkeypatch.setenv("GG_FWD_PATH", "lazy")
    b = h.run_cuda(S, st, grads)
    monkeypatch.delenv("GG_FWD_PATH")
    assert torch.equal(a["color"], b["color"]) and torch.equal(a["alpha"], b["alpha"]) and torch.equal(a["depth"], b["depth"])
    for k in ("means3D", "shs", "opacities", "scales", "rotations"):
        assert h.rel_inf(a["grads"][k], b["grads"][k]) < 1e-4


# ------------------------------------------------------------------------------------ N1: fused mesh binding
@pytest.mark.parametrize("avatar,remembered", [(False, False), (True, False), (False, True), (True, True)])
def test_fused_mesh_binding_matches_torch_chain(avatar, remembered):
    """gg_mesh_bind_forward_ex/backward_ex vs autograd (fp64) through oracle/mesh_chain.py -- the golden-pinned
    restatement of scene/mesh_gaussian_model.py:90-128 and, with `avatar`, of the barycentric anchor of
    scene/avatar_gaussian_model.py:140-159 (get_xyz AND get_final_xyz); `remembered` = face_scaling_remembered branch
    (scene/mesh_gaussian_model.py:98-110).  Includes the gradient to mesh.v."""
    from oracle import mesh_chain as mc
    dev = torch.device("cuda:0")
    g = torch.Generator().manual_seed(11)
    m = gg.scenes.MeshBoundGaussians(n_faces_around=40, n_along=12, per_face=6, seed=5)
    m.mesh_v = m.mesh_v + 0.005 * torch.randn(m.mesh_v.shape, generator=g)
    m._rotation = torch.randn(m._rotation.shape, generator=g)
    N, F = m.binding.shape[0], m.mesh_f.shape[0]
    bc = None
    if avatar:
        bc = torch.rand(N, 3, generator=g) + 0.05
        bc = bc / bc.sum(1, keepdim=True)
    rem = (0.8 + 0.4 * torch.rand(F, 1, generator=g)) * 0.01 if remembered else None
    local_final = m._xyz + 0.02 * torch.randn(N, 3, generator=g)          # AvatarNet: local_xyz = _xyz + offset
    G = [torch.randn(N, 3, generator=g), torch.randn(N, 3, generator=g), torch.randn(N, 4, generator=g)]

    for final in ((False, True) if avatar else (False,)):
        lx_src = local_final if final else m._xyz
        a = [t.detach().clone().to(dev).requires_grad_(True) for t in (m.mesh_v, lx_src, m._scaling, m._rotation)]
        xyz, sc, ro = gg.bind_to_mesh(a[0], m.mesh_f.to(dev), m.binding.to(dev), a[1], a[2], a[3],
                                      barycentric=None if bc is None else [bc[:, k].to(dev) for k in range(3)],
                                      face_scaling_remembered=None if rem is None else rem.to(dev))
        ((xyz * G[0].to(dev)).sum() + (sc * G[1].to(dev)).sum() + (ro * G[2].to(dev)).sum()).backward()

        b = [t.detach().double().requires_grad_(True) for t in (m.mesh_v, lx_src, m._scaling, m._rotation)]
        ref = mc.MeshChain(b[0], m.mesh_f, m.binding, b[1], b[2], b[3], gs_bc=None if bc is None else bc.double(),
                           local_xyz=b[1])
        ref.update_face_coor()
        if rem is not None:
            ref.face_scaling_remembered = rem.double()
        r_xyz = ref.get_final_xyz if final else ref.get_xyz
        ((r_xyz * G[0].double()).sum() + (ref.get_scaling * G[1].double()).sum() + (ref.get_rotation * G[2].double()).sum()).backward()

        assert torch.allclose(xyz.cpu(), r_xyz.float(), atol=2e-6)
        assert torch.allclose(sc.cpu(), ref.get_scaling.float(), atol=1e-7, rtol=1e-5)
        assert torch.allclose(ro.cpu(), ref.get_rotation.float(), atol=2e-6)
        for name, x, y in zip(("mesh_v", "_xyz", "_scaling", "_rotation"), a, b):
            assert h.rel_inf(x.grad.cpu(), y.grad.float()) < 1e-3, (name, final)


def test_fused_mesh_binding_provider_tracks_rebinding_and_rejects_bad_indices():
    """the int32 copies of mesh.f / binding must follow a re-assigned or in-place edited `binding` of the
    SAME length (prune + clone, scene/mesh_gaussian_model.py:130-208); out-of-range indices raise like torch indexing."""
    dev = torch.device("cuda:0")
    m = gg.scenes.MeshBoundGaussians(n_faces_around=24, n_along=6, per_face=2, seed=3).to(dev)
    prov = gg.FusedMeshBinding(m)
    x0, _, _ = prov.world()
    m.update_face_coor()
    assert torch.allclose(x0, m.get_xyz, atol=2e-6)
    m.binding = m.binding.flip(0).contiguous()                 # new tensor, same length
    x1, _, _ = prov.world()
    assert torch.allclose(x1, m.get_xyz, atol=2e-6) and not torch.allclose(x1, x0, atol=1e-4)
    m.binding[:10] = 0                                          # in-place edit (version bump)
    x2, _, _ = prov.world()
    assert torch.allclose(x2, m.get_xyz, atol=2e-6)
    bad = m.binding.clone()
    bad[3] = m.mesh_f.shape[0]                                  # one past the last face
    with pytest.raises(IndexError):
        gg.bind_to_mesh(m.mesh_v, m.mesh_f, bad, m._xyz, m._scaling, m._rotation)
    with pytest.raises(IndexError):
        gg.bind_to_mesh(m.mesh_v, m.mesh_f, m.binding[:-1], m._xyz, m._scaling, m._rotation)


def test_fused_mesh_binding_feeds_the_rasterizer():
    """End to end: mesh.v gradient of a render through fused binding == through the torch chain."""
    dev = torch.device("cuda:0")
    m = gg.scenes.MeshBoundGaussians(n_faces_around=60, n_along=20, per_face=6, sh_degree=0, max_sh_degree=0, seed=9).to(dev)
    cam = gg.scenes.ring_cameras(4, width=320, height=240)[1].to(dev)
    S = h.dgr.GaussianRasterizationSettings(image_height=240, image_width=320, tanfovx=cam.tanfovx, tanfovy=cam.tanfovy,
                                            bg=m.bg, scale_modifier=1.0, viewmatrix=cam.world_view_transform,
                                            projmatrix=cam.full_proj_transform, sh_degree=0, campos=cam.camera_center,
                                            prefiltered=False, debug=False)
    gt = torch.rand(3, 240, 320, device=dev)
    grads = []
    for fused in (False, True):
        m.mesh_v = m.mesh_v.detach().clone().requires_grad_(True)
        if fused:
            xyz, sc, ro = gg.FusedMeshBinding(m).world()
        else:
            m.update_face_coor()
            xyz, sc, ro = m.get_xyz, m.get_scaling, m.get_rotation
        color, *_ = h.dgr.GaussianRasterizer(raster_settings=S)(
            means3D=xyz, means2D=torch.zeros_like(xyz), shs=m.get_features, colors_precomp=None, opacities=m.get_opacity,
            scales=sc, rotations=ro, cov3D_precomp=None)
        (color - gt).abs().mean().backward()
        grads.append(m.mesh_v.grad.clone())
    assert float(grads[0].abs().max()) > 0
    assert h.rel_inf(grads[1], grads[0]) < 2e-3


# ------------------------------------------------------------------------------------ N2: fused photometric loss
@pytest.mark.parametrize("shape,use_mask", [((3, 37, 45), True), ((3, 40, 52), False), ((3, 270, 480), True)])
def test_fused_photometric_loss_matches_oracle(shape, use_mask):
    from oracle import loss_oracle as lo
    dev = torch.device("cuda:0")
    g = torch.Generator().manual_seed(4)
    img = torch.rand(*shape, generator=g)
    gt = (img + 0.2 * torch.randn(*shape, generator=g)).clamp(0, 1)
    mask = (torch.rand(1, *shape[1:], generator=g) > 0.3).float() if use_mask else None
    x = img.to(dev).requires_grad_(True)
    total, l1, ss = gg.photometric_loss(x, gt.to(dev), None if mask is None else mask.to(dev), 0.2)
    (total * 1.7).backward()                       # non-trivial upstream scalar stays on the device
    xr = img.double().requires_grad_(True)
    ref_total = lo.total_loss(xr, gt.double(), None if mask is None else mask.double(), 0.2)
    (ref_total * 1.7).backward()
    assert abs(float(l1) - float(lo.l1_loss(img.double(), gt.double(), None if mask is None else mask.double()))) < 1e-6
    assert abs(float(ss) - float(lo.ssim(img.double(), gt.double(), None if mask is None else mask.double()))) < 2e-5
    assert abs(float(total) - float(ref_total)) < 2e-5
    assert h.rel_inf(x.grad.cpu(), xr.grad.float()) < 1e-3


def test_photometric_loss_inplace_mask_semantics():
    """inplace_mask=True: the reference's call sequence to the letter -- l1_loss on the unmasked tensors, then ssim()
    multiplies image and gt by the mask in place (utils/loss_utils.py:44-46).  Soft mask, so mask*mask != mask."""
    from oracle import loss_oracle as lo
    dev = torch.device("cuda:0")
    g = torch.Generator().manual_seed(12)
    shape = (3, 75, 132)
    img = torch.rand(*shape, generator=g)
    gt = (img + 0.2 * torch.randn(*shape, generator=g)).clamp(0, 1)
    mask = torch.rand(1, *shape[1:], generator=g)
    leaf = img.to(dev).requires_grad_(True)
    x = leaf * 1.0                                  # non-leaf, like the rasterizer's output
    gt_dev = gt.to(dev)
    total, l1, ss = gg.photometric_loss(x, gt_dev, mask.to(dev), 0.2, inplace_mask=True)
    (total * 0.9).backward()
    assert torch.allclose(x.detach().cpu(), img * mask, atol=1e-7) and torch.allclose(gt_dev.cpu(), gt * mask, atol=1e-7)
    xr = img.double().requires_grad_(True)
    l1_r = lo.l1_loss(xr, gt.double(), mask.double())
    ss_r = lo.ssim(xr * mask.double(), gt.double() * mask.double(), None)
    tot_r = l1_r * 0.8 + 1.0 - ss_r * 0.2
    (tot_r * 0.9).backward()
    assert abs(float(l1) - float(l1_r)) < 1e-6 and abs(float(ss) - float(ss_r)) < 2e-5 and abs(float(total) - float(tot_r)) < 2e-5
    assert h.rel_inf(leaf.grad.cpu(), xr.grad.float()) < 1e-3


def test_fused_photometric_loss_golden_from_reference():
    """Directly against values the reference's own l1_loss/ssim produced (tests/golden/loss.npz)."""
    import numpy as np, os
    z = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "loss.npz"))
    dev = torch.device("cuda:0")
    for tag in ("a", "b"):
        x = torch.tensor(z[f"{tag}_img"]).to(dev).requires_grad_(True)
        gt = torch.tensor(z[f"{tag}_gt"]).to(dev)
        mask = torch.tensor(z[f"{tag}_mask"]).to(dev) if f"{tag}_mask" in z else None
        total, l1, ss = gg.photometric_loss(x, gt, mask, 0.2)
        total.backward()
        assert abs(float(l1) - float(z[f"{tag}_l1"])) < 1e-6 and abs(float(ss) - float(z[f"{tag}_ssim"])) < 2e-5
        ref = torch.tensor(z[f"{tag}_grad"])
        assert float((x.grad.cpu() - ref).abs().max()) <= 1e-3 * float(ref.abs().max())


def test_debug_mode_synchronises_and_matches():
    """raster_settings.debug=True (arguments/__init__.py:69): sync + error check after every launch; same results."""
    st = gg.scenes.random_cloud(2000, seed=17)
    cam = gg.scenes.cfg1_camera(160, 128)
    grads = _upstream_grads(128, 160)
    a = h.run_cuda(h.settings_for(cam, st, device=torch.device("cuda:0"), debug=False), st, grads)
    b = h.run_cuda(h.settings_for(cam, st, device=torch.device("cuda:0"), debug=True), st, grads)
    assert torch.equal(a["color"], b["color"]) and torch.equal(a["radii"], b["radii"])
    assert h.rel_inf(a["grads"]["means3D"], b["grads"]["means3D"]) < 1e-4


def test_concurrent_streams_and_threads():
    """Re-entrancy: two host threads rendering different scenes on their own streams get the single-threaded results."""
    import threading
    dev = torch.device("cuda:0")
    jobs = []
    for seed in (3, 4):
        st = gg.scenes.random_cloud(3000, seed=seed)
        cam = gg.scenes.cfg1_camera(192, 160)
        S = h.settings_for(cam, st, device=dev)
        jobs.append((st, S, h.run_cuda(S, st)["color"]))
    out = [None, None]

    def work(k):
        with torch.cuda.stream(torch.cuda.Stream(device=dev)):
            for _ in range(3):
                out[k] = h.run_cuda(jobs[k][1], jobs[k][0])["color"]
            torch.cuda.current_stream().synchronize()

    ts = [threading.Thread(target=work, args=(k,)) for k in range(2)]
    [t.start() for t in ts]
    [t.join() for t in ts]
    for k in range(2):
        assert torch.equal(out[k], jobs[k][2])


# ------------------------------------------------------------------------------------ more full-config parity
def test_cfg1_parity_vs_torch_autograd_oracle():
    """BASELINE configs[0] against the PyTorch oracle whose gradients come from autograd (independent of every
    hand-derived backward formula): RGB <= 1e-4 abs, grads <= 1e-3 rel."""
    st = gg.scenes.random_cloud(10_000)
    cam = gg.scenes.cfg1_camera()
    S = h.settings_for(cam, st, device=torch.device("cuda:0"))
    grads = _upstream_grads(cam.image_height, cam.image_width)
    got = h.run_cuda(S, st, grads)
    leaf = lambda t: t.detach().clone().requires_grad_(True)
    m3, sh, op, sc, ro = (leaf(t) for t in (st.means3D, st.shs, st.opacities, st.scales, st.rotations))
    m2 = torch.zeros_like(m3, requires_grad=True)
    color, radii, depth, alpha, aux = h.torch_oracle.rasterize(h.cpu_settings(S), m3, m2, sh, None, op, sc, ro, None,
                                                               dtype=torch.float32, return_aux=True)
    ((color * grads[0]).sum() + (depth * grads[1]).sum() + (alpha * grads[2]).sum()).backward()
    ref = dict(color=color.detach(), depth=depth.detach(), alpha=alpha.detach(), fragile=aux["fragile"])
    assert int((got["radii"] != radii).sum()) <= 2
    h.assert_images_close(got, ref, max_fragile_frac=0.05)
    refg = dict(means3D=m3.grad, means2D=m2.grad, shs=sh.grad, opacities=op.grad, scales=sc.grad, rotations=ro.grad)
    h.assert_grads_close(got["grads"], refg)


def test_cfg4_style_registration_scene_parity():
    """BASELINE configs[3] shape: 150k mesh-bound Gaussians, SH degree 0 with a [N,1,3] tensor, 1280x720."""
    st = gg.scenes.mesh_bound_state(150_000, sh_degree=0, max_sh_degree=0)
    cam = gg.scenes.ring_cameras(32, width=1280, height=720)[5]
    S = h.settings_for(cam, st, device=torch.device("cuda:0"))
    grads = _upstream_grads(720, 1280, depth_alpha=False)
    got, ref = _parity_masked(S, st, grads)       # 1e-3; threshold pixels (alpha ~ 1/255 flips) masked out of the loss
    assert int((got["radii"] != ref["radii"]).sum()) <= 3


def test_cfg5_style_dense_scene_parity():
    """BASELINE configs[4] shape at a size the oracle sorts in seconds: stress cloud (large, overlapping splats) ->
    thousands of instances per tile, automatically routed through the lazy bucketed forward."""
    st = gg.scenes.stress_cloud(120_000)
    cam = gg.scenes.ring_cameras(8, width=1280, height=720)[2]
    S = h.settings_for(cam, st, device=torch.device("cuda:0"))
    grads = _upstream_grads(720, 1280)
    got, ref = _parity_masked(S, st, grads, max_fragile_frac=0.05)
    off = ref["ctx"].binning()["tile_off"]
    assert int((off[1:] - off[:-1]).max()) > 4096          # dense enough to take the lazy path
    assert int((got["radii"] != ref["radii"]).sum()) <= 3


# ------------------------------------------------------------------------------------ round 2: index-level, sinks, hints
def test_binning_index_level_parity_vs_oracle():
    """Rows a6-a8 at index level (north_star: bit-exact for index work): per tile, the sorted Gaussian-id list the
    kernels blend from must be the oracle's (tile, depth, id)-ordered list with exactly the instances removed that
    the exact-culling predicate proves invisible (no pixel of the tile can reach alpha >= 1/255), in the same order;
    tile offsets must be the running sums of those list lengths."""
    import ctypes as C
    import numpy as np
    from gaussian_garments_b200 import _capi, rasterizer
    st = gg.scenes.random_cloud(10_000)
    cam = gg.scenes.cfg1_camera()
    dev = torch.device("cuda:0")
    S = h.settings_for(cam, st, device=dev)
    cap = {}
    rasterizer.DEBUG_CAPTURE = cap
    try:
        h.run_cuda(S, st)
    finally:
        rasterizer.DEBUG_CAPTURE = None
    lib = _capi.load()
    gx, gy = (cam.image_width + 15) // 16, (cam.image_height + 15) // 16
    T = gx * gy
    offs = torch.zeros(T + 1, dtype=torch.int32, device=dev)
    ids = torch.zeros(max(cap["capacity"], 1), dtype=torch.int32, device=dev)
    sp = torch.cuda.current_stream().cuda_stream
    _capi.check(lib.gg_debug_read_binning(C.byref(cap["view"]), cap["tile_ws"].data_ptr(), cap["record_ws"].data_ptr(),
                                          cap["capacity"], offs.data_ptr(), ids.data_ptr(), 0, sp), "read_binning")
    torch.cuda.synchronize()
    offs, ids = offs.cpu().numpy().astype(np.int64), ids.cpu().numpy().astype(np.int64)
    assert offs[0] == 0 and offs[T] == cap["K"] and np.all(np.diff(offs) >= 0)

    ref = h.run_c_oracle(S, st)
    b = ref["ctx"].binning()
    o_off, o_inst = b["tile_off"].numpy(), b["inst"].numpy().astype(np.int64)
    g = ref["ctx"].geom()
    xy, con = g["xy"].numpy().astype(np.float64), g["conic_opacity"].numpy().astype(np.float64)
    assert o_off[T] >= offs[T]
    dropped_total, max_alpha_dropped = 0, 0.0
    lx, ly = np.meshgrid(np.arange(16), np.arange(16))
    for t in range(T):
        mine = ids[offs[t]:offs[t + 1]]
        full = o_inst[o_off[t]:o_off[t + 1]]
        if mine.size == full.size:
            assert np.array_equal(mine, full), f"tile {t}: order differs"
            continue
        # subsequence check (order preserved) + the dropped entries are provably invisible in this tile
        keep = np.zeros(full.size, dtype=bool)
        j = 0
        for i, gid in enumerate(full):
            if j < mine.size and mine[j] == gid:
                keep[i] = True
                j += 1
        assert j == mine.size, f"tile {t}: kernel list is not an order-preserving subsequence of the oracle list"
        drop = full[~keep]
        dropped_total += drop.size
        px = (t % gx) * 16 + lx.reshape(1, -1)
        py = (t // gx) * 16 + ly.reshape(1, -1)
        dx = xy[drop, 0:1] - px
        dy = xy[drop, 1:2] - py
        power = -0.5 * (con[drop, 0:1] * dx * dx + con[drop, 2:3] * dy * dy) - con[drop, 1:2] * dx * dy
        inside = (px < cam.image_width) & (py < cam.image_height)
        a = np.where((power <= 0) & inside, con[drop, 3:4] * np.exp(np.minimum(power, 0.0)), 0.0)
        max_alpha_dropped = max(max_alpha_dropped, float(a.max()) if a.size else 0.0)
    assert max_alpha_dropped < 1.0 / 255.0, f"a culled instance reaches alpha {max_alpha_dropped}"
    h._note("binning", K_kernel=int(offs[T]), K_oracle=int(o_off[T]), culled=int(dropped_total),
            max_alpha_of_culled=max_alpha_dropped)
    assert dropped_total == int(o_off[T] - offs[T])


def test_two_backwards_accumulate_exactly_with_registered_sinks():
    """with a GradBucket registered, a second backward before .grad is reset must ACCUMULATE
    (g1 + g2), not alias the first result; a freed-and-reused address must not inherit a sink."""
    from gaussian_garments_b200.dist import GradBucket
    dev, st, cam, S = _api_inputs(2500)
    cams = [gg.scenes.cfg1_camera(160, 128).to(dev), gg.scenes.cfg1_camera(160, 128, fov_deg=40.0).to(dev)]
    names = ("means3D", "scales", "rotations", "opacities", "shs")

    def leaves():
        return [getattr(st, k).clone().requires_grad_(True) for k in names]

    def render(ps, cam_):
        S_ = h.settings_for(cam_, st, device=dev)
        color, *_ = h.dgr.GaussianRasterizer(raster_settings=S_)(
            means3D=ps[0], means2D=torch.zeros_like(ps[0]), shs=ps[4], colors_precomp=None, opacities=ps[3],
            scales=ps[1], rotations=ps[2], cov3D_precomp=None)
        return (color - 0.3).abs().mean()

    plain = leaves()
    for c in cams:
        render(plain, c).backward()                      # reference: autograd accumulates g1 + g2
    params = leaves()
    bucket = GradBucket(params, 1)
    try:
        bucket.zero()
        for c in cams:
            render(params, c).backward()
        for i, (a, b) in enumerate(zip(params, plain)):
            assert a.grad.data_ptr() == bucket.view(i).data_ptr(), "first backward was not adopted zero-copy"
            assert h.rel_inf(a.grad, b.grad) < 1e-4, names[i]
        # one loss over two renders of the same leaves (both backward nodes run before AccumulateGrad)
        bucket.zero()
        (render(params, cams[0]) + render(params, cams[1])).backward()
        for a, b in zip(params, plain):
            assert h.rel_inf(a.grad, b.grad) < 1e-4
    finally:
        bucket.unregister()


def test_hinted_forward_equals_synchronous_forward_and_recovers_from_overflow():
    """The sync-free forward (instance capacity from the previous call, K checked after everything is queued) must
    give bit-identical images to the first, synchronous call -- also when the hint is far too small (overflow retry)."""
    from gaussian_garments_b200 import rasterizer
    dev, st, cam, S = _api_inputs(6000, (320, 256), seed=5)
    rasterizer._hints.clear()
    a = h.run_cuda(S, st, _upstream_grads(256, 320))
    n0 = dict(rasterizer.STATS)
    b = h.run_cuda(S, st, _upstream_grads(256, 320))
    assert rasterizer.STATS["hinted"] == n0["hinted"] + 1 and rasterizer.STATS["overflow_retries"] == n0["overflow_retries"]
    for k in ("color", "depth", "alpha"):
        assert torch.equal(a[k], b[k])
    for k in ("means3D", "shs", "opacities", "scales", "rotations"):
        assert h.rel_inf(b["grads"][k], a["grads"][k]) < 1e-4
    for key in list(rasterizer._hints):
        rasterizer._hints[key] = [64, 16]                # absurdly small capacity -> overflow -> retry
    c = h.run_cuda(S, st, _upstream_grads(256, 320))
    assert rasterizer.STATS["overflow_retries"] == n0["overflow_retries"] + 1
    for k in ("color", "depth", "alpha"):
        assert torch.equal(a[k], c[k])
    for k in ("means3D", "shs", "opacities", "scales", "rotations"):
        assert h.rel_inf(c["grads"][k], a["grads"][k]) < 1e-4


# ------------------------------------------------------------------------------------ the reference facade, replayed
def _facade_inputs(sc, z, dev):
    """Rebuild on the GPU what /root/reference/gaussian_renderer/__init__.py hands to the rasterizer in scenario `sc`
    (the trace was recorded from the unmodified facade: tests/golden/make_facade_trace.py)."""
    T = lambda k: torch.tensor(z[k]).to(dev)
    grad_mode = sc["fn"] == "render"
    leaf = (lambda t: t.clone().requires_grad_(True)) if grad_mode else (lambda t: t)
    xyz, log_s, rot, shs = leaf(T("means3D")), leaf(torch.log(T("scales"))), leaf(T("rotations")), leaf(T("shs"))
    logit_o = leaf(torch.logit(T("opacities").clamp(1e-4, 1 - 1e-4)))
    opt = sc["options"]
    with torch.set_grad_enabled(grad_mode):
        screenspace = torch.zeros_like(xyz, requires_grad=True) + 0                       # facade :29 / :132
        if grad_mode:
            screenspace.retain_grad()
        means3D = (xyz + 0.01) * 1.0 + 0.5 if opt.get("avatar") else xyz                  # get_final_xyz vs get_xyz (:56)
        kw = dict(means3D=means3D, means2D=screenspace, opacities=torch.sigmoid(logit_o), shs=None, colors_precomp=None,
                  scales=None, rotations=None, cov3D_precomp=None)
        if sc["pipe"]["compute_cov3D_python"]:
            kw["cov3D_precomp"] = T(f"{sc['name']}__cov3D_precomp")                       # computed by the reference (:70)
            if grad_mode:
                kw["cov3D_precomp"] = kw["cov3D_precomp"] + 0 * log_s.sum()               # non-leaf, requires grad
        else:
            kw["scales"], kw["rotations"] = torch.exp(log_s), torch.nn.functional.normalize(rot)
        if opt.get("override_color"):
            kw["colors_precomp"] = T("override_color")
        elif sc["pipe"]["convert_SHs_python"]:
            kw["colors_precomp"] = T(f"{sc['name']}__colors_precomp") + 0 * shs.sum()    # eval_sh by the reference (:81-85)
        elif opt.get("override_shs"):
            kw["shs"] = T("override_shs")
        else:
            kw["shs"] = shs * 1.5 if opt.get("avatar") else shs                           # pc.shs vs get_features (:87)
        if opt.get("vis_mask"):
            m = T("vis_mask")
            kw = {k: (v[m] if v is not None else None) for k, v in kw.items()}           # :92-100
    return kw, screenspace, dict(xyz=xyz, log_s=log_s, rot=rot, shs=shs, logit_o=logit_o)


def test_reference_facade_trace_replay():
    """Rows a1/a2: every call the reference's render() / doll_render() makes (recorded from the unmodified facade with a
    recording rasterizer) goes through THIS package with the same keywords, dtypes, shapes, strides and autograd
    structure, and what comes back satisfies the facade's return contract (gaussian_renderer/__init__.py:115-122, :221)."""
    import json, os
    import numpy as np
    here = os.path.dirname(os.path.abspath(__file__))
    trace = json.load(open(os.path.join(here, "golden", "facade_trace.json")))
    z = np.load(os.path.join(here, "golden", "facade_trace.npz"))
    dev = torch.device("cuda:0")
    cam = gg.scenes.cfg1_camera(trace["W"], trace["H"]).to(dev)
    N, H, W = trace["N"], trace["H"], trace["W"]
    bg = torch.tensor(z["bg"]).to(dev)
    images = {}
    for sc in trace["scenarios"]:
        kw, screenspace, leaves = _facade_inputs(sc, z, dev)
        # the inputs we rebuilt are what the facade passed: same None pattern, shapes, dtypes, strides, autograd flags
        for k, meta in sc["call"].items():
            if meta is None:
                assert kw[k] is None, (sc["name"], k)
                continue
            t = kw[k]
            assert list(t.shape) == meta["shape"] and str(t.dtype) == "torch." + meta["dtype"], (sc["name"], k)
            assert list(t.stride()) == meta["stride"] and t.requires_grad == meta["requires_grad"], (sc["name"], k)
        # settings: exactly the recorded keyword set, scalars as recorded, tensors from the camera / state
        sk = dict(sc["settings_scalars"])
        assert abs(sk["tanfovx"] - cam.tanfovx) < 1e-12 and abs(sk["tanfovy"] - cam.tanfovy) < 1e-12
        sk.update(bg=bg, viewmatrix=cam.world_view_transform, projmatrix=cam.full_proj_transform,
                  campos=cam.camera_center)
        settings = h.dgr.GaussianRasterizationSettings(**{k: sk[k] for k in sc["settings_keys"]})
        rasterizer = h.dgr.GaussianRasterizer(raster_settings=settings)
        with torch.set_grad_enabled(sc["fn"] == "render"):
            out = rasterizer(**{k: kw[k] for k in sc["call_keys"]})
        assert isinstance(out, tuple) and len(out) == 4
        rendered_image, radii, depth, alpha = out
        images[sc["name"]] = rendered_image.detach().clone()
        n_call = sc["call"]["means3D"]["shape"][0]
        assert rendered_image.shape == (3, H, W) and depth.shape == (1, H, W) and alpha.shape == (1, H, W)
        assert radii.shape == (n_call,) and radii.dtype == torch.int32 and not radii.requires_grad
        vis = radii > 0                                                               # "visibility_filter" (:118)
        assert vis.dtype == torch.bool and 0 < int(vis.sum()) <= n_call
        assert torch.isfinite(rendered_image).all() and float(alpha.max()) <= 1.0 + 1e-5
        if sc["fn"] == "render":
            assert rendered_image.requires_grad
            gt = torch.rand_like(rendered_image)
            loss = (rendered_image - gt).abs().mean()
            rendered_image *= (torch.rand(1, H, W, device=dev) > 0.2).float()           # ssim(): in-place mask (loss_utils.py:45)
            (loss + rendered_image.mean()).backward()
            g = screenspace.grad                                                       # "viewspace_points" (:116, gaussian_model.py:411)
            assert list(g.shape) == sc["returns"]["viewspace_points_grad_shape"] == [N, 3]
            assert float(g[:, 2].abs().max()) == 0.0 and float(g[:, :2].abs().max()) > 0
            if sc["options"].get("vis_mask"):
                assert float(g[~torch.tensor(z["vis_mask"]).to(dev)].abs().max()) == 0.0
            assert leaves["xyz"].grad is not None and torch.isfinite(leaves["xyz"].grad).all()
            assert float(leaves["logit_o"].grad.abs().max()) > 0
            if sc["call"]["shs"] is not None:
                assert float(leaves["shs"].grad.abs().max()) > 0
            if sc["call"]["scales"] is not None:
                assert float(leaves["log_s"].grad.abs().max()) > 0 and float(leaves["rot"].grad.abs().max()) > 0
        else:
            assert not rendered_image.requires_grad                                     # inference.py:462 no_grad
            assert sc["returns"]["order"] == ["rendered_image", "depth", "alpha"]
    # facade branches that must agree pixel for pixel (north_star: 1e-4):
    #  SHs evaluated by the reference's own eval_sh (--convert_SHs_python, :81-85)  ==  SHs evaluated in-kernel
    assert float((images["render_convert_SHs_python"] - images["render_default"]).abs().max()) < 1e-4
    assert float((images["doll_default"] - images["render_default"]).abs().max()) == 0.0
    #  covariance built by the reference's build_scaling_rotation (--compute_cov3D_python, modifier 1.3, :70) == in-kernel
    kw, _, _ = _facade_inputs(trace["scenarios"][0], z, dev)
    sk = h.settings_for(cam, gg.scenes.random_cloud(N, seed=5).to(dev), device=dev, scale_modifier=1.3)._replace(bg=bg)
    with torch.no_grad():
        ref_img, *_ = h.dgr.GaussianRasterizer(raster_settings=sk)(
            means3D=kw["means3D"].detach(), means2D=torch.zeros_like(kw["means3D"]), shs=kw["shs"].detach(),
            colors_precomp=None, opacities=kw["opacities"].detach(), scales=kw["scales"].detach(),
            rotations=kw["rotations"].detach(), cov3D_precomp=None)
    assert float((images["render_compute_cov3D_python"] - ref_img).abs().max()) < 1e-4


# ------------------------------------------------------------------------------------ N3: on-device visibility mask
def _garment_scene(seed=2, n_around=120, n_along=40, per_face=4):
    g = torch.Generator().manual_seed(seed)
    verts, faces = gg.scenes.cylinder_mesh(n_around=n_around, n_along=n_along)
    F = faces.shape[0]
    binding = torch.arange(F).repeat_interleave(per_face)
    bc = torch.rand(binding.shape[0], 3, generator=g) + 0.02
    bc = bc / bc.sum(1, keepdim=True)
    tri = verts[faces][binding]
    pts = bc[:, 0:1] * tri[:, 0] + bc[:, 1:2] * tri[:, 1] + bc[:, 2:3] * tri[:, 2]      # get_barycentric_3d
    return verts, faces, binding, pts


@pytest.mark.parametrize("cam", [(0.0, 0.6, 3.0), (2.1, 1.4, -1.7), (0.05, 0.6, 0.02)])
def test_visibility_mask_is_bit_equal_to_the_cpu_raycast_oracle(cam):
    """Row N3: gg_cast_rays_from_point (projection-grid path AND forced brute force) vs oracle/raycast_oracle.c on
    the synthetic garment cylinder -- primitive ids identical for every ray, hence bit-equal boolean masks for the
    avatar semantics (scene/avatar_gaussian_model.py:227-263).  The third camera sits INSIDE the cylinder: vertices
    fall behind the pinhole plane and the kernel must route itself to the brute-force path."""
    dev = torch.device("cuda:0")
    verts, faces, binding, pts = _garment_scene()
    origin = torch.tensor(cam)
    ref_prim, ref_t = h.c_oracle.cast_rays_from_point(verts, faces, pts, origin)
    for force in (False, True):
        prim, t = gg.cast_rays_from_point(verts.to(dev), faces.to(dev), pts.to(dev), origin.to(dev),
                                          force_bruteforce=force, return_t=True)
        assert torch.equal(prim.cpu(), ref_prim), f"{int((prim.cpu() != ref_prim).sum())} rays differ (force={force})"
        assert torch.equal(t.cpu(), ref_t)
    vis = gg.visible_mask(origin.to(dev), verts.to(dev), faces.to(dev), pts.to(dev), binding.to(dev))
    ref_vis = ref_prim == binding.to(torch.int32)
    assert vis.dtype == torch.bool and torch.equal(vis.cpu(), ref_vis)
    frac = float(ref_vis.float().mean())
    assert 0.2 < frac < 0.99 if abs(cam[0]) + abs(cam[2]) > 1 else frac > 0.9     # outside: the far half is occluded
    h._note("visibility", camera=list(cam), visible_fraction=frac, rays=int(pts.shape[0]), faces=int(faces.shape[0]))


def test_visibility_mask_multi_garment_matches_oracle():
    """inference.py:285-316 semantics: several meshes in one scene (garment cylinder + an inner 'body' cylinder + a
    skirt that hides part of both); a Gaussian is visible iff the first hit lies on its own garment or nothing is hit."""
    dev = torch.device("cuda:0")
    v0, f0, b0, p0 = _garment_scene(seed=3)
    v1, f1 = gg.scenes.cylinder_mesh(n_around=60, n_along=20, radius=0.2, height=1.2, wrinkle_amp=0.0)      # body
    v2, f2 = gg.scenes.cylinder_mesh(n_around=80, n_along=10, radius=0.5, height=0.4, wrinkle_amp=0.0)      # skirt
    g = torch.Generator().manual_seed(9)
    p1 = v1[f1].mean(1)
    p2 = v2[f2].mean(1) + 0.3 * torch.randn(f2.shape[0], 3, generator=g) * 0          # on the surface
    free = torch.tensor([[5.0, 5.0, 5.0], [0.0, 3.0, 0.0]])                            # rays that hit nothing
    pts = torch.cat([p0, p1, p2, free])
    gid = torch.cat([torch.zeros(p0.shape[0]), torch.ones(p1.shape[0]), torch.full((p2.shape[0],), 2.0),
                     torch.zeros(2)]).to(torch.int32)
    cam = torch.tensor([1.8, 0.9, 2.2])
    vis = gg.visible_mask_multi(cam.to(dev), [(v0.to(dev), f0.to(dev)), (v1.to(dev), f1.to(dev)), (v2.to(dev), f2.to(dev))],
                                pts.to(dev), gid.to(dev))
    # oracle: concatenate the meshes the same way
    verts = torch.cat([v0, v1, v2])
    faces = torch.cat([f0, f1 + v0.shape[0], f2 + v0.shape[0] + v1.shape[0]])
    tri_geom = torch.cat([torch.zeros(f0.shape[0]), torch.ones(f1.shape[0]), torch.full((f2.shape[0],), 2.0)]).to(torch.int32)
    prim, _ = h.c_oracle.cast_rays_from_point(verts, faces, pts, cam)
    ref = (tri_geom[prim.clamp_min(0).long()] == gid) | (prim < 0)
    assert torch.equal(vis.cpu(), ref)
    assert bool(ref[-2:].all())                                   # nothing hit -> kept (reference: geometry_ids >= num_gs)
    assert not bool(ref[p0.shape[0]:p0.shape[0] + p1.shape[0]].any())   # the body is entirely behind garment 0


# ------------------------------------------------------------------------------------ N4: densification on the device
def test_densification_cycle_on_device_feeds_the_rasterizer():
    """Row N4 end to end on the GPU: render -> backward -> the means2D side channel drives add_densification_stats ->
    densify_and_prune (clone / split / prune with the keep-one-per-face rule, Adam moments carried along) -> the grown
    model renders again through the fused mesh binding.  The same cycle with the same generator seed on a second copy
    gives identical tensors (what makes replicated ranks stay in lock-step)."""
    import numpy as np, os
    from test_densify_cpu import Model, NAMES
    from gaussian_garments_b200 import densify as D
    dev = torch.device("cuda:0")
    z = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "densify.npz"))
    cam = gg.scenes.ring_cameras(4, width=256, height=192)[1].to(dev)
    bg = torch.tensor([0.1, 0.2, 0.3], device=dev)

    def render(m):
        xyz, sc, ro = gg.FusedMeshBinding(m).world()
        shs = torch.cat([m._features_dc, m._features_rest], dim=1)
        S = h.dgr.GaussianRasterizationSettings(image_height=192, image_width=256, tanfovx=cam.tanfovx, tanfovy=cam.tanfovy,
                                                bg=bg, scale_modifier=1.0, viewmatrix=cam.world_view_transform,
                                                projmatrix=cam.full_proj_transform, sh_degree=3, campos=cam.camera_center,
                                                prefiltered=False, debug=False)
        screenspace = torch.zeros_like(xyz, requires_grad=True) + 0
        screenspace.retain_grad()
        color, radii, _, _ = h.dgr.GaussianRasterizer(raster_settings=S)(
            means3D=xyz, means2D=screenspace, shs=shs, colors_precomp=None, opacities=torch.sigmoid(m._opacity),
            scales=sc, rotations=ro, cov3D_precomp=None)
        return color, radii, screenspace

    outs = []
    for _ in range(2):
        m = Model(z, "s0", device=dev)
        m.mesh_v, m.mesh_f = m.mesh.v, m.mesh.f
        m._scaling.data += 3.0                            # the golden model's splats are tiny at this camera distance
        color, radii, screenspace = render(m)
        (color - 0.5).abs().mean().backward()
        assert float(screenspace.grad[:, :2].abs().max()) > 0
        D.add_densification_stats(m, screenspace, radii > 0)
        m.max_radii2D = torch.max(m.max_radii2D, radii.float())
        n0 = m._xyz.shape[0]
        thr = float((m.xyz_gradient_accum / m.denom.clamp_min(1)).median())
        D.densify_and_prune(m, thr, 0.005, 2.0, 4000, generator=D.rank_consistent_generator(31359, 500, dev))
        assert m._xyz.shape[0] > n0 and int(m.binding_counter.min()) >= 1
        assert int(m.binding_counter.sum()) == m._xyz.shape[0] == m.binding.shape[0]
        for k, a in NAMES.items():                        # Adam moments follow their rows
            p = getattr(m, a)
            assert m.optimizer.state[p]["exp_avg"].shape == p.shape
        color2, radii2, _ = render(m)
        assert radii2.shape[0] == m._xyz.shape[0] and torch.isfinite(color2).all()
        outs.append({k: getattr(m, a).detach().clone() for k, a in NAMES.items()} | {"binding": m.binding.clone()})
    for k in outs[0]:
        assert torch.equal(outs[0][k], outs[1][k]), k


def test_late_colour_forward_path_equals_default_path():
    """Multi-GPU overlap path (dist.GradBucket sets rasterizer.COLOR_GATE): the colour stage runs behind emission and
    sorting and colours are scattered into the packed records afterwards.  Images must be bit-identical to the default
    order, gradients equal up to atomics reordering -- also through the overflow retry."""
    from gaussian_garments_b200 import rasterizer
    dev, st, cam, S = _api_inputs(6000, (320, 256), seed=7)
    grads = _upstream_grads(256, 320)
    rasterizer._hints.clear()
    a = h.run_cuda(S, st, grads)                       # first call: synchronous, default order; leaves a hint
    gate = torch.cuda.Event()
    gate.record(torch.cuda.current_stream())
    rasterizer.COLOR_GATE[0] = gate
    try:
        b = h.run_cuda(S, st, grads)                   # hinted + gated -> late-colour call
        for key in list(rasterizer._hints):
            rasterizer._hints[key] = [64, 16]          # overflow -> retry, still late colour (gate already passed)
        n_retry = rasterizer.STATS["overflow_retries"]
        c = h.run_cuda(S, st, grads)
        assert rasterizer.STATS["overflow_retries"] == n_retry + 1
    finally:
        rasterizer.COLOR_GATE.pop(0, None)
    for other in (b, c):
        for k in ("color", "depth", "alpha", "radii"):
            assert torch.equal(a[k], other[k]), k
        for k in ("means3D", "shs", "opacities", "scales", "rotations"):
            assert h.rel_inf(other["grads"][k], a["grads"][k]) < 1e-4, k


def test_photometric_l1_with_8bit_ground_truth_equals_float_path():
    """The end-to-end loop ships 8-bit frames over PCIe; the fused L1 kernels dequantise them on the fly (value / 255)."""
    dev = torch.device("cuda:0")
    g = torch.Generator().manual_seed(8)
    u8 = torch.randint(0, 256, (3, 136, 200), generator=g, dtype=torch.uint8).to(dev)
    img = torch.rand(3, 136, 200, generator=g).to(dev)
    mask = (torch.rand(1, 136, 200, generator=g) > 0.3).float().to(dev)
    for m in (None, mask):
        a = img.clone().requires_grad_(True)
        b = img.clone().requires_grad_(True)
        ta, l1a, _ = gg.photometric_loss(a, u8, m, 0.0)
        tb, l1b, _ = gg.photometric_loss(b, u8.float() / 255.0, m, 0.0)
        (ta * 1.3).backward()
        (tb * 1.3).backward()
        assert abs(float(l1a) - float(l1b)) < 1e-7 and abs(float(ta) - float(tb)) < 1e-6
        assert torch.equal(a.grad, b.grad)
    # SSIM path with an 8-bit ground truth falls back to a plain conversion
    t1 = gg.photometric_loss(img, u8, None, 0.2)[0]
    t2 = gg.photometric_loss(img, u8.float() / 255.0, None, 0.2)[0]
    assert abs(float(t1) - float(t2)) < 1e-6


# ------------------------------------------------------------------------------------ BASELINE-size cfg4 / cfg5 parity
_V1_CHILD = r"""
import sys, torch
sys.path.insert(0, {root!r}); sys.path.insert(0, {tests!r})
import helpers as h
import gaussian_garments_b200 as gg
st = gg.scenes.random_cloud(6000, seed=23)
cam = gg.scenes.cfg1_camera(352, 240)
g = torch.Generator().manual_seed(5)
grads = (torch.randn(3, 240, 352, generator=g), torch.randn(1, 240, 352, generator=g), torch.randn(1, 240, 352, generator=g))
out = h.run_cuda(h.settings_for(cam, st, device=torch.device("cuda:0")), st, grads)
torch.save(out, {dst!r})
"""


def test_decoupled_warp_blend_kernels_equal_the_barrier_synchronised_ones(tmp_path):
    """The v2 blend kernels recycle their TMA stages through mbarriers (no __syncthreads); compute-sanitizer's racecheck
    does not model that hand-off and flags it.  The round-1 kernels (GG_FWD_KERNEL=v1 / GG_BWD_PATH=v1, CTA-barrier
    recycling, racecheck-clean) run in a child process on the same seeded scene: the forward images must agree to the
    last bits (same per-pixel accumulation order), the gradients up to the float-atomic summation order."""
    import subprocess, sys, os
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    outs = {}
    for tag, env in (("v2", {}), ("v1", {"GG_FWD_KERNEL": "v1", "GG_BWD_PATH": "v1"})):
        dst = str(tmp_path / f"{tag}.pt")
        code = _V1_CHILD.format(root=root, tests=os.path.join(root, "tests"), dst=dst)
        e = dict(os.environ, **env)
        r = subprocess.run([sys.executable, "-c", code], env=e, capture_output=True, text=True, timeout=600)
        assert r.returncode == 0, r.stderr[-2000:]
        outs[tag] = torch.load(dst)
    a, b = outs["v2"], outs["v1"]
    assert torch.equal(a["radii"], b["radii"])
    for k in ("color", "depth", "alpha"):          # same per-pixel accumulation order; FMA contraction may differ
        assert float((a[k] - b[k]).abs().max()) <= 2e-6 * max(1.0, float(b[k].abs().max())), k
    for k, ga in a["grads"].items():
        if ga is not None:
            assert h.rel_inf(ga, b["grads"][k]) <= 2e-5, k


def test_cfg4_full_size_mixed_resolution_with_mesh_vertex_gradient():
    """BASELINE configs[3] at full size: 150k mesh-bound Gaussians (25k faces x 6), SH degree 0 with an [N,1,3] tensor
    (s2_registration.py:158), one 1280x720 and one 1920x1080 camera of the 32-camera ring, gradients chained to mesh.v
    ONLY (scene/mesh_gaussian_model.py:366-371) through the fused mesh binding.  Reference: C oracle for the rasterizer
    (its dL/d{means3D, scales, rotations}) chained through autograd (fp64) over oracle/mesh_chain.py to mesh.v."""
    from oracle import mesh_chain as mc
    dev = torch.device("cuda:0")
    model = gg.scenes.MeshBoundGaussians(250, 50, 6, sh_degree=0, max_sh_degree=0)          # 25 000 faces -> 150 000
    assert model.binding.shape[0] == 150_000
    ring720, ring1080 = gg.scenes.ring_cameras(32, width=1280, height=720), gg.scenes.ring_cameras(32, width=1920, height=1080)
    total_ref = torch.zeros_like(model.mesh_v, dtype=torch.float64)
    total_got = torch.zeros_like(model.mesh_v)
    for cam in (ring720[6], ring1080[7]):
        H, W = cam.image_height, cam.image_width
        # product: fused binding -> rasterizer -> backward to mesh.v
        m = gg.scenes.MeshBoundGaussians(250, 50, 6, sh_degree=0, max_sh_degree=0).to(dev)
        m.mesh_v = m.mesh_v.detach().clone().requires_grad_(True)
        xyz, sc, ro = gg.FusedMeshBinding(m).world()
        # The rasterizer oracle is handed the world-space tensors the fused binding produced (its own parity against the
        # reference chain is test_fused_mesh_binding_*): the six splats of a face are coplanar, so a last-bit difference
        # in a centre would legitimately swap two depth keys and move pixels by ~1e-2 on either side.
        st = gg.scenes.GaussianState(xyz.detach().cpu(), sc.detach().cpu(), ro.detach().cpu(), m.get_opacity.detach().cpu().contiguous(),
                                     m.get_features.detach().cpu().contiguous(), 0, model.bg)
        S = h.settings_for(cam, st, device=dev)
        ref = h.run_c_oracle(S, st, None)
        grads = h.mask_upstream(_upstream_grads(H, W, seed=H, depth_alpha=False), ref["fragile"])
        rg = ref["ctx"].backward(*grads)
        # reference chain to mesh.v (fp64 autograd through the golden-pinned restatement of the binding)
        v64 = model.mesh_v.double().requires_grad_(True)
        chain = mc.MeshChain(v64, model.mesh_f, model.binding, model._xyz.double(), model._scaling.double(), model._rotation.double())
        chain.update_face_coor()
        # q and -q are the same rotation: where the fp32 kernel and the fp64 chain pick different branches of the
        # matrix -> quaternion conversion (near-ties of its four candidates) the two world quaternions differ in sign.
        # rg was evaluated at the product's sign, so the chain's quaternion is aligned to it before the inner product.
        ro_chain = chain.get_rotation
        sgn = torch.sign((ro_chain.detach() * ro.detach().cpu().double()).sum(-1, keepdim=True))
        flipped = int((sgn < 0).sum())
        assert flipped <= 0.01 * sgn.numel(), flipped
        assert float((ro_chain.detach() * sgn - ro.detach().cpu().double()).abs().max()) < 5e-6
        ((chain.get_xyz * rg["means3D"].double()).sum() + (chain.get_scaling * rg["scales"].double()).sum() +
         (ro_chain * sgn * rg["rotations"].double()).sum()).backward()
        total_ref += v64.grad
        color, radii, _, _ = h.dgr.GaussianRasterizer(raster_settings=S)(
            means3D=xyz, means2D=torch.zeros_like(xyz), shs=m.get_features, colors_precomp=None, opacities=m.get_opacity,
            scales=sc, rotations=ro, cov3D_precomp=None)
        assert int((radii.cpu() != ref["radii"]).sum()) <= 3
        h.assert_images_close(dict(color=color.detach().cpu(), depth=ref["depth"], alpha=ref["alpha"]), ref)
        (color * grads[0].to(dev)).sum().backward()
        total_got += m.mesh_v.grad.cpu()
        ref["ctx"].close()
    err = h.rel_inf(total_got, total_ref.float())
    h._note("cfg4_full", mesh_v_rel_inf=err, verts=int(model.mesh_v.shape[0]), quaternion_sign_flips_last_view=flipped)
    assert err <= h.GRAD_TOL, f"mesh.v gradient over the 720p + 1080p views: rel-inf {err:.3e}"


def test_cfg5_half_million_gaussians_at_2160p_parity():
    """BASELINE configs[4] shape at a size the C oracle finishes in about a minute: 500k stress-cloud Gaussians on one
    3840x2160 camera (thousands of instances per tile -> the lazy bucketed forward), forward + gradients at 1e-4 / 1e-3."""
    st = gg.scenes.stress_cloud(500_000)
    cam = gg.scenes.ring_cameras(160, width=3840, height=2160)[11]
    S = h.settings_for(cam, st, device=torch.device("cuda:0"))
    ref = h.run_c_oracle(S, st, None)
    grads = h.mask_upstream(_upstream_grads(2160, 3840, seed=9), ref["fragile"])
    got = h.run_cuda(S, st, grads)
    ref["grads"] = ref["ctx"].backward(*grads)
    off = ref["ctx"].binning()["tile_off"]
    assert int((off[1:] - off[:-1]).max()) > 4096
    assert int((got["radii"] != ref["radii"]).sum()) <= 5
    h.assert_images_close(got, ref, max_fragile_frac=0.05)
    h.assert_grads_close(got["grads"], ref["grads"])
    h._note("cfg5_500k", K_oracle=int(off[-1]), max_tile=int((off[1:] - off[:-1]).max()))
    ref["ctx"].close()
