"""Developer diagnostic (GPU box): parity of the CUDA path vs the oracles at cfg1 + a timing peek at cfg2."""
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import torch
import helpers as h

gg = h.gg


def parity(N=10000, res=512):
    st = gg.scenes.random_cloud(N)
    cam = gg.scenes.cfg1_camera(res, res)
    S = h.settings_for(cam, st, device=torch.device("cuda:0"))
    g = torch.Generator().manual_seed(1)
    H, W = cam.image_height, cam.image_width
    grads = (torch.randn(3, H, W, generator=g), torch.randn(1, H, W, generator=g) * 0.3, torch.randn(1, H, W, generator=g))
    got = h.run_cuda(S, st, grads)
    ref = h.run_c_oracle(S, st, grads, fragile_eps=1e-3)
    print("K oracle", ref["K"], "radii mismatches", int((got["radii"] != ref["radii"]).sum()))
    for name in ("color", "depth", "alpha"):
        d = (got[name] - ref[name]).abs().amax(0)
        print(f"{name}: max {float(d.max()):.3e}  n>1e-4 {int((d > 1e-4).sum())}  of which fragile {int(((d > 1e-4) & ref['fragile']).sum())}")
    print("fragile px", int(ref["fragile"].sum()))
    for k, v in ref["grads"].items():
        if v is None:
            continue
        gv = got["grads"][k]
        print(f"grad {k}: rel-inf {h.rel_inf(gv, v):.3e}  finite {bool(torch.isfinite(gv).all())} max {float(v.abs().max()):.3e}")


def timing(N=300_000):
    dev = torch.device("cuda:0")
    st = gg.scenes.mesh_bound_state(N).to(dev)
    cams = [c.to(dev) for c in gg.scenes.cfg2_cameras(8)]
    gt = torch.rand(3, 1080, 1920, device=dev)
    leaves = [t.clone().requires_grad_(True) for t in (st.means3D, st.shs, st.opacities, st.scales, st.rotations)]
    import diff_gaussian_rasterization_depth_alpha as dgr
    def step(cam):
        S = h.settings_for(cam, st, device=dev)
        m2 = torch.zeros_like(leaves[0], requires_grad=True)
        color, radii, depth, alpha = dgr.GaussianRasterizer(S)(means3D=leaves[0], means2D=m2, shs=leaves[1], colors_precomp=None,
                                                              opacities=leaves[2], scales=leaves[3], rotations=leaves[4], cov3D_precomp=None)
        loss = (color - gt).abs().mean()
        loss.backward()
        return loss, radii
    for i in range(5):
        loss, radii = step(cams[i % 8])
    torch.cuda.synchronize()
    t0 = time.time()
    n = 40
    for i in range(n):
        loss, radii = step(cams[i % 8])
    torch.cuda.synchronize()
    dt = (time.time() - t0) / n
    print(f"cfg2 fwd+bwd {dt*1e3:.3f} ms/view -> {1/dt:.1f} views/s; visible {int((radii>0).sum())} loss {float(loss):.4f}")
    from torch.profiler import profile, ProfilerActivity
    with profile(activities=[ProfilerActivity.CUDA]) as prof:
        for i in range(8):
            step(cams[i % 8])
        torch.cuda.synchronize()
    print(prof.key_averages().table(sort_by="cuda_time_total", row_limit=18, max_name_column_width=60))


if __name__ == "__main__":
    print(torch.cuda.get_device_name(0))
    parity(2000, 256)
    parity(10000, 512)
    timing()
