#!/usr/bin/env python
"""Measure the BASELINE.json configs other than the bench line: prints one JSON line per run.

    python tools/run_configs.py --config cfg1
    python tools/run_configs.py --config cfg4 --views 8                      # single GPU smoke of cfg4
    python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29544 \
        tools/run_configs.py --config cfg4
    ... --nproc-per-node 8 ... --config cfg5 [--views 160]

cfg1  10k random Gaussians, 512x512 (parity gate config; timed here, parity in tests/)
cfg4  s2_registration-style: 150k mesh-bound Gaussians, SH degree 0 (M=1), cameras alternating 1280x720 / 1920x1080,
      gradients chained to mesh.v ONLY and only mesh.v all-reduced (scene/mesh_gaussian_model.py:366-371)
cfg5  stress: 2M random Gaussians, 3840x2160, per-kernel time split (tile sort vs blend)
"""
import argparse
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--config", required=True, choices=["cfg1", "cfg4", "cfg5"])
    ap.add_argument("--views", type=int, default=0, help="total views per step (default: the config's own)")
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=1)
    ap.add_argument("--gaussians", type=int, default=0)
    ap.add_argument("--fused-binding", action="store_true", help="cfg4: fused mesh-binding kernels instead of the torch chain")
    args = ap.parse_args()
    rank, world, local = (int(os.environ.get(k, d)) for k, d in (("RANK", 0), ("WORLD_SIZE", 1), ("LOCAL_RANK", 0)))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    import diff_gaussian_rasterization_depth_alpha as dgr
    import gaussian_garments_b200 as gg
    from gaussian_garments_b200 import _capi, rasterizer
    from gaussian_garments_b200.dist import GradBucket, shard_views

    if args.config == "cfg1":
        st = gg.scenes.random_cloud(args.gaussians or 10_000).to(dev)
        cams = [gg.scenes.cfg1_camera().to(dev)]
        n_views = args.views or 1
        params = [t.clone().requires_grad_(True) for t in (st.means3D, st.scales, st.rotations, st.opacities, st.shs)]
        state_fn = lambda: (params[0], params[1], params[2], params[3], params[4])
        sh_degree, bg = 3, st.bg
        leafs = params
    elif args.config == "cfg4":
        n = args.gaussians or 150_000
        model = gg.scenes.MeshBoundGaussians(250, max(1, (n // 6) // 500), 6, sh_degree=0, max_sh_degree=0).to(dev)
        model.mesh_v.requires_grad_(True)
        cams = []
        for i, c in enumerate(gg.scenes.ring_cameras(32, width=1280, height=720)):
            cams.append((c if i % 2 == 0 else gg.scenes.ring_cameras(32, width=1920, height=1080)[i]).to(dev))
        n_views = args.views or 32

        fused = gg.FusedMeshBinding(model) if args.fused_binding else None

        def state_fn():
            if fused is not None:
                xyz, sc, ro = fused.world()
                return (xyz, sc, ro, model.get_opacity, model.get_features)
            model.update_face_coor()
            return (model.get_xyz, model.get_scaling, model.get_rotation, model.get_opacity, model.get_features)
        sh_degree, bg = 0, model.bg
        leafs = [model.mesh_v]
    else:
        st = gg.scenes.stress_cloud(args.gaussians or 2_000_000).to(dev)
        n_cam = args.views or 160
        cams = [c.to(dev) for c in gg.scenes.ring_cameras(n_cam, width=3840, height=2160)]
        n_views = n_cam
        params = [t.clone().requires_grad_(True) for t in (st.means3D, st.scales, st.rotations, st.opacities, st.shs)]
        state_fn = lambda: (params[0], params[1], params[2], params[3], params[4])
        sh_degree, bg = 3, st.bg
        leafs = params

    bucket = GradBucket(leafs, world)
    lib = _capi.load()
    mine = shard_views(n_views, rank, world)
    gts = {}

    def render_view(vi):
        cam = cams[vi % len(cams)]
        H, W = cam.image_height, cam.image_width
        if (H, W) not in gts:
            gts[(H, W)] = torch.rand(3, H, W, device=dev)
        S = dgr.GaussianRasterizationSettings(image_height=H, image_width=W, tanfovx=cam.tanfovx, tanfovy=cam.tanfovy,
                                              bg=bg, scale_modifier=1.0, viewmatrix=cam.world_view_transform,
                                              projmatrix=cam.full_proj_transform, sh_degree=sh_degree,
                                              campos=cam.camera_center, prefiltered=False, debug=False)
        m3, sc, ro, op, sh = state_fn()
        color, radii, depth, alpha = dgr.GaussianRasterizer(raster_settings=S)(
            means3D=m3, means2D=torch.zeros_like(m3, requires_grad=True), shs=sh, colors_precomp=None, opacities=op,
            scales=sc, rotations=ro, cov3D_precomp=None)
        return (color - gts[(H, W)]).abs().mean() / len(mine)

    def step():
        bucket.zero()                        # first view's gradients land in the bucket zero-copy, the rest accumulate
        for vi in mine:                      # this rank's views; gradients accumulate locally
            render_view(vi).backward()
        bucket.all_reduce()                  # ONE collective per step (leaf-parameter level)

    for _ in range(args.warmup):
        step()
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(args.steps):
        step()
    e1.record()
    torch.cuda.synchronize()
    ms = torch.tensor([e0.elapsed_time(e1)], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    ms = float(ms.item())

    out = None
    if rank == 0:
        _capi.kernel_timing(True)
        split, Ks = {}, []
        sample = mine[: min(4, len(mine))]
        phase = torch.zeros(2, dtype=torch.int64, device=dev)        # lazy fused forward: [ordering, pack + blend] cycles
        lib.gg_debug_lazy_phase_counters(phase.data_ptr())
        for vi in sample:
            for p in leafs:
                p.grad = None
            render_view(vi).backward()
            torch.cuda.synchronize()
            Ks.append(rasterizer.LAST_NUM_RENDERED)
            for k, v in _capi.kernel_times().items():
                split[k] = split.get(k, 0.0) + v / len(sample)
        _capi.kernel_timing(False)
        torch.cuda.synchronize()
        lib.gg_debug_lazy_phase_counters(None)
        ph = [int(v) for v in phase.tolist()]
        lazy_share = ph[0] / max(1, ph[0] + ph[1]) if (ph[0] + ph[1]) > 0 else None
        fused = lazy_share is not None and split.get("sort_pack", 0.0) == 0.0
        # dense scenes run the fused lazy forward: ONE kernel orders (depth buckets + per-bucket sorts) and blends a tile.
        # Its in-kernel cycle counters give the honest split of that kernel's time.
        sort_in_fwd = split.get("blend_fwd", 0.0) * lazy_share if fused else 0.0
        tot = sum(split.values())
        out = {"config": args.config, "n_gpus": world, "views_per_step": n_views, "steps": args.steps,
               "ms_per_step": ms / args.steps, "views_per_s": n_views * args.steps / (ms * 1e-3),
               "num_rendered_sample": Ks, "kernel_ms_per_view": {k: round(v, 4) for k, v in sorted(split.items(), key=lambda kv: -kv[1])},
               "sort_vs_blend": {"tile_sort_ms": round(split.get("sort_pack", 0) + split.get("emit", 0) + sort_in_fwd, 4),
                                 "blend_fwd_ms": round(split.get("blend_fwd", 0) - sort_in_fwd, 4),
                                 "blend_bwd_ms": round(split.get("blend_bwd", 0), 4),
                                 "sort_share": round((split.get("sort_pack", 0) + split.get("emit", 0) + sort_in_fwd) / max(tot, 1e-9), 3),
                                 "forward_path": "lazy fused (ordering inside the forward kernel; split by in-kernel cycle counters)" if fused else "sort_pack + blend_fwd",
                                 "lazy_ordering_cycle_share": None if lazy_share is None else round(lazy_share, 4),
                                 "emit_ms": round(split.get("emit", 0), 4)},
               "grad_allreduce": "mesh.v only" if args.config == "cfg4" else "5 Gaussian tensors (flat bucket)",
               "mesh_binding": ("fused kernels" if args.fused_binding else "torch chain") if args.config == "cfg4" else None,
               "mem_gb": round(torch.cuda.max_memory_allocated() / 1e9, 2)}
        print(json.dumps(out), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
