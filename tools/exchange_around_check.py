#!/usr/bin/env python
"""Multi-GPU check of the graph + eager-exchange ('around') step order with the device-side colour gate.

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29577 \
        tools/exchange_around_check.py [--steps 24] [--gaussians 300000]

Every step: replay the captured step graph (forward + loss + backward, no exchange inside), snapshot the LOCAL bucket,
issue both exchanges behind it (geometry on the compute stream, SH on the side stream + gate signal), and let the next
replay start immediately.  The SH block of step i is checked, after the fact, against the all-gathered mean of the
snapshots: a gate that let backward(i+1) overwrite the SH block while exchange(i) was still reading would show up here.
Prints one JSON line on rank 0.
"""
import argparse
import copy
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--steps", type=int, default=24)
    ap.add_argument("--gaussians", type=int, default=300_000)
    args = ap.parse_args()
    rank, world, local = (int(os.environ.get(k, d)) for k, d in (("RANK", 0), ("WORLD_SIZE", 1), ("LOCAL_RANK", 0)))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    import diff_gaussian_rasterization_depth_alpha as dgr
    import gaussian_garments_b200 as gg
    from gaussian_garments_b200.dist import GradBucket

    st = gg.scenes.mesh_bound_state(args.gaussians).to(dev)
    cams = [c.to(dev) for c in gg.scenes.ring_cameras(8, width=1920, height=1080)]
    params = [t.detach().clone().requires_grad_(True) for t in (st.means3D, st.scales, st.rotations, st.opacities, st.shs)]
    bucket = GradBucket(params, world, deferred=(4,))
    if bucket.impl != "nvls_multimem":
        if rank == 0:
            print(json.dumps({"skipped": f"no multicast exchange here: {bucket.nvls_error}"}))
        dist.destroy_process_group()
        return
    H, W = 1080, 1920
    gt = torch.rand(3, H, W, device=dev)
    slot_cam = [tuple(torch.empty_like(t) for t in (cams[0].world_view_transform, cams[0].full_proj_transform,
                                                    cams[0].camera_center)) for _ in range(2)]
    camobj = []
    for k in range(2):
        c = copy.copy(cams[0])
        c.world_view_transform, c.full_proj_transform, c.camera_center = slot_cam[k]
        camobj.append(c)

    def load(k, ci):
        for d, s in zip(slot_cam[k], (cams[ci].world_view_transform, cams[ci].full_proj_transform, cams[ci].camera_center)):
            d.copy_(s, non_blocking=True)

    def body(k):
        cam = camobj[k]
        S = dgr.GaussianRasterizationSettings(image_height=H, image_width=W, tanfovx=cam.tanfovx, tanfovy=cam.tanfovy,
                                              bg=st.bg, scale_modifier=1.0, viewmatrix=cam.world_view_transform,
                                              projmatrix=cam.full_proj_transform, sh_degree=st.sh_degree,
                                              campos=cam.camera_center, prefiltered=False, debug=False)
        color, radii, depth, alpha = dgr.GaussianRasterizer(raster_settings=S)(
            means3D=params[0], means2D=torch.zeros_like(params[0], requires_grad=True), shs=params[4],
            colors_precomp=None, opacities=params[3], scales=params[1], rotations=params[2], cov3D_precomp=None)
        loss = gg.photometric_loss(color, gt, None, 0.0)[0]
        bucket.zero()
        loss.backward()
        bucket.adopt()

    side = torch.cuda.Stream(device=dev)
    side.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(side):                      # eager warm-up: sizes the capacity hints (late-colour forward)
        for i in range(4):
            load(i % 2, (i * world + rank) % len(cams))
            body(i % 2)
            bucket.all_reduce()
    torch.cuda.current_stream().wait_stream(side)
    bucket.wait()
    torch.cuda.synchronize()
    dist.barrier()
    bucket.use_device_gate(True)
    graphs = []
    for k in range(2):
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g):
            body(k)
        graphs.append(g)
    torch.cuda.synchronize()

    snaps = []
    results = []          # exchanged buckets, cloned right after the exchange of step i is complete
    for i in range(args.steps):
        k = i % 2
        load(k, (i * world + rank) % len(cams))
        graphs[k].replay()
        snaps.append(bucket.flat.clone())                       # local gradients of step i (compute stream)
        bucket.exchange_immediate()
        bucket.exchange_deferred_async()
        # the exchanged SH block, read on the SIDE stream right behind its exchange (before backward(i+1) may touch it)
        with torch.cuda.stream(bucket.comm_stream):
            sh_after = bucket.flat[bucket.split:].clone()
        geo_after = bucket.flat[:bucket.split].clone()
        results.append((geo_after, sh_after))
    bucket.wait()
    torch.cuda.synchronize()
    timed_out = bucket.device_gate.timed_out()
    words = [int(v) for v in bucket.device_gate.words.tolist()]
    bucket.use_device_gate(False)

    worst = 0.0
    for i in range(args.steps):
        parts = [torch.empty_like(snaps[i]) for _ in range(world)]
        dist.all_gather(parts, snaps[i])
        mean = torch.stack(parts).double().mean(0)
        got = torch.cat([results[i][0], results[i][1]]).double()
        scale = mean.abs().max().clamp_min(1e-30)
        worst = max(worst, float((got - mean).abs().max() / scale))
    w = torch.tensor([worst], dtype=torch.float64, device=dev)
    dist.all_reduce(w, op=dist.ReduceOp.MAX)
    if rank == 0:
        print(json.dumps({"check": "graph step + eager exchanges + device colour gate", "n_gpus": world, "steps": args.steps,
                          "max_rel_err_vs_gathered_mean": float(w.item()), "ok": bool(float(w.item()) <= 1e-6 and not timed_out),
                          "gate_words_G_X_timeout": words[:3], "gate_timed_out": timed_out}), flush=True)
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
