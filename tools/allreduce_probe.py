"""dev probe (torchrun): the gradient exchange in isolation.  NCCL all-reduce vs the in-switch multimem kernel
(csrc/allreduce.cu) on the cfg2 bucket (70.8 MB) and on its two blocks (13.2 MB geometry / 57.6 MB SH), correctness of the
multimem result against an all-gather, and the NCCL algorithm NCCL_DEBUG reports.  Output: one JSON line (rank 0)."""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import torch.distributed as dist

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
import diff_gaussian_rasterization_depth_alpha  # noqa: F401,E402
from gaussian_garments_b200.dist import GradBucket  # noqa: E402

N = 300_000
params = [torch.zeros(N, 3, device=dev, requires_grad=True), torch.zeros(N, 3, device=dev, requires_grad=True),
          torch.zeros(N, 4, device=dev, requires_grad=True), torch.zeros(N, 1, device=dev, requires_grad=True),
          torch.zeros(N, 16, 3, device=dev, requires_grad=True)]


def timeit(fn, iters=20, warm=5):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    dist.barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        fn()
    e1.record()
    torch.cuda.synchronize()
    t = torch.tensor([e0.elapsed_time(e1) / iters], device=dev)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t)


res = {"world": world}
for symmetric in (False, True):
    b = GradBucket(params, world, register=False, deferred=(4,), symmetric=symmetric)
    tag = b.impl
    nbytes = b.numel * 4
    g = torch.Generator(device=dev).manual_seed(100 + rank)
    b.flat.copy_(torch.randn(b.numel, device=dev, generator=g))
    local_copy = b.flat.clone()
    for i, p in enumerate(params):
        p.grad = b.view(i)                   # as after a backward through the grad sinks
    b.all_reduce(overlap=False)
    chk = b.check_against_gather(local_copy)
    full = timeit(lambda: b.all_reduce(overlap=False))
    geom = timeit(lambda: b._reduce_range(0, b.split, 0))
    sh = timeit(lambda: b._reduce_range(b.split, b.numel, 1))
    res[tag if symmetric else "nccl"] = {
        "impl": tag, "nvls_error": b.nvls_error, "bytes": nbytes, "check": chk,
        "full_ms": full, "geometry_ms": geom, "sh_ms": sh,
        "full_busbw_gbs": nbytes / (full * 1e-3) / 1e9 * 2 * (world - 1) / world,
        "full_algbw_gbs": nbytes / (full * 1e-3) / 1e9}
    del b
small = torch.randn(1024, device=dev)
res["nccl_allreduce_4KB_ms"] = timeit(lambda: dist.all_reduce(small))
if rank == 0:
    print(json.dumps(res), flush=True)
dist.destroy_process_group()
