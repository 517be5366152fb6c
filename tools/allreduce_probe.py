"""dev probe (torchrun): where does the multi-GPU step overhead go?  NCCL all-reduce of the 71 MB gradient bucket alone,
AVG vs SUM, and per-rank compute-time skew."""
import os, sys, time, json
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT)
import torch, torch.distributed as dist
rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local); dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
n = 17_700_000 + 64 * 5
flat = torch.randn(n, device=dev)
def timeit(fn, iters=30, warm=5):
    for _ in range(warm): fn()
    torch.cuda.synchronize(); dist.barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters): fn()
    e1.record(); torch.cuda.synchronize()
    t = torch.tensor([e0.elapsed_time(e1) / iters], device=dev); dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t)
res = {"world": world, "bytes": n * 4}
res["allreduce_avg_ms"] = timeit(lambda: dist.all_reduce(flat, op=dist.ReduceOp.AVG))
res["allreduce_sum_ms"] = timeit(lambda: dist.all_reduce(flat, op=dist.ReduceOp.SUM))
small = torch.randn(1024, device=dev)
res["allreduce_4KB_ms"] = timeit(lambda: dist.all_reduce(small))
# reduce_scatter + all_gather (what a sharded optimizer would use)
pad = (n + world - 1) // world * world
big = torch.randn(pad, device=dev); shard = torch.empty(pad // world, device=dev)
res["reduce_scatter_ms"] = timeit(lambda: dist.reduce_scatter_tensor(shard, big, op=dist.ReduceOp.AVG))
res["all_gather_ms"] = timeit(lambda: dist.all_gather_into_tensor(big, shard))
if rank == 0: print(json.dumps(res), flush=True)
dist.destroy_process_group()
