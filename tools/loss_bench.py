"""dev/measurement helper (GPU box): fused photometric loss vs the torch formulation of the reference (fwd+bwd, 1080p)."""
import os, sys, json
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT)
import torch
import diff_gaussian_rasterization_depth_alpha  # noqa
import gaussian_garments_b200 as gg
from oracle import loss_oracle as lo
dev = torch.device("cuda:0")
H, W = 1080, 1920
g = torch.Generator().manual_seed(0)
img = torch.rand(3, H, W, generator=g).to(dev); gt = torch.rand(3, H, W, generator=g).to(dev)
mask = (torch.rand(1, H, W, generator=g) > 0.3).float().to(dev)
def timeit(fn, n=30):
    for _ in range(5): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n
def fused(lam, m):
    x = img.detach().requires_grad_(True)
    gg.photometric_loss(x, gt, m, lam)[0].backward()
def torch_ref(lam, m):
    x = img.detach().requires_grad_(True)
    (lo.total_loss(x, gt, m, lam) if lam else lo.l1_loss(x, gt, m)).backward()
out = {}
for lam in (0.2, 0.0):
    for name, m in (("mask", mask), ("nomask", None)):
        out[f"fused_lam{lam}_{name}_ms"] = round(timeit(lambda: fused(lam, m)), 4)
        out[f"torch_lam{lam}_{name}_ms"] = round(timeit(lambda: torch_ref(lam, m)), 4)
print(json.dumps(out))
