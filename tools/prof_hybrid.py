"""Kernel-level breakdown of the hybrid loss (text + area-image + smoothness) through compute_loss, X on the device, bf16 or fp32 X
(bench.py key `hybrid`).   python tools/prof_hybrid.py [fp32|bf16] [builder]"""
import os, sys
import numpy as np
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench as BN
import rangeclip_b200 as R
from rangeclip_b200 import pool_objects_per_image

kind = sys.argv[1] if len(sys.argv) > 1 else "fp32"
builder = sys.argv[2] if len(sys.argv) > 2 else "reference"
dev = torch.device("cuda")
c = BN.CFG
B, H, W, D, K, C = 64, c["H"], c["W"], c["D"], c["K"], c["C"]
g = torch.Generator(device=dev).manual_seed(1234)
x = torch.empty(B, D, H, W, device=dev, dtype=torch.bfloat16)
for b in range(B):
    xb = torch.randn(D, H * W, device=dev, generator=g)
    x[b] = (xb / xb.norm(dim=0, keepdim=True)).view(D, H, W).to(torch.bfloat16)
if kind == "fp32":
    x = x.float()
seg = BN.make_labels(B, H, W, c["G"], torch.Generator().manual_seed(1234)).to(dev)
text = torch.nn.functional.normalize(torch.randn(C, D, device=dev, generator=g), dim=1)
rest = torch.arange(c["G"] + 1, C, device=dev)
contrast = torch.unique(torch.cat([torch.arange(1, c["G"] + 1, device=dev), rest[torch.randperm(rest.numel(), device=dev, generator=g)[: K - c["G"]]]]))
sets = BN.similarity_sets(contrast.tolist(), c["G"], C)
img = torch.nn.functional.normalize(torch.randn(B, D, device=dev, generator=g), dim=1)
labels = seg[:, 128, 128].tolist()
model = BN.TemperatureHolder(c["tau"]).to(dev)
xg = x.detach().requires_grad_(True)
params = [xg] + list(model.parameters())


def f():
    with torch.no_grad():
        area = pool_objects_per_image(xg, seg, list(range(B)), labels)
    total, info = R.compute_loss(model, xg, seg, text, sets, area, img, W_text=1.0, W_image=0.5, W_smooth=2e2, percent_image_sampling=c["pct_sampling"],
                                 k_distractors=K - c["G"], pct_medium=0.0, pct_hard=1.0, pct_rand=0.0, contrast_builder=builder)
    torch.autograd.grad(total, params, allow_unused=True)


for _ in range(3):
    f()
torch.cuda.synchronize()
with torch.profiler.profile(activities=[torch.profiler.ProfilerActivity.CUDA]) as prof:
    for _ in range(3):
        f()
    torch.cuda.synchronize()
print(prof.key_averages().table(sort_by="cuda_time_total", row_limit=22, max_name_column_width=90))
