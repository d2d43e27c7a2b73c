"""Kernel-level breakdown of one configs[2] training step (tools/full_step.py) with torch.profiler: where the loss path's
time goes next to the backbone.  python tools/prof_full_step.py [shared2x2|full] [B]"""
import os
import sys
import types

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench as BN  # noqa: E402
import rangeclip_b200 as R  # noqa: E402
from tools.full_step import Backbone  # noqa: E402
from tools.stage_reference import import_reference_model  # noqa: E402

variant = sys.argv[1] if len(sys.argv) > 1 else "shared2x2"
B = int(sys.argv[2]) if len(sys.argv) > 2 else 64
dev = torch.device("cuda")
c = BN.CFG
H, W, D, K, C = c["H"], c["W"], c["D"], c["K"], c["C"]
DepthUNet = import_reference_model()
torch.manual_seed(0)
core = DepthUNet('resnet', dev, embedding_dim=D, use_batch_norm=True, activation_func='relu').to(dev)
net = Backbone(core, variant == "shared2x2")
opt = torch.optim.Adam(core.parameters(), lr=1e-4)
g = torch.Generator(device=dev).manual_seed(1234)
depth = torch.rand(B, 1, H, W, device=dev, generator=g) + 0.5
seg = BN.make_labels(B, H, W, c["G"], torch.Generator().manual_seed(1234)).to(dev)
text = torch.nn.functional.normalize(torch.randn(C, D, device=dev, generator=g), dim=1)
rest = torch.arange(c["G"] + 1, C, device=dev)
contrast = torch.unique(torch.cat([torch.arange(1, c["G"] + 1, device=dev), rest[torch.randperm(rest.numel(), device=dev, generator=g)[: K - c["G"]]]]))
sets = BN.similarity_sets(contrast.tolist(), c["G"], C)
img = torch.nn.functional.normalize(torch.randn(B, D, device=dev, generator=g), dim=1)
labs = seg[:, H // 2, W // 2].tolist()
kw = dict(W_text=1.0, W_image=0.5, W_smooth=2e2, percent_image_sampling=0.7, k_distractors=K - c["G"], pct_medium=0.0, pct_hard=1.0, pct_rand=0.0,
          contrast_builder=os.environ.get("BUILDER", "device"))


def step():
    with torch.autocast("cuda", dtype=torch.bfloat16):
        e = net(depth)
    with torch.profiler.record_function("LOSS_PATH"):
        with torch.no_grad():
            area = R.pool_objects_per_image(e, seg, list(range(B)), labs, shared2x2=(variant == "shared2x2"))
        fn = R.compute_loss_shared2x2 if variant == "shared2x2" else R.compute_loss
        loss, info = fn(core, e, seg, text, sets, area, img, **kw)
    loss.backward()
    opt.step(); opt.zero_grad(set_to_none=True)


for _ in range(3):
    step()
torch.cuda.synchronize()
with torch.profiler.profile(activities=[torch.profiler.ProfilerActivity.CPU, torch.profiler.ProfilerActivity.CUDA]) as prof:
    for _ in range(2):
        step()
    torch.cuda.synchronize()
tab = prof.key_averages().table(sort_by="cuda_time_total", row_limit=45, max_name_column_width=70)
print(tab)
os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
open(os.path.join(ROOT, "gpurun_out", f"r2_full_step_profile_{variant}.txt"), "w").write(tab)
