"""Where the public API's time goes beyond the fused kernel (bench.py api_device): the phases of
rangeclip_b200.text_contrastive_loss + backward at the headline size, each bracketed by a device synchronize.
python tools/prof_api.py  ->  gpurun_out/r2_api_breakdown.json"""
import json
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench as BN  # noqa: E402
import rangeclip_b200 as R  # noqa: E402
from rangeclip_b200 import losses, ops  # noqa: E402

dev = torch.device("cuda")
c = BN.CFG
B, D, H, W, K, C = c["B"], c["D"], c["H"], c["W"], c["K"], c["C"]
wl = BN.make_device_workload(dev, 1234, B)
x, text, seg = wl["x"], wl["text"], wl["seg"]
sets = BN.similarity_sets(wl["contrast"].tolist(), c["G"], C)
model = BN.TemperatureHolder(c["tau"]).to(dev)
T = {}


def tick(name, t0):
    torch.cuda.synchronize()
    T.setdefault(name, []).append((time.perf_counter() - t0) * 1e3)
    return time.perf_counter()


for it in range(6):
    xg = x.detach().requires_grad_(True)
    np.random.seed(0); torch.manual_seed(0)
    torch.cuda.synchronize()
    t = time.perf_counter()
    hw = H * W
    n_samples = int(c["pct_sampling"] * hw)
    rand_indices = torch.randint(0, hw, (B, n_samples), device=dev)
    t = tick("randint", t)
    target_flat = seg.reshape(B, -1)
    label_samples = torch.gather(target_flat, 1, rand_indices)
    label_samples = label_samples[label_samples > 0]
    t = tick("gather+filter labels", t)
    unique_labels = torch.unique(label_samples)
    t = tick("unique", t)
    contrast = losses.build_contrast_indices(unique_labels, C, sets, K - c["G"], 0.0, 1.0, 0.0, dev)
    t = tick("build_contrast_indices (host sets, isin, randperm, unique)", t)
    label_map = torch.full((C,), -1, dtype=torch.int32, device=dev)
    label_map[contrast] = torch.arange(contrast.shape[0], device=dev, dtype=torch.int32)
    w, y = torch.ops.rangeclip.sample_weights(target_flat, rand_indices, label_map)
    t = tick("label_map + sample_weights", t)
    t_norm, tb, ttb = torch.ops.rangeclip.text_prepare(text, contrast)
    t = tick("text_prepare", t)
    loss = ops.infonce(xg, t_norm, model.log_temperature_text, y, w, "auto", t_bf16=(tb, ttb))
    t = tick("infonce op forward (weight_sum + fused kernel + tau sync)", t)
    total = 1.0 * loss
    host = torch.stack([total.detach().float(), loss.detach().float()]).tolist()
    t = tick("loss_info readback", t)
    total.backward()
    t = tick("backward (late scale)", t)
    # whole call for comparison
    xg2 = x.detach().requires_grad_(True)
    np.random.seed(0); torch.manual_seed(0)
    torch.cuda.synchronize(); t = time.perf_counter()
    tot, info = R.compute_loss(model, xg2, seg, text, sets, None, None, W_text=1.0, W_image=0.0, W_smooth=0.0,
                               percent_image_sampling=c["pct_sampling"], k_distractors=K - c["G"], pct_medium=0.0, pct_hard=1.0, pct_rand=0.0)
    tot.backward()
    t = tick("compute_loss + backward (one call, one sync)", t)
out = {k: float(np.median(v[1:])) for k, v in T.items()}
print(json.dumps(out, indent=1))
os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
json.dump(out, open(os.path.join(ROOT, "gpurun_out", "r2_api_breakdown.json"), "w"), indent=1)
