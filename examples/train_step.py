#!/usr/bin/env python
"""One-process-per-GPU training steps with the drop-in losses (BASELINE configs[2] shape: B images of 256x256 per GPU,
D=512, K = ground-truth labels + curriculum distractors), on a stand-in backbone (examples/standin_model.py).

    python examples/train_step.py [--batch 16] [--steps 5] [--shared2x2]
    python -m torch.distributed.run --nproc-per-node 2 --master-addr 127.0.0.1 examples/train_step.py --ddp

Follows the reference loop (train_util.py:286-345) minus what it does not need on this path: bf16 autocast without a
GradScaler, no per-step ``torch.cuda.empty_cache()``.  Prints per-step time and the share of the loss path.
"""
import argparse
import os
import sys
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import rangeclip_b200 as R                                  # noqa: E402
from examples.standin_model import StandInDepthUNet         # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--batch", type=int, default=16)
    ap.add_argument("--size", type=int, default=256)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--classes", type=int, default=1024)
    ap.add_argument("--shared2x2", action="store_true", help="take the loss below the decoder tail (compute_loss_shared2x2)")
    ap.add_argument("--ddp", action="store_true")
    args = ap.parse_args()
    rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
    torch.cuda.set_device(local)
    device = torch.device("cuda", local)
    if args.ddp:
        torch.distributed.init_process_group("nccl", device_id=device)
    torch.manual_seed(1234 + rank)
    model = StandInDepthUNet().to(device)
    net = torch.nn.parallel.DistributedDataParallel(model, device_ids=[local]) if args.ddp else model
    opt = torch.optim.AdamW(net.parameters(), lr=1e-4)
    B, S, C = args.batch, args.size, args.classes
    text = torch.nn.functional.normalize(torch.randn(C, 512, device=device), dim=1)
    sets = {"medium": {}, "hard": {i: [(i * 7 + j) % C for j in range(1, 9)] for i in range(C)}}
    g = torch.Generator(device=device).manual_seed(rank)
    for step in range(args.steps):
        depth = torch.rand(B, 1, S, S, device=device, generator=g) + 0.5
        seg = torch.randint(0, 64, (B, S // 32, S // 32), device=device, generator=g).repeat_interleave(32, 1).repeat_interleave(32, 2)
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        with torch.autocast("cuda", dtype=torch.bfloat16):
            emb, _, _ = net(depth, skip_tail=args.shared2x2)
        torch.cuda.synchronize()
        t1 = time.perf_counter()
        # W_image keeps the reference default: without area embeddings the image term is `dummy * exp(log_tau_image) * 0`
        # (model.py:325-326), which still gives log_temperature_image a (zero) gradient -- DDP wants one for every parameter
        loss_fn = R.compute_loss_shared2x2 if args.shared2x2 else R.compute_loss
        loss, info = loss_fn(model, emb, seg, text, sets, None, None, W_text=1.0, W_image=0.5, W_smooth=2e2, k_distractors=192,
                             pct_medium=0.0, pct_hard=0.75, pct_rand=0.25)
        torch.cuda.synchronize()
        t2 = time.perf_counter()
        opt.zero_grad(set_to_none=True)
        loss.backward()
        opt.step()
        torch.cuda.synchronize()
        t3 = time.perf_counter()
        if rank == 0:
            print(f"step {step}: loss {info['total_loss']:.4f} (text {info['text_contrastive_loss']:.4f}, smooth {info['smoothness_loss']:.5f}) "
                  f"backbone fwd {1e3 * (t1 - t0):.1f} ms, losses fwd {1e3 * (t2 - t1):.1f} ms, backward + optimizer {1e3 * (t3 - t2):.1f} ms, "
                  f"{world * B * S * S / (t3 - t0) / 1e6:.1f} Mpix/s", flush=True)
    if args.ddp:
        torch.distributed.destroy_process_group()


if __name__ == "__main__":
    main()
