"""BASELINE configs[2]: the full training step -- the reference's own ResNet-18-UNet+ASPP backbone (unmodified, from
baseline/_ref, random init) + the fused text / area-image / smoothness losses of this repo -- under the reference's DDP
wrapping (train_util.py:173-175), B images per GPU, 1 / 2 / 4 / 8 GPUs:

    python bench.py --workload full_step --gpus 1 --steps 8 --warmup 3 [--batch 64] [--variant shared2x2|full|eager]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P bench.py \
        --workload full_step --gpus N --steps 8 --warmup 3

variant  shared2x2  compute_loss_shared2x2 on the decoder's output_conv result (the 128x128 distinct embeddings)
         full       compute_loss on the decoder's full-resolution fp32 output (decoder.py:113-115 as written)
         eager      the reference's own compute_loss (model.py:178-355) on the same tensors -- the GPU incumbent
One JSON line: step ms (max over ranks) and Mpix/s over all ranks.  --phase backbone times the same step with the loss
replaced by a mean (the loss path's share follows), --phase nosync the step under DDP.no_sync (no gradient all-reduce: the
exposed NCCL share follows); tools/run_full_step.sh runs the three phases and merges the lines."""
import json
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def decoder_pre_tail(decoder, spatial_feature_map, encoder_features):
    """The reference decoder up to (and including) output_conv -- utils/src/decoder.py:99-112, its own modules called in
    its own order -- without the nearest x2 / normalize tail (:113-115) that compute_loss_shared2x2 folds into the loss."""
    skip_features = encoder_features[:-1][::-1]
    x = decoder.up_blocks[0](spatial_feature_map)
    for i in range(1, len(decoder.up_blocks)):
        x = decoder.up_blocks[i](x, skip_features[i - 1])
    return decoder.output_conv(x)


class Backbone(torch.nn.Module):
    """The reference encoder + decoder behind ONE forward, so that DistributedDataParallel sees the call (its reducer is
    armed in DDP.forward); ``pre_tail`` stops after output_conv (decoder_pre_tail)."""

    def __init__(self, core, pre_tail):
        super().__init__()
        self.core, self.pre_tail = core, pre_tail

    def forward(self, depth):
        _, feats, final = self.core.depth_encoder(depth)
        if self.pre_tail:
            return decoder_pre_tail(self.core.depth_decoder, final, feats)             # [B, D, H/2, W/2]
        return self.core.depth_decoder(final, feats, depth.shape[2:])                   # [B, D, H, W] fp32, normalised


def run_full_step(args):
    import torch.distributed as dist
    import bench as BN
    import rangeclip_b200 as R
    from rangeclip_b200 import _lib
    from tools.stage_reference import import_reference_model
    rank, world, local = BN.dist_env()
    torch.cuda.set_device(local)
    device = torch.device("cuda", local)
    if os.environ.get("NCCL_DEBUG", "VERSION").upper() == "VERSION":
        os.environ["NCCL_DEBUG"] = "WARN"
    os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")
    if world > 1:
        dist.init_process_group("nccl", device_id=device)
    c = BN.CFG
    B, H, W, D, K, C = args.batch, c["H"], c["W"], c["D"], c["K"], c["C"]
    DepthUNet = import_reference_model()
    torch.manual_seed(0)                       # identical initial weights on every rank, as DDP would broadcast
    core = DepthUNet('resnet', device, embedding_dim=D, use_batch_norm=True, activation_func='relu').to(device)   # train_util.py:133-144
    net = core if args.variant == "eager" else Backbone(core, args.variant == "shared2x2")
    # train_util.py:173-174: DistributedDataParallel + _set_static_graph() (the static graph is what lets DDP live with the
    # temperatures used outside forward and the backbone parameters that never receive a gradient).  A static graph must
    # not change between iterations, so one process times ONE phase (--phase step | backbone | nosync).
    if world > 1 and args.phase != "nosync":
        model = torch.nn.parallel.DistributedDataParallel(net, device_ids=[local])
        model._set_static_graph()
    else:       # --phase nosync: the same step on every rank WITHOUT the wrapper = the step minus the gradient all-reduce
        model = net      # (DDP.no_sync() is not available under a static graph)
    opt = torch.optim.Adam(core.parameters(), lr=1e-4)
    g = torch.Generator(device=device).manual_seed(1234 + rank)
    depth = torch.rand(B, 1, H, W, device=device, generator=g) + 0.5
    seg = BN.make_labels(B, H, W, c["G"], torch.Generator().manual_seed(1234 + rank)).to(device)
    text = torch.nn.functional.normalize(torch.randn(C, D, device=device, generator=g), dim=1)
    rest = torch.arange(c["G"] + 1, C, device=device)
    distract = rest[torch.randperm(rest.numel(), device=device, generator=g)[: K - c["G"]]]
    contrast = torch.unique(torch.cat([torch.arange(1, c["G"] + 1, device=device), distract]))
    sets = BN.similarity_sets(contrast.tolist(), c["G"], C)
    img = torch.nn.functional.normalize(torch.randn(B, D, device=device, generator=g), dim=1)       # frozen CLIP crop embeddings
    obj_labels = seg[:, H // 2, W // 2].tolist()                                                     # one object per image
    loss_kw = dict(W_text=1.0, W_image=0.5, W_smooth=2e2, percent_image_sampling=c["pct_sampling"], k_distractors=K - c["G"],
                   pct_medium=0.0, pct_hard=1.0, pct_rand=0.0)
    ours_kw = dict(loss_kw, contrast_builder=args.builder)
    scaler = torch.amp.GradScaler("cuda") if args.variant == "eager" else None
    last = {}

    def forward_backward(mode, sync=True):
        """mode: 'loss' = the variant's loss path, 'backbone' = same backbone work with the loss replaced by a mean."""
        if True:
            if args.variant == "eager":
                # the reference as written: DepthUNet.forward (fp16 autocast inside, model.py:110), its own compute_loss, GradScaler
                emb, _, _ = model(depth)
                if mode == "backbone":
                    loss = emb.float().mean() + 0.0 * (core.log_temperature_text + core.log_temperature_image)
                else:
                    with torch.no_grad():
                        area = torch.stack([(emb[b].float() * (seg[b] == obj_labels[b])[None]).sum(dim=(1, 2)) /
                                            (seg[b] == obj_labels[b]).sum().clamp_min(1) for b in range(B)])
                    with torch.autocast("cuda", dtype=torch.float16):
                        loss, info = core.compute_loss(emb, seg, text, sets, area, img, **loss_kw)
                    last["info"] = info
                scaler.scale(loss).backward()
                return
            with torch.autocast("cuda", dtype=torch.bfloat16):      # bf16 autocast, no GradScaler
                e = model(depth)
            if mode == "backbone":      # same parameter set receives gradients (DDP expects every parameter each step)
                loss = e.float().mean() + 0.0 * (core.log_temperature_text + core.log_temperature_image)
            elif args.variant == "shared2x2":
                with torch.no_grad():
                    area = R.pool_objects_per_image(e, seg, list(range(B)), obj_labels, shared2x2=True)
                loss, info = R.compute_loss_shared2x2(core, e, seg, text, sets, area, img, **ours_kw)
                last["info"] = info          # (a LazyLossInfo under the device builder: read once, after the timed loop)
            else:
                with torch.no_grad():
                    area = R.pool_objects_per_image(e, seg, list(range(B)), obj_labels)
                loss, info = R.compute_loss(core, e, seg, text, sets, area, img, **ours_kw)
                last["info"] = info
            loss.backward()

    def train_step(mode="loss", sync=True):
        forward_backward(mode, sync)
        if scaler is not None:
            scaler.step(opt)
            scaler.update()
        else:
            opt.step()
        opt.zero_grad(set_to_none=True)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, steps, warmup):
        for _ in range(warmup):
            fn()
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(steps):
            fn()
        e1.record()
        barrier()
        t = torch.tensor([e0.elapsed_time(e1)], device=device, dtype=torch.float64)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t) / steps

    np.random.seed(0)
    n0 = _lib.launch_count()
    phase = args.phase
    with BN.ClockSampler(local) as clk:
        if phase == "backbone":
            t_full = timed(lambda: train_step("backbone"), args.steps, max(args.warmup, 3))
        elif phase == "nosync":
            t_full = timed(lambda: train_step("loss", sync=False), args.steps, max(args.warmup, 3))
        else:
            t_full = timed(train_step, args.steps, max(args.warmup, 3))
    launches = (_lib.launch_count() - n0)
    n_params = sum(p.numel() for p in core.parameters())
    if rank == 0:
        line = {
            "metric": "full_training_step_throughput", "value": world * B * H * W / (t_full * 1e-3) / 1e6, "unit": "Mpix/s", "n_gpus": world,
            "steps": args.steps, "warmup": max(args.warmup, 3), "ms_per_step": t_full, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "fp16 autocast (reference)" if args.variant == "eager" else "bf16", "data": "synthetic",
            "phase": phase,
            "config": {"workload": f"configs[2]: reference ResNet-18-UNet+ASPP fwd+bwd + hybrid loss (text K=256 + area-image n=B + smoothness) + Adam, B={B}/GPU, 256x256, D=512",
                       "variant": args.variant, "contrast_builder": "reference (model.py:231-270 as written)" if args.variant == "eager" else args.builder,
                       "parallelism": f"ddp{world}" if world > 1 else "single"},
            "clocks": clk.summary(), "gpu_launches": int(launches),
            "grad_bytes_allreduced_per_step": n_params * 4 if world > 1 else 0, "params_M": n_params / 1e6,
            "loss_info": {k: v for k, v in dict(last.get("info", {})).items() if isinstance(v, (int, float))},
        }
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()
