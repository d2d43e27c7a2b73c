import sys, torch, time
import os; sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from rangeclip_b200 import ops
dev = torch.device("cuda:0")
g = torch.Generator(device=dev).manual_seed(0)
B, D, H, W = 16, 512, 256, 256
x = torch.randn(B, D, H, W, device=dev, generator=g).to(torch.bfloat16)
for K in (256, 1024):
    t = torch.nn.functional.normalize(torch.randn(K, D, device=dev, generator=g), dim=1)
    tb = ops.text_to_bf16(t)[0]
    im = torch.arange(K, device=dev)
    for k in (1, 5):
        ops.eval_topk(x, t, im, k, "bf16", t_bf16=tb); torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(3): ops.eval_topk(x, t, im, k, "bf16", t_bf16=tb)
        e1.record(); torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 3
        print(f"K={K} k={k}: {ms:.3f} ms  {B*H*W/ms/1e3:.1f} Mpix/s  {2*K*D*B*H*W/ms/1e9:.1f} TFLOP/s", flush=True)
