import sys, torch
import os; sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from rangeclip_b200 import ops
dev = torch.device("cuda:0")
g = torch.Generator(device=dev).manual_seed(0)
B, D, H, W, K, k = 8, 512, 256, 256, 1024, int(sys.argv[1]) if len(sys.argv) > 1 else 5
x = torch.randn(B, D, H, W, device=dev, generator=g).to(torch.bfloat16)
t = torch.nn.functional.normalize(torch.randn(K, D, device=dev, generator=g), dim=1)
tb = ops.text_to_bf16(t)[0]
im = torch.arange(K, device=dev)
for _ in range(3):
    ops.eval_topk(x, t, im, k, "bf16", t_bf16=tb)
torch.cuda.synchronize()
