import sys, os, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from rangeclip_b200 import ops
dev = torch.device("cuda:0")
B, D, H, W = 16, 512, 256, 256
g = torch.Generator(device=dev).manual_seed(0)
x = torch.randn(B, D, H, W, device=dev, generator=g).to(torch.bfloat16)
seg = torch.arange(64, device=dev).view(8, 8).repeat_interleave(32, 0).repeat_interleave(32, 1)[None].repeat(B, 1, 1).contiguous()
lut = torch.arange(B * 64, device=dev, dtype=torch.int32).view(B, 64)
for _ in range(3):
    ops.pool_forward(x, seg, lut, True, B * 64)
    ops.tv_sums(x)
torch.cuda.synchronize()
