"""A few launches of the fused InfoNCE backward kernel at the headline size, for ncu (TS kernel by default; SS=1 selects
the shared-memory-operand kernel).  python tools/prof_ts.py"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from rangeclip_b200 import _lib, ops  # noqa: E402

dev = torch.device("cuda")
L = _lib.lib()
st = torch.cuda.current_stream().cuda_stream
B, D, HW, K = int(os.environ.get("PROF_B", 64)), 512, 65536, 256
g = torch.Generator(device=dev).manual_seed(1)
x = torch.empty(B, D, HW, device=dev, dtype=torch.bfloat16)
for b in range(B):
    xb = torch.randn(D, HW, device=dev, generator=g)
    x[b] = (xb / xb.norm(dim=0, keepdim=True)).to(torch.bfloat16)
text = torch.nn.functional.normalize(torch.randn(K, D, device=dev, generator=g), dim=1)
_, tb, ttb = ops.text_prepare(text, None, want_f32=False, want_bf16=True)
y = torch.randint(0, K, (B * HW,), device=dev, generator=g, dtype=torch.int32)
w = torch.randint(0, 3, (B * HW,), device=dev, generator=g).float()
M = B * HW
acc = torch.zeros(4, device=dev, dtype=torch.float64)
lse = torch.empty(M, device=dev, dtype=torch.float32)
dx = torch.empty(B, D, HW, device=dev, dtype=torch.bfloat16)
ws_bytes = int(L.rc_infonce_workspace_bytes(B, D, HW, K, _lib.RC_BF16))
ws = torch.empty(ws_bytes, device=dev, dtype=torch.uint8)
_lib.check(L.rc_weight_sum(w.data_ptr(), y.data_ptr(), M, acc[3:].data_ptr(), st), "wsum")
flags = 0 if os.environ.get("SS") else 8
for _ in range(int(os.environ.get("PROF_REPS", 4))):
    _lib.check(L.rc_infonce_bf16(x.data_ptr(), _lib.RC_BF16, B, D, HW, tb.data_ptr(), ttb.data_ptr(), K, y.data_ptr(), w.data_ptr(),
                                 1.0 / 0.07, lse.data_ptr(), acc[0:].data_ptr(), acc[1:].data_ptr(), acc[3:].data_ptr(), None,
                                 dx.data_ptr(), None, acc[2:].data_ptr(), ws.data_ptr(), ws_bytes, flags, st), "rc_infonce_bf16")
torch.cuda.synchronize()
print("ok", float(acc[0] / acc[1]))
