"""One warm-up + one measured launch of every HBM-bound kernel of the path (pooling, smoothness, row norms,
evaluation histograms, sampling weights), for an `ncu --set full` capture:

  ncu --set full --clock-control none --import-source on \
      -k regex:"pool_|tv_|rownorm|eval_hist|eval_fold|sample_|weight_sum" -o /tmp/prof_hbm python tools/prof_hbm.py
  ncu -i /tmp/prof_hbm.ncu-rep --page raw --csv > gpurun_out/prof_hbm_raw.csv      (the report itself is > 64 MiB)

Sizes: B=16 of the B=64 batch (X bf16 = 1.07 GB, f32 = 2.15 GB, both far above the 126 MB L2)."""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from rangeclip_b200 import _lib, ops

dev = torch.device("cuda:0")
B, D, H, W = 16, 512, 256, 256
g = torch.Generator(device=dev).manual_seed(0)
REPS = int(os.environ.get("PROF_REPS", "2"))      # 1 under ncu (every replay pass is cold-cache anyway)
L = _lib.lib()
st = torch.cuda.current_stream().cuda_stream
seg = torch.arange(64, device=dev).view(8, 8).repeat_interleave(32, 0).repeat_interleave(32, 1)[None].repeat(B, 1, 1).contiguous()
lut = torch.arange(B * 64, device=dev, dtype=torch.int32).view(B, 64)
scale = torch.tensor([1e-9, 1e-9], device=dev)
one = torch.ones(1, device=dev)

for dtype in (torch.bfloat16, torch.float32):
    x = torch.randn(B, D, H, W, device=dev, generator=g).to(dtype)
    rcdt = _lib.RC_F32 if dtype == torch.float32 else _lib.RC_BF16
    dx = torch.empty_like(x)
    out, cnt = ops.pool_forward(x, seg, lut, True, B * 64)
    gup = torch.randn_like(out)
    for rep in range(REPS):
        ops.tv_sums(x)
        _lib.check(L.rc_tv_bwd(x.data_ptr(), rcdt, B * D, H, W, scale.data_ptr(), dx.data_ptr(), 0, None, st), "tv_bwd")
        _lib.check(L.rc_tv_bwd(x.data_ptr(), rcdt, B * D, H, W, scale.data_ptr(), dx.data_ptr(), 1, one.data_ptr(), st), "tv_bwd")
        ops.pool_forward(x, seg, lut, True, B * 64)
        ops.pool_backward(gup, cnt, seg, lut, True, tuple(x.shape), dtype)
        ws_bytes = int(L.rc_infonce_workspace_bytes(B, D, H * W, 256, rcdt))
        ws = torch.empty(ws_bytes, device=dev, dtype=torch.uint8)
        _lib.check(L.rc_infonce_prepass(x.data_ptr(), rcdt, B, D, H * W, ws.data_ptr(), ws_bytes, st), "prepass")
        del ws
    del x, dx, out, gup

# evaluation histograms (K = 1024 vocabulary, top-5) and the sampling-weight kernels
C, k = 1024, 5
gt = (torch.randint(0, C, (B, H, W), device=dev, generator=g) // 37 * 37) % C
topk = torch.randint(0, C, (B, k, H, W), device=dev, generator=g)
topk[:, 0] = torch.where(torch.rand(B, H, W, device=dev, generator=g) < 0.5, gt, topk[:, 0])
E = torch.eye(C, dtype=torch.uint8, device=dev)
cmap = torch.arange(C, device=dev)
hist = torch.zeros(5, C, device=dev, dtype=torch.int64)
cnt3 = torch.zeros(3, device=dev, dtype=torch.int64)
acc = torch.zeros(4, C, device=dev, dtype=torch.int64)
first_seen = torch.full((C,), 2**31 - 1, device=dev, dtype=torch.int32)
label_map = torch.full((C,), -1, device=dev, dtype=torch.int32)
label_map[1:257] = torch.arange(256, device=dev, dtype=torch.int32)
rand_idx = torch.randint(0, H * W, (B, int(0.7 * H * W)), device=dev, generator=g)
for rep in range(REPS):
    hist.zero_()
    ops.eval_hist(gt, topk, E, cmap, hist, cnt3)
    ops.eval_fold(hist, rep, acc, first_seen)
    ops.sample_weights(gt % 257, rand_idx, label_map)
torch.cuda.synchronize()
print("ok")
