"""fp32 X at the headline size: the separate pre-pass (bf16 copy + row norms) and smoothness-sum kernels against the fused
rc_infonce_prepass_tv (one read of X).   python tools/bench_prepass.py   [RANGECLIP_B200_LIB=<variant .so>]"""
import os, sys, json
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from rangeclip_b200 import _lib, ops
dev = torch.device("cuda:0")
B, D, H, W = 64, 512, 256, 256
x = torch.empty(B, D, H, W, device=dev)
g = torch.Generator(device=dev).manual_seed(0)
for b in range(B):
    x[b] = torch.randn(D, H, W, device=dev, generator=g)
L = _lib.lib(); st = torch.cuda.current_stream().cuda_stream
wsb = int(L.rc_infonce_workspace_bytes(B, D, H * W, 256, _lib.RC_F32)); ws = torch.empty(wsb, device=dev, dtype=torch.uint8)
tv = torch.zeros(2, device=dev, dtype=torch.float64)
def t(fn, n=5):
    for _ in range(2): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n
gb = x.numel() * 4 / 1e9
out = {"lib": os.path.basename(_lib.LIB_PATH)}
out["prepass_ms"] = t(lambda: _lib.check(L.rc_infonce_prepass(x.data_ptr(), _lib.RC_F32, B, D, H * W, ws.data_ptr(), wsb, st), "prepass"))
out["tv_fwd_ms"] = t(lambda: ops.tv_sums(x))
out["fused_ms"] = t(lambda: _lib.check(L.rc_infonce_prepass_tv(x.data_ptr(), B, D, H, W, ws.data_ptr(), wsb, tv.data_ptr(), None, st), "prepass_tv"))
codes = torch.empty(B, D, H, W // 8, device=dev, dtype=torch.int32)
out["fused_codes_ms"] = t(lambda: _lib.check(L.rc_infonce_prepass_tv(x.data_ptr(), B, D, H, W, ws.data_ptr(), wsb, tv.data_ptr(), codes.data_ptr(), st), "prepass_tv"))
out["fused_GBps_algorithmic"] = 1.5 * gb / out["fused_ms"] * 1e3
out["separate_GBps_algorithmic"] = 2.5 * gb / (out["prepass_ms"] + out["tv_fwd_ms"]) * 1e3
tv.zero_(); _lib.check(L.rc_infonce_prepass_tv(x.data_ptr(), B, D, H, W, ws.data_ptr(), wsb, tv.data_ptr(), codes.data_ptr(), st), "prepass_tv")
ref = ops.tv_sums(x)
out["rel_err"] = float(((tv - ref).abs() / ref).max())
del ws
dxb = torch.randn(B, D, H, W, device=dev, dtype=torch.bfloat16)
scale = torch.tensor([1e-3, 2e-3], device=dev); one = torch.ones((), device=dev)
out["tv_bwd_from_x_ms"] = t(lambda: ops.tv_backward(x, scale, dxb, one), 3)
out["tv_bwd_codes_ms"] = t(lambda: ops.tv_backward_codes(codes, scale, dxb, one), 3)
a = ops.tv_backward(x, scale, dxb, one); bq = ops.tv_backward_codes(codes, scale, dxb, one)
out["bwd_max_abs_diff"] = float((a - bq).abs().max())
print(json.dumps(out))
