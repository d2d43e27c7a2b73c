"""Per-kernel throughput of the HBM-bound kernels at BASELINE config sizes (CUDA events, inputs > L2).
Prints one JSON line per kernel with achieved algorithmic GB/s against the measured HBM peak."""
import json, os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from rangeclip_b200 import _lib, ops
import rangeclip_b200 as R

dev = torch.device("cuda:0")
peak = json.load(open(os.path.join(os.path.dirname(__file__), "..", "MEASURED_PEAKS.json")))["hbm_gbs"] if os.path.exists(
    os.path.join(os.path.dirname(__file__), "..", "MEASURED_PEAKS.json")) else 6650.0
B = int(sys.argv[1]) if len(sys.argv) > 1 else 64
D, H, W = 512, 256, 256
g = torch.Generator(device=dev).manual_seed(0)


def timeit(fn, reps=5, warm=2):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


def report(name, ms, bytes_alg, extra=None):
    gbs = bytes_alg / (ms * 1e-3) / 1e9
    d = {"kernel": name, "ms": round(ms, 4), "algorithmic_GB": round(bytes_alg / 1e9, 3), "GBps": round(gbs, 1),
         "frac_of_measured_hbm": round(gbs / peak, 3), "Gpix_per_s": round(B * H * W / (ms * 1e-3) / 1e9, 3)}
    if extra:
        d.update(extra)
    print(json.dumps(d), flush=True)


for dtype, esz in ((torch.bfloat16, 2), (torch.float32, 4)):
    x = torch.empty(B, D, H, W, device=dev, dtype=dtype)
    for b in range(B):
        x[b] = torch.randn(D, H, W, device=dev, generator=g).to(dtype)
    n = x.numel()
    # smoothness forward / backward
    report(f"tv_fwd[{dtype}]", timeit(lambda: ops.tv_sums(x)), n * esz)
    scale = torch.tensor([1e-9, 1e-9], device=dev)
    dx = torch.empty_like(x)
    L = _lib.lib(); st = torch.cuda.current_stream().cuda_stream
    rcdt = _lib.RC_F32 if dtype == torch.float32 else _lib.RC_BF16
    report(f"tv_bwd[{dtype}]", timeit(lambda: _lib.check(L.rc_tv_bwd(x.data_ptr(), rcdt, B * D, H, W, scale.data_ptr(),
                                                                     dx.data_ptr(), 0, None, st), "tv_bwd")), 2 * n * esz)
    one = torch.ones(1, device=dev)
    report(f"tv_bwd_accumulate[{dtype}]", timeit(lambda: _lib.check(L.rc_tv_bwd(x.data_ptr(), rcdt, B * D, H, W, scale.data_ptr(),
                                                                                dx.data_ptr(), 1, one.data_ptr(), st), "tv_bwd")), 3 * n * esz)
    del dx
    # pooling: config 3 -- 64 object masks per image (8x8 grid of 32x32 blocks), n = 64 B slots
    seg = torch.arange(64, device=dev).view(8, 8).repeat_interleave(32, 0).repeat_interleave(32, 1)[None].repeat(B, 1, 1).contiguous()
    lut = torch.arange(B * 64, device=dev, dtype=torch.int32).view(B, 64)
    report(f"pool_fwd[{dtype}]", timeit(lambda: ops.pool_forward(x, seg, lut, True, B * 64)), n * esz + B * H * W * 8,
           {"slots": B * 64})
    out, cnt = ops.pool_forward(x, seg, lut, True, B * 64)
    gup = torch.randn_like(out)
    report(f"pool_bwd[{dtype}]", timeit(lambda: ops.pool_backward(gup, cnt, seg, lut, True, tuple(x.shape), dtype)),
           n * esz + B * H * W * 8)
    if dtype == torch.bfloat16:
        ws_bytes = int(L.rc_infonce_workspace_bytes(B, D, H * W, 256, rcdt))
        ws = torch.empty(ws_bytes, device=dev, dtype=torch.uint8)
        report("infonce_prepass[bf16]", timeit(lambda: _lib.check(L.rc_infonce_prepass(x.data_ptr(), rcdt, B, D, H * W, ws.data_ptr(),
                                                                                      ws_bytes, st), "prepass")), n * esz + B * H * W * 4)
        del ws
    del x, out, gup

# evaluation histograms: K = 1024 vocabulary, top-5
C, k = 1024, 5
gt = torch.randint(0, C, (B, H, W), device=dev, generator=g)
gt = (gt // 37 * 37) % C
blocky = gt.view(B, H // 32, 32, W // 32, 32)[:, :, :1, :, :1].expand(B, H // 32, 32, W // 32, 32).reshape(B, H, W).contiguous()
topk = torch.randint(0, C, (B, k, H, W), device=dev, generator=g)
topk[:, 0] = torch.where(torch.rand(B, H, W, device=dev, generator=g) < 0.5, blocky, topk[:, 0])
E = torch.eye(C, dtype=torch.uint8, device=dev); cmap = torch.arange(C, device=dev)
hist = torch.zeros(5, C, device=dev, dtype=torch.int64); cnt = torch.zeros(3, device=dev, dtype=torch.int64)
report("eval_hist[K=1024,k=5]", timeit(lambda: ops.eval_hist(blocky, topk, E, cmap, hist, cnt)), B * H * W * 48)
x32 = torch.randn(2, D, H, W, device=dev, generator=g)
tn = torch.nn.functional.normalize(torch.randn(C, D, device=dev, generator=g), dim=1)
ms = timeit(lambda: ops.eval_topk(x32, tn, torch.arange(C, device=dev), 5), reps=2, warm=1)
print(json.dumps({"kernel": "eval_topk_f32[K=1024] (CUDA-core fp32 path, B=2)", "ms": round(ms, 3),
                  "Mpix_per_s": round(2 * H * W / (ms * 1e-3) / 1e6, 2)}))
