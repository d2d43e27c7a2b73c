import os, sys, json, torch
sys.path.insert(0, "/root/repo"); os.chdir("/root/repo")
import bench
from rangeclip_b200 import _lib, ops
dev = torch.device("cuda:0"); torch.cuda.set_device(0)
c = bench.CFG; B, D, H, W, K = c["B"], c["D"], c["H"], c["W"], c["K"]; HW = H*W; M = B*HW
wl = bench.make_device_workload(dev, 1234, B)
x, text, seg = wl["x"], wl["text"], wl["seg"]
L = _lib.lib(); st = torch.cuda.current_stream().cuda_stream
label_map = torch.full((c["C"],), -1, dtype=torch.int32, device=dev); label_map[wl["contrast"]] = torch.arange(K, device=dev, dtype=torch.int32)
_, tb, ttb = ops.text_prepare(text, wl["contrast"], want_f32=False, want_bf16=True)
ws_bytes = int(L.rc_infonce_workspace_bytes(B, D, HW, K, _lib.RC_BF16)); ws = torch.empty(ws_bytes, device=dev, dtype=torch.uint8)
dx = torch.empty(B, D, HW, device=dev, dtype=torch.bfloat16); lse = torch.empty(M, device=dev); acc = torch.zeros(4, device=dev, dtype=torch.float64)
w, y = ops.sample_weights(seg.view(B, HW), wl["rand_idx"], label_map)
_lib.check(L.rc_weight_sum(w.data_ptr(), y.data_ptr(), M, acc[3:].data_ptr(), st), "ws")
def k(xx, yy, ww):
    _lib.check(L.rc_infonce_bf16(xx.data_ptr(), _lib.RC_BF16, B, D, HW, tb.data_ptr(), ttb.data_ptr(), K, yy.data_ptr(), ww.data_ptr(), 1/0.07, lse.data_ptr(), acc[0:].data_ptr(), acc[1:].data_ptr(), acc[3:].data_ptr(), None, dx.data_ptr(), None, acc[2:].data_ptr(), ws.data_ptr(), ws_bytes, 1, st), "k")
def timeit(fn, n=10, sync_each=False):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    tot = 0.0
    if sync_each:
        for _ in range(n):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(); fn(); e1.record(); torch.cuda.synchronize(); tot += e0.elapsed_time(e1)
        return tot / n
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n
yr = torch.randint(0, K, (M,), device=dev, dtype=torch.int32); wo = torch.ones(M, device=dev)
xr = torch.empty_like(x)
g = torch.Generator(device=dev).manual_seed(0)
for b in range(B): xr[b] = torch.randn(D, H, W, device=dev, generator=g).to(torch.bfloat16)
import subprocess
def smi():
    return subprocess.run(["nvidia-smi","--query-gpu=clocks.sm,clocks.mem,power.draw,temperature.gpu,clocks_throttle_reasons.active","--format=csv,noheader"],capture_output=True,text=True).stdout.strip()
for rep in range(3):
    print("randn x, random y, ones w  ", round(timeit(lambda: k(xr, yr, wo)),3), smi())
    print("bench workload back-to-back", round(timeit(lambda: k(x, y, w)),3), smi())
    print("bench x, random y, ones w  ", round(timeit(lambda: k(x, yr, wo)),3))
    print("bench x, bench y, ones w   ", round(timeit(lambda: k(x, y, wo)),3))
    print("bench x, random y, bench w ", round(timeit(lambda: k(x, yr, w)),3))
    print("bench workload sync each   ", round(timeit(lambda: k(x, y, w), sync_each=True),3))
