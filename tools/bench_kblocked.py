import sys, torch, time
sys.path.insert(0, ".")
from rangeclip_b200 import ops
dev = torch.device("cuda:0")
g = torch.Generator(device=dev).manual_seed(0)
B, D, HW = 64, 512, 65536
x = torch.empty(B, D, HW, device=dev, dtype=torch.bfloat16)
for b in range(B): x[b] = torch.randn(D, HW, device=dev, generator=g).to(torch.bfloat16)
for K in (256, 300, 512):
    t = torch.nn.functional.normalize(torch.randn(K, D, device=dev, generator=g), dim=1)
    y = torch.randint(0, K, (B * HW,), device=dev, generator=g, dtype=torch.int32); w = torch.ones(B * HW, device=dev)
    for _ in range(2): r = ops.infonce_raw(x, t, y, w, 1 / 0.07, True, False, "auto")
    torch.cuda.synchronize(); t0 = time.perf_counter()
    for _ in range(3): r = ops.infonce_raw(x, t, y, w, 1 / 0.07, True, False, "auto")
    torch.cuda.synchronize(); print(K, r["precision"], f"{(time.perf_counter() - t0) / 3 * 1e3:.2f} ms", float(r["loss_sum"] / r["w_sum"]), f"{torch.cuda.max_memory_allocated() / 1e9:.1f} GB")

# the pieces of the K-blocked form: one forward-only (logsumexp) launch and one backward launch of a 256-row block
from rangeclip_b200 import _lib
L = _lib.lib(); st = torch.cuda.current_stream().cuda_stream
K = 256
t = torch.nn.functional.normalize(torch.randn(K, D, device=dev, generator=g), dim=1)
_, tb, ttb = ops.text_prepare(t, None, want_f32=False, want_bf16=True)
y = torch.randint(0, K, (B * HW,), device=dev, generator=g, dtype=torch.int32); w = torch.ones(B * HW, device=dev)
M = B * HW
acc = torch.zeros(4, device=dev, dtype=torch.float64); lse = torch.empty(M, device=dev); acc[3] = M
dx = torch.empty(B, D, HW, device=dev, dtype=torch.bfloat16)
wsb = int(L.rc_infonce_workspace_bytes(B, D, HW, K, _lib.RC_BF16)); ws = torch.empty(wsb, device=dev, dtype=torch.uint8)
def fwd(): _lib.check(L.rc_infonce_bf16(x.data_ptr(), _lib.RC_BF16, B, D, HW, tb.data_ptr(), ttb.data_ptr(), K, y.data_ptr(), w.data_ptr(), 1 / 0.07,
                                       lse.data_ptr(), acc[0:].data_ptr(), acc[1:].data_ptr(), None, None, None, None, None, ws.data_ptr(), wsb, 1 | 2, st), "f")
def bwd(fl): _lib.check(L.rc_infonce_bf16(x.data_ptr(), _lib.RC_BF16, B, D, HW, tb.data_ptr(), ttb.data_ptr(), K, y.data_ptr(), w.data_ptr(), 1 / 0.07,
                                          lse.data_ptr(), acc[0:].data_ptr(), acc[1:].data_ptr(), acc[3:].data_ptr(), None, dx.data_ptr(), None,
                                          acc[2:].data_ptr(), ws.data_ptr(), wsb, fl, st), "b")
def tm(fn, n=5):
    for _ in range(2): fn()
    torch.cuda.synchronize(); e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize(); return e0.elapsed_time(e1) / n
print("forward-only launch (lse)", round(tm(fwd), 3), "ms")
print("backward launch, lse given, plain store", round(tm(lambda: bwd(1 | 2 | 4)), 3), "ms")
print("backward launch, lse given, reduce-add store", round(tm(lambda: bwd(1 | 2 | 4 | 16)), 3), "ms")
print("plain fused launch", round(tm(lambda: bwd(0)), 3), "ms")
