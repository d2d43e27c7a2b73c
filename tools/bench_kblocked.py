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
