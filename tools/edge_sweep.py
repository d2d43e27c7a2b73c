"""Edge-shape sweep of every op against the CPU oracle (tiny / odd / very wide / many-image shapes that the parity tests
do not enumerate); prints one line per case and the number of failures.  Run on a GPU box: python tools/edge_sweep.py"""
import sys, torch, numpy as np, traceback
sys.path.insert(0, ".")
from rangeclip_b200 import ops
import rangeclip_b200 as R
from oracle import rangeclip_oracle as O
dev = torch.device("cuda:0")
def maxrel(a, b):
    a = np.asarray(a, dtype=np.float64); b = np.asarray(b, dtype=np.float64)
    return float(np.abs(a - b).max() / max(np.abs(b).max(), 1e-30))
bad = 0
def run(name, fn):
    global bad
    try:
        r = fn()
        print("ok  ", name, r)
    except Exception as e:
        bad += 1
        print("FAIL", name, repr(e)[:300])
g = torch.Generator().manual_seed(0)
# smoothness: wide / tall / tiny
for shape, dt in [((1, 1, 3, 2048), torch.bfloat16), ((1, 1, 2, 4096), torch.float32), ((1, 2, 1, 8), torch.float32), ((1, 2, 8, 1), torch.float32),
                  ((1, 1, 300, 8), torch.bfloat16), ((2, 3, 7, 2040), torch.bfloat16), ((1, 1, 2, 6000), torch.float32)]:
    def f(shape=shape, dt=dt):
        x = torch.randn(shape, generator=g).to(dt)
        ref = O.smoothness(x.double()); gr = O.smoothness_grad(x.float())
        xd = x.to(dev).requires_grad_(True); l = ops.smoothness(xd); l.backward()
        e1 = abs(float(l) - float(ref)) / max(abs(float(ref)), 1e-30) if np.isfinite(float(ref)) else (0.0 if not np.isfinite(float(l)) else 1.0)
        return e1, maxrel(xd.grad.float().cpu(), gr)
    run(f"tv {shape} {dt}", f)
# pooling tiny / odd
for (B, D, H, W), dt in [((1, 1, 1, 8), torch.float32), ((1, 3, 2, 4), torch.bfloat16), ((2, 65, 4, 64), torch.bfloat16), ((1, 520, 2, 136), torch.float32),
                          ((1, 8, 1, 264), torch.bfloat16), ((3, 64, 1, 520), torch.bfloat16)]:
    def f(B=B, D=D, H=H, W=W, dt=dt):
        x = torch.randn(B, D, H, W, generator=g).to(dt).float()
        seg = torch.randint(0, 4, (B, H, W), generator=g)
        items = [b for b in range(B) for _ in range(4)]; labels = [l for _ in range(B) for l in range(4)]
        ref = O.area_pool_per_image(x.double(), seg, items, labels)
        out = R.pool_objects_per_image(x.to(dev).to(dt), seg.to(dev), items, labels)
        return maxrel(out.float().cpu(), ref)
    run(f"pool {(B, D, H, W)} {dt}", f)
# infonce bf16 extremes
for (B, D, H, W, K) in [(1, 256, 1, 8, 1), (1, 512, 1, 8, 256), (300, 256, 1, 8, 5), (1, 512, 3, 8, 255), (2, 128, 1, 8, 2), (1, 384, 2, 64, 65), (1, 512, 1, 136, 129)]:
    def f(B=B, D=D, H=H, W=W, K=K):
        x = torch.randn(B, D, H, W, generator=g).to(torch.bfloat16).float()
        t = torch.nn.functional.normalize(torch.randn(K, D, generator=g), dim=1).to(torch.bfloat16).float()
        y = torch.randint(-1, K, (B, H * W), generator=g, dtype=torch.int32); w = torch.randint(0, 3, (B, H * W), generator=g).float()
        ref = O.infonce_dense(x.permute(0, 2, 3, 1).reshape(-1, D), t, y.reshape(-1), w.reshape(-1), 1 / 0.07)
        r = ops.infonce_raw(x.to(dev).to(torch.bfloat16), t.to(dev), y.to(dev), w.to(dev), 1 / 0.07, True, False, "bf16")
        torch.cuda.synchronize()
        loss = float(r["loss_sum"] / r["w_sum"]) if float(r["w_sum"]) > 0 else 0.0
        dref = ref["dx"].reshape(B, H, W, D).permute(0, 3, 1, 2)
        return abs(loss - float(ref["loss"])), maxrel(r["dx"].float().cpu(), dref) if float(dref.abs().max()) > 1e-12 else float(r["dx"].float().abs().max())
    run(f"infonce bf16 {(B, D, H, W, K)}", f)
# rep4 extremes
for (B, D, h, w, K) in [(1, 256, 1, 8, 3), (130, 512, 1, 8, 256), (1, 512, 5, 40, 77)]:
    def f(B=B, D=D, h=h, w=w, K=K):
        x = torch.randn(B, D, h, w, generator=g).to(torch.bfloat16).float()
        t = torch.nn.functional.normalize(torch.randn(K, D, generator=g), dim=1).to(torch.bfloat16).float()
        M = B * h * w
        y4 = torch.randint(-1, K, (M, 4), generator=g, dtype=torch.int32); w4 = torch.randint(0, 3, (M, 4), generator=g).float()
        ref = O.infonce_dense_rep(x.permute(0, 2, 3, 1).reshape(-1, D), t, y4, w4, 1 / 0.07)
        r = ops.infonce_raw(x.to(dev).to(torch.bfloat16), t.to(dev), y4.to(dev), w4.to(dev), 1 / 0.07, True, False, "bf16", rep=4)
        torch.cuda.synchronize()
        loss = float(r["loss_sum"] / r["w_sum"])
        dref = ref["dx"].reshape(B, h, w, D).permute(0, 3, 1, 2)
        return abs(loss - float(ref["loss"])), maxrel(r["dx"].float().cpu(), dref)
    run(f"rep4 {(B, D, h, w, K)}", f)
# eval topk extremes
for (B, D, H, W, K, k) in [(1, 64, 1, 8, 1, 1), (1, 512, 1, 8, 5, 5), (2, 128, 2, 8, 2000, 5), (1, 256, 1, 136, 257, 8), (70, 64, 1, 8, 300, 3)]:
    def f(B=B, D=D, H=H, W=W, K=K, k=k):
        x = torch.randn(B, D, H, W, generator=g).to(torch.bfloat16)
        t = torch.nn.functional.normalize(torch.randn(K, D, generator=g), dim=1).to(torch.bfloat16).float()
        out = ops.eval_topk(x.to(dev), t.to(dev), torch.arange(K, device=dev), k, "bf16").cpu()
        logits = torch.einsum("bdn,cd->bcn", x.float().view(B, D, H * W), t)
        ref = logits.topk(k, dim=1)
        got = logits.gather(1, out.view(B, k, H * W))
        return float((got - ref.values).abs().max())
    run(f"topk {(B, D, H, W, K, k)}", f)
# sample weights degenerate
def f():
    seg = torch.randint(0, 5, (2, 64), generator=g).to(dev)
    lm = torch.arange(5, dtype=torch.int32).to(dev) - 1
    w, y = ops.sample_weights(seg, None, lm)
    w2, y2 = ops.sample_weights(seg, torch.zeros(2, 0, dtype=torch.int64, device=dev), lm)
    return float(w.sum()), float(w2.sum())
run("sample_weights none/empty", f)
print("failures:", bad)
