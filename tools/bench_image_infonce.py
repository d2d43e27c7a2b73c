import sys, torch, time
sys.path.insert(0, ".")
import rangeclip_b200 as R
dev = torch.device("cuda:0")
g = torch.Generator(device=dev).manual_seed(0)
for n in (64, 1024, 4096):
    area = torch.randn(n, 512, device=dev, generator=g, requires_grad=True)
    img = torch.nn.functional.normalize(torch.randn(n, 512, device=dev, generator=g), dim=1)
    lt = torch.log(torch.tensor(0.1, device=dev)).requires_grad_(True)
    def step():
        area.grad = None
        l = R.image_contrastive_loss(area, img, lt)
        l.backward()
        return l
    for _ in range(2): step()
    torch.cuda.synchronize(); t0 = time.perf_counter()
    for _ in range(5): l = step()
    torch.cuda.synchronize(); ms = (time.perf_counter() - t0) / 5 * 1e3
    # eager PyTorch on the same GPU (the reference's op sequence, model.py:304-321)
    def eager():
        area.grad = None
        a = torch.nn.functional.normalize(area, dim=1); logits = a @ img.t() / torch.exp(lt)
        l = torch.nn.functional.cross_entropy(logits, torch.arange(n, device=dev)); l.backward(); return l
    for _ in range(2): eager()
    torch.cuda.synchronize(); t0 = time.perf_counter()
    for _ in range(5): le = eager()
    torch.cuda.synchronize(); ms_e = (time.perf_counter() - t0) / 5 * 1e3
    print(f"n={n}: ours {ms:.3f} ms (loss {float(l):.5f}), eager fp32 {ms_e:.3f} ms (loss {float(le):.5f})")
