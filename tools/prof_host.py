"""Host-side cost of the sync-free compute_loss (contrast_builder='device') + backward: cProfile over a loop without any
device synchronisation inside (so wall time per call == host enqueue time once the GPU queue is deeper than the host).
python tools/prof_host.py [B]"""
import cProfile, os, pstats, sys, time
import numpy as np
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import rangeclip_b200 as R

B = int(sys.argv[1]) if len(sys.argv) > 1 else 16
dev = torch.device("cuda:0")
D, H, W, C, K = 512, 256, 256, 1024, 256
g = torch.Generator(device=dev).manual_seed(0)
x = torch.randn(B, D, H, W, device=dev, generator=g).to(torch.bfloat16).requires_grad_(True)
text = torch.randn(C, D, device=dev, generator=g)
seg = torch.randint(0, 64, (B, H // 32, W // 32), device=dev, generator=g).repeat_interleave(32, 1).repeat_interleave(32, 2)
rng = np.random.default_rng(0)
sets = {"hard": {c: [int(v) for v in rng.choice(C, 50, replace=False)] for c in range(C)}, "medium": {}}


class M(torch.nn.Module):
    def __init__(self):
        super().__init__()
        self.log_temperature_text = torch.nn.Parameter(torch.log(torch.tensor(0.07)))
        self.log_temperature_image = torch.nn.Parameter(torch.log(torch.tensor(0.1)))


model = M().to(dev)
params = [x] + list(model.parameters())


def step(builder):
    total, info = R.compute_loss(model, x, seg, text, sets, None, None, W_text=1.0, W_image=0.0, W_smooth=0.0, k_distractors=193,
                                 pct_medium=0.0, pct_hard=1.0, pct_rand=0.0, contrast_builder=builder)
    torch.autograd.grad(total, params, allow_unused=True)
    return info


for builder in ("device", "reference"):
    for _ in range(3):
        step(builder)
    torch.cuda.synchronize()
    n = int(os.environ.get('PROF_STEPS', 30))
    t0 = time.perf_counter()
    for _ in range(n):
        step(builder)
    t1 = time.perf_counter()
    torch.cuda.synchronize()
    t2 = time.perf_counter()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10):
        step(builder)
    e1.record(); torch.cuda.synchronize()
    print(f"builder={builder}: device-timed {e0.elapsed_time(e1) / 10:.3f} ms/step")
    print(f"builder={builder}: host enqueue {1e3 * (t1 - t0) / n:.3f} ms/step, with final drain {1e3 * (t2 - t0) / n:.3f} ms/step (B={B})")
pr = cProfile.Profile()
pr.enable()
for _ in range(int(os.environ.get('PROF_STEPS', 30))):
    step("device")
pr.disable()
torch.cuda.synchronize()
st = pstats.Stats(pr)
st.sort_stats("cumulative").print_stats(45)
