"""Same-box A/B of library variants on the evaluation kernel (one batch of 64 realistic maps, K=1024, top-5 and top-1, ids
written): make -C rangeclip_b200/csrc variant NAME=x VSRC=eval_topk_umma DEFS=..;  python tools/ab_topk.py base x ..."""
import json, os, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if len(sys.argv) > 1 and sys.argv[1] == "child":
    sys.path.insert(0, ROOT)
    os.environ["RANGECLIP_B200_LIB"] = os.environ["RC_AB_LIB"]
    import torch
    from rangeclip_b200 import ops
    from bench import make_eval_batch
    dev = torch.device("cuda:0")
    B, D, H, W, C = 64, 512, 256, 256, 1024
    g = torch.Generator(device=dev).manual_seed(99)
    text = torch.nn.functional.normalize(torch.randn(C, D, device=dev, generator=g), dim=1)
    _, tb, _ = ops.text_prepare(text, None, want_f32=False, want_bf16=True)
    im = torch.arange(C, device=dev)
    x, _ = make_eval_batch(dev, text, B, H, W, 0.3, 500)
    out = {"lib": os.path.basename(os.environ["RC_AB_LIB"])}
    for k in (5, 1):
        for _ in range(2): ops.eval_topk(x, text, im, k, "bf16", t_bf16=tb)
        torch.cuda.synchronize()
        ts = []
        for _ in range(7):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(); ops.eval_topk(x, text, im, k, "bf16", t_bf16=tb); e1.record(); torch.cuda.synchronize(); ts.append(e0.elapsed_time(e1))
        out[f"top{k}_ms_min"] = round(min(ts), 4); out[f"top{k}_ms_med"] = round(sorted(ts)[3], 4)
    print(json.dumps(out))
else:
    for rep in range(2):
        for n in sys.argv[1:]:
            lib = os.path.join(ROOT, "rangeclip_b200", "librangeclip_b200_bringup.so" if n == "base" else f"librangeclip_b200_var_{n}.so")
            r = subprocess.run([sys.executable, __file__, "child"], env=dict(os.environ, RC_AB_LIB=lib), capture_output=True, text=True, timeout=300)
            print(r.stdout.strip() or r.stderr[-600:], flush=True)
