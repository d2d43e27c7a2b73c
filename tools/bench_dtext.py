"""dText path at the headline size: fused launch with G store + the dT = G^T X kernel, CTA-pair form against the single-CTA form
(bring-up build, RANGECLIP_B200_DT=1cta).  python tools/bench_dtext.py"""
import json, os, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if len(sys.argv) > 1 and sys.argv[1] == "child":
    sys.path.insert(0, ROOT)
    os.environ["RANGECLIP_B200_LIB"] = os.path.join(ROOT, "rangeclip_b200", "librangeclip_b200_bringup.so")
    import torch
    from rangeclip_b200 import _lib, ops
    dev = torch.device("cuda"); L = _lib.lib(); st = torch.cuda.current_stream().cuda_stream
    B, D, HW, K = 64, 512, 65536, 256
    g = torch.Generator(device=dev).manual_seed(1)
    x = torch.empty(B, D, HW, device=dev, dtype=torch.bfloat16)
    for b in range(B):
        xb = torch.randn(D, HW, device=dev, generator=g); x[b] = (xb / xb.norm(dim=0, keepdim=True)).to(torch.bfloat16)
    text = torch.nn.functional.normalize(torch.randn(K, D, device=dev, generator=g), dim=1)
    _, tb, ttb = ops.text_prepare(text, None, want_f32=False, want_bf16=True)
    y = torch.randint(0, K, (B * HW,), device=dev, generator=g, dtype=torch.int32)
    w = torch.randint(0, 3, (B * HW,), device=dev, generator=g).float()
    M = B * HW
    acc = torch.zeros(4, device=dev, dtype=torch.float64); lse = torch.empty(M, device=dev)
    dx = torch.empty(B, D, HW, device=dev, dtype=torch.bfloat16); dt = torch.zeros(K, D, device=dev)
    wsb = int(L.rc_infonce_workspace_bytes_dt(B, D, HW, K, _lib.RC_BF16)); ws = torch.empty(wsb, device=dev, dtype=torch.uint8)
    _lib.check(L.rc_weight_sum(w.data_ptr(), y.data_ptr(), M, acc[3:].data_ptr(), st), "wsum")
    def run(with_dt):
        _lib.check(L.rc_infonce_bf16(x.data_ptr(), _lib.RC_BF16, B, D, HW, tb.data_ptr(), ttb.data_ptr(), K, y.data_ptr(), w.data_ptr(), 1 / 0.07,
                                     lse.data_ptr(), acc[0:].data_ptr(), acc[1:].data_ptr(), acc[3:].data_ptr(), None, dx.data_ptr(),
                                     dt.data_ptr() if with_dt else None, acc[2:].data_ptr(), ws.data_ptr(), wsb, 0, st), "infonce")
    out = {"dt": os.environ.get("RANGECLIP_B200_DT", "pair")}
    for with_dt in (False, True):
        for _ in range(3): run(with_dt)
        torch.cuda.synchronize(); ts = []
        for _ in range(7):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(); run(with_dt); e1.record(); torch.cuda.synchronize(); ts.append(e0.elapsed_time(e1))
        out["with_dt_ms" if with_dt else "no_dt_ms"] = sorted(ts)[3]
    out["dt_sum"] = float(dt.abs().sum())
    print(json.dumps(out))
else:
    for rep in range(2):
        for v in ("pair", "1cta"):
            r = subprocess.run([sys.executable, __file__, "child"], env=dict(os.environ, RANGECLIP_B200_DT=v), capture_output=True, text=True, timeout=300)
            print(r.stdout.strip() or r.stderr[-400:], flush=True)
