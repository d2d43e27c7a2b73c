"""Bring-up check of the TS-mode InfoNCE kernel (csrc/infonce_ts.cu) on a B200: the TS building-block GEMM, parity of
the fused kernel against the shared-memory-operand kernel and the fp64 oracle, and A/B timing at the headline size.
Writes gpurun_out/r2_ts_check.json.  Usage: python tools/r2_ts_check.py [stage ...]   (stages: gemm parity time)"""
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from rangeclip_b200 import _lib, ops  # noqa: E402

dev = torch.device("cuda")
L = _lib.lib()
st = torch.cuda.current_stream().cuda_stream
out = {}


def gemm_stage():
    res = []
    for N, Kd in ((128, 256), (256, 64), (64, 128), (128, 64)):
        g = torch.Generator(device=dev).manual_seed(N * 1000 + Kd)
        a = torch.randn(256, Kd, device=dev, generator=g).to(torch.bfloat16)
        b = torch.randn(N, Kd, device=dev, generator=g).to(torch.bfloat16)
        c = torch.empty(256, N, device=dev, dtype=torch.float32)
        _lib.check(_lib.bringup_lib().rc_debug_umma_gemm_ts_2sm(a.data_ptr(), b.data_ptr(), N, Kd, c.data_ptr(), st), "ts gemm")
        torch.cuda.synchronize()
        ref = a.float() @ b.float().t()
        err = float((c - ref).abs().max())
        res.append(dict(N=N, Kd=Kd, max_abs_err=err, ref_max=float(ref.abs().max())))
        print("ts gemm", res[-1], flush=True)
    out["gemm"] = res


def run_infonce(x, tb, ttb, K, y, w, inv_tau, flags, gs=None):
    B, D = x.shape[0], x.shape[1]
    HW = x[0, 0].numel()
    M = B * HW
    acc = torch.zeros(4, device=dev, dtype=torch.float64)
    lse = torch.empty(M, device=dev, dtype=torch.float32)
    dx = torch.empty(B, D, HW, device=dev, dtype=torch.bfloat16)
    ws_bytes = int(L.rc_infonce_workspace_bytes(B, D, HW, K, _lib.RC_BF16))
    ws = torch.empty(ws_bytes, device=dev, dtype=torch.uint8)
    _lib.check(L.rc_weight_sum(w.data_ptr(), y.data_ptr(), M, acc[3:].data_ptr(), st), "wsum")
    _lib.check(L.rc_infonce_bf16(x.data_ptr(), _lib.RC_BF16, B, D, HW, tb.data_ptr(), ttb.data_ptr(), K, y.data_ptr(), w.data_ptr(),
                                 float(inv_tau), lse.data_ptr(), acc[0:].data_ptr(), acc[1:].data_ptr(), acc[3:].data_ptr(),
                                 None if gs is None else gs.data_ptr(), dx.data_ptr(), None, acc[2:].data_ptr(), ws.data_ptr(),
                                 ws_bytes, flags, st), "rc_infonce_bf16")
    torch.cuda.synchronize()
    return dict(loss=float(acc[0] / acc[1]), dlogtau=float(acc[2]), lse=lse, dx=dx)


def make_case(B, D, HW, K, seed, frac_ignored=0.2):
    g = torch.Generator(device=dev).manual_seed(seed)
    x = torch.randn(B, D, HW, device=dev, generator=g)
    x = (x / x.norm(dim=1, keepdim=True) * (0.5 + torch.rand(B, 1, HW, device=dev, generator=g))).to(torch.bfloat16)
    text = torch.nn.functional.normalize(torch.randn(K, D, device=dev, generator=g), dim=1)
    _, tb, ttb = ops.text_prepare(text, None, want_f32=False, want_bf16=True)
    y = torch.randint(0, K, (B * HW,), device=dev, generator=g, dtype=torch.int32)
    w = torch.randint(0, 3, (B * HW,), device=dev, generator=g).float()
    ign = torch.rand(B * HW, device=dev, generator=g) < frac_ignored
    y = torch.where(ign, torch.full_like(y, -1), y)
    return x, text, tb, ttb, y, w


def parity_stage():
    sys.path.insert(0, ROOT)
    from oracle import rangeclip_oracle as O
    res = []
    for (B, D, HW, K, tau) in ((2, 512, 2048, 256, 0.07), (3, 256, 1000, 100, 0.07), (1, 512, 640, 64, 0.2), (2, 512, 4096, 200, 0.02)):
        x, text, tb, ttb, y, w = make_case(B, D, HW, K, 7 + B + K)
        r_ts = run_infonce(x, tb, ttb, K, y, w, 1.0 / tau, 8)
        r_ss = run_infonce(x, tb, ttb, K, y, w, 1.0 / tau, 0)
        rec = dict(B=B, D=D, HW=HW, K=K, tau=tau, loss_ts=r_ts["loss"], loss_ss=r_ss["loss"], dlt_ts=r_ts["dlogtau"], dlt_ss=r_ss["dlogtau"])
        d_ts, d_ss = r_ts["dx"].float(), r_ss["dx"].float()
        rec["dx_ts_vs_ss_maxrel"] = float((d_ts - d_ss).abs().max() / d_ss.abs().max())
        rec["lse_ts_vs_ss"] = float((r_ts["lse"] - r_ss["lse"]).abs().max())
        # fp64 oracle on the same bf16-rounded inputs
        try:
            rows = x.float().cpu().permute(0, 2, 1).reshape(B * HW, D)
            o = O.infonce_dense(rows, tb[:K].float().cpu(), y.cpu().long(), w.cpu(), 1.0 / tau)
            rec["loss_oracle"] = float(o["loss"])
            dxo = o["dx"].view(B, HW, D).permute(0, 2, 1).float()
            rec["dx_ts_vs_oracle_maxrel"] = float((d_ts.cpu() - dxo).abs().max() / dxo.abs().max())
            rec["dx_ss_vs_oracle_maxrel"] = float((d_ss.cpu() - dxo).abs().max() / dxo.abs().max())
            rec["dlt_oracle"] = float(o["dlogtau"])
        except Exception as e:  # the oracle's helper signature may differ; TS-vs-SS is the gate here
            rec["oracle_error"] = repr(e)[:200]
        res.append(rec)
        print("parity", rec, flush=True)
    out["parity"] = res


def time_stage():
    B, D, HW, K = 64, 512, 65536, 256
    g = torch.Generator(device=dev).manual_seed(1)
    x = torch.empty(B, D, HW, device=dev, dtype=torch.bfloat16)
    for b in range(B):
        xb = torch.randn(D, HW, device=dev, generator=g)
        x[b] = (xb / xb.norm(dim=0, keepdim=True)).to(torch.bfloat16)
    text = torch.nn.functional.normalize(torch.randn(K, D, device=dev, generator=g), dim=1)
    _, tb, ttb = ops.text_prepare(text, None, want_f32=False, want_bf16=True)
    y = torch.randint(0, K, (B * HW,), device=dev, generator=g, dtype=torch.int32)
    w = torch.randint(0, 3, (B * HW,), device=dev, generator=g).float()
    M = B * HW
    acc = torch.zeros(4, device=dev, dtype=torch.float64)
    lse = torch.empty(M, device=dev, dtype=torch.float32)
    dx = torch.empty(B, D, HW, device=dev, dtype=torch.bfloat16)
    ws_bytes = int(L.rc_infonce_workspace_bytes(B, D, HW, K, _lib.RC_BF16))
    ws = torch.empty(ws_bytes, device=dev, dtype=torch.uint8)
    _lib.check(L.rc_weight_sum(w.data_ptr(), y.data_ptr(), M, acc[3:].data_ptr(), st), "wsum")
    res = {}
    for name, flags in (("ts", 8), ("ss", 0), ("ts2", 8), ("ss2", 0)):
        def launch():
            _lib.check(L.rc_infonce_bf16(x.data_ptr(), _lib.RC_BF16, B, D, HW, tb.data_ptr(), ttb.data_ptr(), K, y.data_ptr(), w.data_ptr(),
                                         1.0 / 0.07, lse.data_ptr(), acc[0:].data_ptr(), acc[1:].data_ptr(), acc[3:].data_ptr(), None,
                                         dx.data_ptr(), None, acc[2:].data_ptr(), ws.data_ptr(), ws_bytes, flags, st), "rc_infonce_bf16")
        for _ in range(3):
            launch()
        torch.cuda.synchronize()
        ts = []
        for _ in range(10):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(); launch(); e1.record()
            torch.cuda.synchronize()
            ts.append(e0.elapsed_time(e1))
        ts.sort()
        clk = None
        try:
            import pynvml as N
            N.nvmlInit()
            clk = N.nvmlDeviceGetClockInfo(N.nvmlDeviceGetHandleByIndex(0), N.NVML_CLOCK_SM)
        except Exception:
            pass
        res[name] = dict(ms_min=ts[0], ms_med=ts[len(ts) // 2], tflops_med=4.0 * M * K * D / (ts[len(ts) // 2] * 1e-3) / 1e12,
                         sm_mhz_after=clk, lib=os.path.basename(_lib.LIB_PATH), spb=os.environ.get("RANGECLIP_B200_TS_SPB"))
        print("time", name, res[name], flush=True)
    out["time"] = res


if __name__ == "__main__":
    stages = sys.argv[1:] or ["gemm", "parity", "time"]
    try:
        for s_ in stages:
            {"gemm": gemm_stage, "parity": parity_stage, "time": time_stage}[s_]()
    finally:
        os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
        tag = os.environ.get("RC_TAG", "")
        with open(os.path.join(ROOT, "gpurun_out", "r2_ts_check_%s%s.json" % ("_".join(stages), tag)), "w") as f:
            json.dump(out, f, indent=1, default=str)
