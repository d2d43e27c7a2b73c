"""Thin PyTorch-facing wrappers over the C ABI (``include/rangeclip_b200.h``).

PyTorch is used for device memory, streams and autograd plumbing only; every computation is a
kernel in ``librangeclip_b200.so``.  All functions require CUDA tensors on an sm_100 device and
raise ``RuntimeError`` otherwise -- there is no CPU or eager-PyTorch fallback.
"""
from __future__ import annotations

from typing import Optional, Tuple

import torch

from . import _lib
from ._lib import RC_BF16, RC_F32, check

_I32_MAX = 2**31 - 1


def _need_cuda(*tensors: torch.Tensor) -> None:
    for t in tensors:
        if t is not None and not t.is_cuda:
            raise RuntimeError("rangeclip_b200 ops need CUDA tensors (no CPU fallback exists)")


def _p(t: Optional[torch.Tensor]):
    return None if t is None else t.data_ptr()


def _stream(t: torch.Tensor):
    return torch.cuda.current_stream(t.device).cuda_stream


def _dt(t: torch.Tensor) -> int:
    if t.dtype == torch.float32:
        return RC_F32
    if t.dtype == torch.bfloat16:
        return RC_BF16
    raise RuntimeError(f"rangeclip_b200: unsupported embedding dtype {t.dtype} (float32 or bfloat16)")


def _emb3(x: torch.Tensor) -> Tuple[torch.Tensor, int, int, int]:
    """[B, D, H, W] (or [B, D, HW]) contiguous view -> (x, B, D, HW)."""
    if x.dim() not in (3, 4):
        raise RuntimeError("pixel embeddings must be [B, D, H, W]")
    x = x.contiguous()
    B, D = x.shape[0], x.shape[1]
    return x, B, D, x[0, 0].numel() if B and D else 0


def bf16_path_supported(D: int, HW: int, K: int) -> bool:
    """Shapes the tcgen05 InfoNCE kernel covers (everything else runs on the fp32 kernel)."""
    return 1 <= K <= 256 and D in (128, 256, 384, 512) and HW % 8 == 0 and HW > 0


# ----------------------------------------------------------------------------------------------
# text preparation / sampling weights
# ----------------------------------------------------------------------------------------------

def text_prepare(text: torch.Tensor, idx: Optional[torch.Tensor], want_f32=True, want_bf16=False):
    """F.normalize(text[idx], dim=1) (model.py:272) as f32 [K,D] and/or bf16 [Kp,D] + [D,Kp]."""
    _need_cuda(text, idx)
    text = text.float().contiguous()
    K = int(idx.numel()) if idx is not None else text.shape[0]
    D = text.shape[1]
    Kp = (K + 63) // 64 * 64
    t32 = torch.empty(K, D, device=text.device, dtype=torch.float32) if want_f32 else None
    tb = torch.empty(Kp, D, device=text.device, dtype=torch.bfloat16) if want_bf16 else None
    ttb = torch.empty(D, Kp, device=text.device, dtype=torch.bfloat16) if want_bf16 else None
    if idx is not None:
        idx = idx.to(torch.int64).contiguous()
    check(_lib.lib().rc_text_prepare(_p(text), text.stride(0), _p(idx), K, D, _p(t32), _p(tb), _p(ttb),
                                     _stream(text)), "rc_text_prepare")
    return t32, tb, ttb


def text_to_bf16(t_norm: torch.Tensor):
    """bf16 [Kp,D] and transposed [D,Kp] copies of already-normalised rows (tiny; plumbing only)."""
    K, D = t_norm.shape
    Kp = (K + 63) // 64 * 64
    tb = torch.zeros(Kp, D, device=t_norm.device, dtype=torch.bfloat16)
    tb[:K] = t_norm.detach().to(torch.bfloat16)
    return tb, tb.t().contiguous()


def sample_weights(seg: torch.Tensor, rand_idx: Optional[torch.Tensor], label_map: torch.Tensor):
    """Dense form of model.py:220-228,276-284: w[b,p] = multiplicity of p among rand_idx[b] (1 when
    rand_idx is None) zeroed on background / unmapped labels; y[b,p] = label_map[seg] or -1."""
    _need_cuda(seg, rand_idx, label_map)
    B = seg.shape[0]
    HW = seg[0].numel() if B else 0
    seg = seg.to(torch.int64).contiguous()
    label_map = label_map.to(torch.int32).contiguous()
    w = torch.empty(B, HW, device=seg.device, dtype=torch.float32)
    y = torch.empty(B, HW, device=seg.device, dtype=torch.int32)
    n_s = 0
    if rand_idx is not None:
        rand_idx = rand_idx.to(torch.int64).contiguous()
        n_s = rand_idx.shape[1]
    check(_lib.lib().rc_sample_weights(_p(seg), _p(rand_idx), B, HW, n_s, _p(label_map), label_map.numel(),
                                       _p(w), _p(y), _stream(seg)), "rc_sample_weights")
    return w, y


# ----------------------------------------------------------------------------------------------
# InfoNCE
# ----------------------------------------------------------------------------------------------

def infonce_raw(x: torch.Tensor, t_norm: torch.Tensor, y: torch.Tensor, w: torch.Tensor, inv_tau: float,
                need_dx: bool, need_dt: bool, precision: str = "auto",
                grad_scale: Optional[torch.Tensor] = None, t_bf16=None, rep: int = 1):
    """One fused pass: returns dict(loss_sum, w_sum (double[1] tensors), lse, dx, dt, dlogtau).
    loss = loss_sum / w_sum; dx/dt/dlogtau are gradients of that mean loss times grad_scale.
    rep = 4: every row of x is the embedding shared by a 2x2 block of pixels (decoder.py:113, Q8);
    y / w are [rows, 4] and dx is the gradient w.r.t. the shared row (tensor-core path only)."""
    _need_cuda(x, t_norm, y, w)
    x, B, D, HW = _emb3(x)
    K = t_norm.shape[0]
    dev = x.device
    M = B * HW
    y = y.reshape(-1).to(torch.int32).contiguous()
    w = w.reshape(-1).to(torch.float32).contiguous()
    if rep not in (1, 4):
        raise RuntimeError("infonce: rep must be 1 or 4")
    if y.numel() != M * rep or w.numel() != M * rep:
        raise RuntimeError("infonce: y / w must have `rep` entries per embedding row")
    if rep == 4:
        if precision == "fp32" or need_dt and not need_dx or D not in (256, 512) or not bf16_path_supported(D, HW, K):
            raise RuntimeError(f"infonce(rep=4): needs the tensor-core path (D in (256, 512), K <= 256, HW % 8 == 0); "
                               f"got D={D}, HW={HW}, K={K}, precision={precision!r}")
        precision = "bf16"
    if precision == "auto" and rep == 1 and K > 256 and not need_dt and kblocked_supported(D, HW):
        # more candidates than one launch takes: tensor cores over blocks of 256 candidate rows instead of the CUDA-core
        # kernel (which needs ~100x longer at full size)
        r = infonce_kblocked_raw(x, t_norm, y, w, inv_tau, need_dx)
        dx = r["dx"]
        if dx is not None and grad_scale is not None:
            dx = dx * grad_scale.detach().reshape(1).to(device=dev, dtype=dx.dtype)
        dlt = r["dlogtau"] if r["dlogtau"] is not None else torch.zeros((), device=dev, dtype=torch.float64)
        if grad_scale is not None:
            dlt = dlt * grad_scale.detach().reshape(()).to(device=dev, dtype=torch.float64)
        return dict(loss_sum=r["loss"] * r["w_sum"], w_sum=r["w_sum"], dlogtau=dlt, lse=r["lse"], dx=dx, dt=None,
                    precision="bf16-kblocked")
    dt_on_tc = need_dt and need_dx and D in (256, 512)       # tensor-core dText: pair kernel + split-K GEMM
    if precision == "auto":
        precision = "bf16" if (bf16_path_supported(D, HW, K) and (not need_dt or dt_on_tc)) else "fp32"
    acc = torch.zeros(4, device=dev, dtype=torch.float64)        # loss_sum, w_sum, dlogtau, w_sum_in
    lse = torch.empty(M, device=dev, dtype=torch.float32)
    need_grad = need_dx or need_dt
    L = _lib.lib()
    st = _stream(x)
    if need_grad:
        check(L.rc_weight_sum(_p(w), _p(y), M * rep, acc[3:].data_ptr(), st), "rc_weight_sum")
    gs = None
    if grad_scale is not None:
        gs = grad_scale.detach().reshape(1).to(device=dev, dtype=torch.float32)
    dt = torch.zeros(K, D, device=dev, dtype=torch.float32) if need_dt else None
    if precision == "fp32":
        xf = x if x.dtype == torch.float32 else x.float()
        tf = t_norm.detach().float().contiguous()
        dx = torch.empty_like(xf) if need_dx else None
        check(L.rc_infonce_f32(_p(xf), B, D, HW, D * HW, _p(tf), K, _p(y), _p(w), float(inv_tau), _p(lse),
                               acc[0:].data_ptr(), acc[1:].data_ptr(),
                               acc[3:].data_ptr() if need_grad else None, _p(gs),
                               _p(dx), _p(dt), acc[2:].data_ptr() if need_grad else None, st), "rc_infonce_f32")
        if dx is not None and dx.dtype != x.dtype:
            dx = dx.to(x.dtype)
    elif precision == "bf16":
        if not bf16_path_supported(D, HW, K):
            raise RuntimeError(f"infonce: bf16 tensor-core path does not cover D={D}, HW={HW}, K={K}")
        if need_dt and not dt_on_tc:
            raise RuntimeError("infonce: dText on the bf16 path needs dx and D in (256, 512); use precision='fp32'")
        if t_bf16 is None:
            t_bf16 = text_to_bf16(t_norm)
        tb, ttb = t_bf16
        xdt = _dt(x)
        ws_bytes = int((L.rc_infonce_workspace_bytes_dt if need_dt else L.rc_infonce_workspace_bytes)(B, D, HW, K, xdt))
        ws = torch.empty(ws_bytes, device=dev, dtype=torch.uint8)
        dxb = torch.empty(B, D, HW, device=dev, dtype=torch.bfloat16) if need_dx else None
        entry = L.rc_infonce_bf16 if rep == 1 else L.rc_infonce_bf16_rep4
        check(entry(_p(x), xdt, B, D, HW, _p(tb), _p(ttb), K, _p(y), _p(w), float(inv_tau), _p(lse),
                    acc[0:].data_ptr(), acc[1:].data_ptr(),
                    acc[3:].data_ptr() if need_grad else None, _p(gs), _p(dxb), _p(dt),
                    acc[2:].data_ptr() if need_dx else None, _p(ws), ws_bytes, 0, st),
              "rc_infonce_bf16" if rep == 1 else "rc_infonce_bf16_rep4")
        dx = None
        if dxb is not None:
            dx = dxb.view(x.shape) if x.dtype == torch.bfloat16 else dxb.view(x.shape).to(x.dtype)
    else:
        raise RuntimeError(f"infonce: unknown precision {precision!r}")
    return dict(loss_sum=acc[0], w_sum=acc[1], dlogtau=acc[2], lse=lse, dx=dx, dt=dt, precision=precision)


class _InfoNCE(torch.autograd.Function):
    """loss = weighted InfoNCE(x, t_norm, y, w) / tau; gradients for x, t_norm and log_tau are
    produced by the same fused kernel launch as the loss (single pass over X)."""

    @staticmethod
    def forward(ctx, x, t_norm, log_tau, y, w, precision, rep=1):
        need_dx = x.requires_grad
        need_dt = t_norm.requires_grad
        need_tau = log_tau.requires_grad
        inv_tau = float(torch.exp(-log_tau.detach().float()))          # one scalar sync, as .item() in the reference
        r = infonce_raw(x.detach(), t_norm.detach(), y, w, inv_tau, need_dx or need_tau, need_dt, precision, rep=rep)
        wsum = r["w_sum"]
        loss = torch.where(wsum > 0, r["loss_sum"] / wsum.clamp_min(1e-300), torch.zeros_like(wsum)).float()
        ctx.save_for_backward(r["dx"] if need_dx else None, r["dt"], r["dlogtau"].float())
        ctx.flags = (need_dx, need_dt, need_tau)
        ctx.x_dtype = x.dtype
        return loss

    @staticmethod
    def backward(ctx, g):
        dx, dt, dlt = ctx.saved_tensors
        need_dx, need_dt, need_tau = ctx.flags
        gx = gt = gl = None
        if need_dx:
            gx = dx
            gdev = g.detach().reshape(1).float()
            check(_lib.lib().rc_scale(_p(gx), _dt(gx), gx.numel(), _p(gdev), _stream(gx)), "rc_scale")
        if need_dt:
            gt = dt * g
        if need_tau:
            gl = (dlt * g).reshape(())
        return gx, gt, gl, None, None, None, None


class _PixelLosses(torch.autograd.Function):
    """Text InfoNCE + smoothness of the SAME pixel embeddings as one autograd node, so that the backward is a
    single pass: dX = g_text * dX_text (computed in the forward launch) + g_smooth * d(TV)/dX, written in
    place by rc_tv_bwd(accumulate, dx_scale) -- X is read once, dX read and written once (model.py:272-291,
    332-334 and their autograd)."""

    @staticmethod
    def forward(ctx, x, t_norm, log_tau, y, w, precision):
        need_dx = x.requires_grad
        need_dt = t_norm.requires_grad
        need_tau = log_tau.requires_grad
        inv_tau = float(torch.exp(-log_tau.detach().float()))
        r = infonce_raw(x.detach(), t_norm.detach(), y, w, inv_tau, need_dx or need_tau, need_dt, precision)
        wsum = r["w_sum"]
        text = torch.where(wsum > 0, r["loss_sum"] / wsum.clamp_min(1e-300), torch.zeros_like(wsum)).float()
        sums = tv_sums(x.detach())
        dh, dv = tv_denominators(x.shape)
        nan = torch.full((), float("nan"), device=x.device, dtype=torch.float64)
        smooth = ((sums[0] / dh if dh > 0 else nan) + (sums[1] / dv if dv > 0 else nan)).float()
        ctx.save_for_backward(x, r["dx"] if need_dx else None, r["dt"], r["dlogtau"].float())
        ctx.flags = (need_dx, need_dt, need_tau)
        return text, smooth

    @staticmethod
    def backward(ctx, g_text, g_smooth):
        x, dx, dt, dlt = ctx.saved_tensors
        need_dx, need_dt, need_tau = ctx.flags
        gx = gt = gl = None
        if need_dx:
            dh, dv = tv_denominators(x.shape)
            gs = g_smooth.float()
            scale = torch.stack([gs / dh if dh > 0 else gs * 0, gs / dv if dv > 0 else gs * 0])
            gx = tv_backward(x.detach(), scale, dx=dx, dx_scale=g_text.float())
        if need_dt:
            gt = dt * g_text
        if need_tau:
            gl = (dlt * g_text).reshape(())
        return gx, gt, gl, None, None, None


RC_INFONCE_KEEP_WEIGHT, RC_INFONCE_LSE_GIVEN = 2, 4


def kblocked_supported(D: int, HW: int) -> bool:
    return D in (256, 512) and HW % 8 == 0 and HW > 0


def infonce_kblocked_raw(x: torch.Tensor, t_norm: torch.Tensor, y: torch.Tensor, w: torch.Tensor, inv_tau: float,
                         need_dx: bool, block: int = 256):
    """InfoNCE against MORE than 256 candidates on the tensor cores (model.py:304-321 at thousands of objects): the
    candidate rows are split into launches of <= 256.  Round 1: forward launches give the per-block logsumexp; their
    logsumexp is the row's lse over all candidates.  Round 2: fwd+bwd launches with that lse given produce per-block dx /
    dlogtau that add up to the full gradient.  Returns dict(loss, lse, dx, dlogtau, w_sum); bf16 operands, fp32 sums."""
    _need_cuda(x, t_norm, y, w)
    x, B, D, HW = _emb3(x)
    if not kblocked_supported(D, HW):
        raise RuntimeError(f"infonce_kblocked: needs D in (256, 512) and HW % 8 == 0; got D={D}, HW={HW}")
    K = t_norm.shape[0]
    M = B * HW
    dev = x.device
    y = y.reshape(-1).to(torch.int32)
    w = (w.reshape(-1).to(torch.float32) * (y >= 0)).contiguous()        # ignored rows: weight 0 (targets may leave the block)
    L = _lib.lib()
    st = _stream(x)
    xdt = _dt(x)
    ws_bytes = int(L.rc_infonce_workspace_bytes(B, D, HW, block, xdt))
    ws = torch.empty(ws_bytes, device=dev, dtype=torch.uint8)
    check(L.rc_infonce_prepass(_p(x), xdt, B, D, HW, _p(ws), ws_bytes, st), "rc_infonce_prepass")    # once for all launches
    starts = list(range(0, K, block))
    nb = len(starts)
    if B == 1 and HW % 256 == 0 and block == 256:
        # one image of rows (the area-image loss): every candidate block in ONE launch per round -- the kernel's image
        # index runs over the blocks (rc_infonce_bf16_kblocks)
        tb_all = torch.zeros(nb * 256, D, device=dev, dtype=torch.bfloat16)
        tb_all[:K] = t_norm.detach().to(torch.bfloat16)
        ttb_all = tb_all.t().contiguous()
        y_rel = (y[None, :] - 256 * torch.arange(nb, device=dev, dtype=torch.int32)[:, None]).contiguous()
        w_rep = w[None, :].expand(nb, M).contiguous()
        wsum = w.double().sum()
        acc = torch.zeros(4, device=dev, dtype=torch.float64)
        acc[3] = wsum * nb                       # (unused by the forward round)
        lse_blk = torch.empty(nb, M, device=dev, dtype=torch.float32)
        check(L.rc_infonce_bf16_kblocks(_p(x), xdt, D, HW, _p(tb_all), _p(ttb_all), K, nb, _p(y_rel), _p(w_rep), float(inv_tau),
                                        _p(lse_blk), acc[0:].data_ptr(), acc[1:].data_ptr(), None, None, None, None, _p(ws),
                                        ws_bytes, 1, st), "rc_infonce_bf16_kblocks(forward)")
        lse = torch.logsumexp(lse_blk, dim=0)
        wz = (lse_blk.double() * w.double()[None, :]).sum() - acc[0]
        loss = torch.where(wsum > 0, ((lse.double() * w.double()).sum() - wz) / wsum.clamp_min(1e-300), torch.zeros_like(wsum))
        dx = dlogtau = None
        if need_dx:
            dxb = torch.empty(nb, D, HW, device=dev, dtype=torch.bfloat16)
            acc2 = torch.zeros(4, device=dev, dtype=torch.float64)
            acc2[3] = wsum
            check(L.rc_infonce_bf16_kblocks(_p(x), xdt, D, HW, _p(tb_all), _p(ttb_all), K, nb, _p(y_rel), _p(w_rep), float(inv_tau),
                                            _p(lse), acc2[0:].data_ptr(), acc2[1:].data_ptr(), acc2[3:].data_ptr(), None, _p(dxb),
                                            acc2[2:].data_ptr(), _p(ws), ws_bytes, 1 | RC_INFONCE_LSE_GIVEN, st),
                  "rc_infonce_bf16_kblocks(backward)")
            dx = dxb.float().sum(dim=0).view(x.shape).to(x.dtype)
            dlogtau = acc2[2].clone()
        return dict(loss=loss, lse=lse, dx=dx, dlogtau=dlogtau, w_sum=wsum)
    # bf16 operand copies of the candidate blocks: all full blocks with two tensor ops, a shorter last block on its own
    n_full = K // block if block % 64 == 0 else 0
    texts = []
    if n_full:
        tb_full = t_norm[:n_full * block].detach().to(torch.bfloat16).contiguous()
        ttb_full = tb_full.view(n_full, block, D).transpose(1, 2).contiguous()
        texts = [(tb_full[i * block:(i + 1) * block], ttb_full[i]) for i in range(n_full)]
    texts += [text_to_bf16(t_norm[s0:s0 + block]) for s0 in starts[n_full:]]
    ys = [(y - s0).contiguous() for s0 in starts]
    lse_blk = torch.empty(nb, M, device=dev, dtype=torch.float32)
    acc = torch.zeros(nb, 4, device=dev, dtype=torch.float64)             # per block: loss_sum, w_sum, dlogtau, w_sum_in
    wsum = w.double().sum()
    acc[:, 3] = wsum
    for i, s0 in enumerate(starts):
        Kb = min(block, K - s0)
        check(L.rc_infonce_bf16(_p(x), xdt, B, D, HW, _p(texts[i][0]), _p(texts[i][1]), Kb, _p(ys[i]), _p(w), float(inv_tau),
                                lse_blk[i].data_ptr(), acc[i, 0:].data_ptr(), acc[i, 1:].data_ptr(), None, None, None, None, None,
                                _p(ws), ws_bytes, 1 | RC_INFONCE_KEEP_WEIGHT, st), "rc_infonce_bf16(K block, forward)")
    lse = torch.logsumexp(lse_blk, dim=0)
    # sum_p w z[p, y_p] from the blocks that own the targets: loss_sum_b = sum_p w lse_b - sum_{y in b} w z_y
    wz = ((lse_blk.double() * w.double()).sum(dim=1) - acc[:, 0]).sum()
    loss = torch.where(wsum > 0, ((lse.double() * w.double()).sum() - wz) / wsum.clamp_min(1e-300), torch.zeros_like(wsum))
    dx = dlogtau = None
    if need_dx:
        dxb = torch.empty(B, D, HW, device=dev, dtype=torch.bfloat16)     # the tensor-core kernel always writes bf16
        dx = None                                                          # fp32 sum of the per-block gradients
        acc2 = torch.zeros(nb, 4, device=dev, dtype=torch.float64)
        acc2[:, 3] = wsum
        for i, s0 in enumerate(starts):
            Kb = min(block, K - s0)
            check(L.rc_infonce_bf16(_p(x), xdt, B, D, HW, _p(texts[i][0]), _p(texts[i][1]), Kb, _p(ys[i]), _p(w), float(inv_tau),
                                    _p(lse), acc2[i, 0:].data_ptr(), acc2[i, 1:].data_ptr(), acc2[i, 3:].data_ptr(), None, _p(dxb),
                                    None, acc2[i, 2:].data_ptr(), _p(ws), ws_bytes,
                                    1 | RC_INFONCE_KEEP_WEIGHT | RC_INFONCE_LSE_GIVEN, st), "rc_infonce_bf16(K block, backward)")
            dx = dxb.float() if dx is None else dx.add_(dxb)
        dlogtau = acc2[:, 2].sum()
        dx = dx.view(x.shape).to(x.dtype)
    return dict(loss=loss, lse=lse, dx=dx, dlogtau=dlogtau, w_sum=wsum)


class _InfoNCEKBlocked(torch.autograd.Function):
    """Autograd wrapper of ``infonce_kblocked_raw`` (gradients for x and log_tau; the candidate rows are constants)."""

    @staticmethod
    def forward(ctx, x, t_norm, log_tau, y, w):
        need = x.requires_grad or log_tau.requires_grad
        inv_tau = float(torch.exp(-log_tau.detach().float()))
        r = infonce_kblocked_raw(x.detach(), t_norm.detach(), y, w, inv_tau, need)
        ctx.save_for_backward(r["dx"], r["dlogtau"].float() if need else None)
        ctx.flags = (x.requires_grad, log_tau.requires_grad)
        return r["loss"].float()

    @staticmethod
    def backward(ctx, g):
        dx, dlt = ctx.saved_tensors
        need_dx, need_tau = ctx.flags
        return (dx * g.to(dx.dtype)) if need_dx else None, None, (dlt * g).reshape(()) if need_tau else None, None, None


def infonce_kblocked(x, t_norm, log_tau, y, w):
    return _InfoNCEKBlocked.apply(x, t_norm, log_tau, y, w)


def pixel_losses(x, t_norm, log_tau, y, w, precision="auto"):
    """(text InfoNCE, smoothness) with a fused single-pass backward."""
    return _PixelLosses.apply(x, t_norm, log_tau, y, w, precision)


def infonce(x, t_norm, log_tau, y, w, precision="auto", rep=1):
    """Autograd-aware fused InfoNCE; x [B,D,H,W], t_norm [K,D] normalised, y/w per pixel
    (rep = 4: x holds the embeddings shared by 2x2 pixel blocks, y/w are [B*H*W, 4])."""
    return _InfoNCE.apply(x, t_norm, log_tau, y, w, precision, rep)


# ----------------------------------------------------------------------------------------------
# smoothness (TV-L1)
# ----------------------------------------------------------------------------------------------

def tv_sums(x: torch.Tensor) -> torch.Tensor:
    """double[2]: sum |dx_w|, sum |dx_h| over [B,D,H,W] (model.py:332-333 numerators)."""
    _need_cuda(x)
    x = x.contiguous()
    B, D, H, W = x.shape
    sums = torch.zeros(2, device=x.device, dtype=torch.float64)
    check(_lib.lib().rc_tv_fwd(_p(x), _dt(x), B * D, H, W, _p(sums), _stream(x)), "rc_tv_fwd")
    return sums


def tv_denominators(shape) -> Tuple[float, float]:
    B, D, H, W = shape
    return float(B * D * H * (W - 1)), float(B * D * (H - 1) * W)


def tv_backward(x: torch.Tensor, scale: torch.Tensor, dx: Optional[torch.Tensor] = None,
                dx_scale: Optional[torch.Tensor] = None) -> torch.Tensor:
    """dx (= or +=) scale[0] * d(sum_h)/dx + scale[1] * d(sum_v)/dx; if dx is given it is first
    multiplied by dx_scale (device scalar) -- the fused late upstream scaling."""
    x = x.contiguous()
    B, D, H, W = x.shape
    acc = 1
    if dx is None:
        dx = torch.empty_like(x)
        acc = 0
    scale = scale.detach().to(device=x.device, dtype=torch.float32).contiguous()
    ds = None if dx_scale is None else dx_scale.detach().reshape(1).to(device=x.device, dtype=torch.float32)
    check(_lib.lib().rc_tv_bwd(_p(x), _dt(x), B * D, H, W, _p(scale), _p(dx), acc, _p(ds), _stream(x)), "rc_tv_bwd")
    return dx


class _Smoothness(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, denominators=None):
        sums = tv_sums(x.detach())
        dh, dv = tv_denominators(x.shape) if denominators is None else denominators
        ctx.den = (dh, dv)
        ctx.save_for_backward(x)
        # l1_loss over an empty slice is NaN in the reference (W == 1 or H == 1); keep that
        th = sums[0] / dh if dh > 0 else torch.full((), float("nan"), device=x.device, dtype=torch.float64)
        tv = sums[1] / dv if dv > 0 else torch.full((), float("nan"), device=x.device, dtype=torch.float64)
        return (th + tv).float()

    @staticmethod
    def backward(ctx, g):
        (x,) = ctx.saved_tensors
        dh, dv = ctx.den
        scale = torch.stack([g.float() / dh if dh > 0 else g.float() * 0, g.float() / dv if dv > 0 else g.float() * 0])
        return tv_backward(x.detach(), scale), None


def smoothness(x: torch.Tensor, denominators=None) -> torch.Tensor:
    """sum_h / dh + sum_v / dv; the denominators default to the element counts of x's own slices
    (model.py:332-333) and can be overridden when x stands for a larger tensor (shared 2x2 blocks)."""
    return _Smoothness.apply(x, denominators)


# ----------------------------------------------------------------------------------------------
# masked pooling
# ----------------------------------------------------------------------------------------------

def _seg_views(seg: torch.Tensor, B: int):
    """One label map per pass: a tensor, or a list of tensors that label the SAME embedding rows (shared 2x2
    blocks: the four full-resolution label sub-grids of a half-resolution embedding map)."""
    segs = list(seg) if isinstance(seg, (list, tuple)) else [seg]
    return [s_.reshape(B, -1).to(torch.int64).contiguous() for s_ in segs]


def pool_forward(x: torch.Tensor, seg, lut: torch.Tensor, lut_per_image: bool, n_slots: int):
    """sum/count per slot then mean; lut [B,C] (per image) or [C] (batch-wide) int32, -1 = none.
    ``seg`` may be a list of label maps over the same rows: every pass adds into the same sums and counts."""
    _need_cuda(x, lut)
    x, B, D, HW = _emb3(x)
    segs = _seg_views(seg, B)
    lut = lut.to(torch.int32).contiguous()
    C = lut.shape[-1]
    out = torch.zeros(n_slots, D, device=x.device, dtype=torch.float32)
    cnt = torch.zeros(max(n_slots, 1), device=x.device, dtype=torch.int32)
    L = _lib.lib()
    for sg in segs:
        _need_cuda(sg)
        check(L.rc_pool_fwd(_p(x), _dt(x), B, D, HW, _p(sg), _p(lut), C if lut_per_image else 0, C, n_slots,
                            _p(out), _p(cnt), _stream(x)), "rc_pool_fwd")
    check(L.rc_pool_finish(_p(out), _p(cnt), n_slots, D, _stream(x)), "rc_pool_finish")
    return out, cnt


def pool_backward(g: torch.Tensor, cnt: torch.Tensor, seg, lut: torch.Tensor, lut_per_image: bool,
                  shape, dtype) -> torch.Tensor:
    B, D = shape[0], shape[1]
    HW = 1
    for s in shape[2:]:
        HW *= s
    segs = _seg_views(seg, B)
    lut = lut.to(torch.int32).contiguous()
    C = lut.shape[-1]
    g = g.float().contiguous()
    dx = torch.empty(shape, device=g.device, dtype=dtype)
    for i, sg in enumerate(segs):          # later passes accumulate: a row's gradient is the sum over its label maps
        check(_lib.lib().rc_pool_bwd(_p(g), _p(cnt), B, D, HW, _p(sg), _p(lut), C if lut_per_image else 0, C,
                                     g.shape[0], _p(dx), RC_F32 if dtype == torch.float32 else RC_BF16, 1 if i else 0,
                                     _stream(g)), "rc_pool_bwd")
    return dx


class _MaskedPool(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, seg, lut, lut_per_image, n_slots):
        out, cnt = pool_forward(x.detach(), seg, lut, lut_per_image, n_slots)
        segs = list(seg) if isinstance(seg, (list, tuple)) else [seg]
        ctx.save_for_backward(cnt, lut, *segs)
        ctx.meta = (lut_per_image, tuple(x.shape), x.dtype)
        return out.to(x.dtype) if x.dtype != torch.float32 else out

    @staticmethod
    def backward(ctx, g):
        cnt, lut, *segs = ctx.saved_tensors
        lut_per_image, shape, dtype = ctx.meta
        return pool_backward(g, cnt, segs, lut, lut_per_image, shape, dtype), None, None, None, None


def masked_pool(x, seg, lut, lut_per_image, n_slots):
    return _MaskedPool.apply(x, seg, lut, lut_per_image, n_slots)


# ----------------------------------------------------------------------------------------------
# evaluation
# ----------------------------------------------------------------------------------------------

def topk_bf16_supported(D: int, HW: int) -> bool:
    return D % 64 == 0 and 64 <= D <= 512 and HW % 8 == 0 and HW > 0


def eval_topk(x: torch.Tensor, t_norm: torch.Tensor, index_map: torch.Tensor, k: int, precision="auto", t_bf16=None):
    """out[b, j, h, w] = index_map[j-th best text row by cosine logit] (model.py:164-173).
    precision 'fp32' = CUDA-core kernel (reference arithmetic), 'bf16' = tcgen05 kernel, 'auto' = bf16 when
    the shape allows it."""
    _need_cuda(x, t_norm, index_map)
    x, B, D, HW = _emb3(x)
    K = t_norm.shape[0]
    k = min(k, K)
    out = torch.empty((B, k) + tuple(x.shape[2:]), device=x.device, dtype=torch.int64)
    index_map = index_map.to(torch.int64).contiguous()
    if precision == "auto":
        precision = "bf16" if topk_bf16_supported(D, HW) else "fp32"
    L = _lib.lib()
    if precision == "fp32":
        xf = x if x.dtype == torch.float32 else x.float()
        tf = t_norm.float().contiguous()
        check(L.rc_eval_topk_f32(_p(xf), B, D, HW, D * HW, _p(tf), K, _p(index_map), k, _p(out), _stream(x)),
              "rc_eval_topk_f32")
    elif precision == "bf16":
        if not topk_bf16_supported(D, HW):
            raise RuntimeError(f"eval_topk: bf16 tensor-core path does not cover D={D}, HW={HW}")
        tb = t_bf16 if t_bf16 is not None else text_to_bf16(t_norm)[0]
        xdt = _dt(x)
        ws, ws_bytes = None, 0
        if xdt == RC_F32:
            ws_bytes = int(L.rc_infonce_workspace_bytes(B, D, HW, K, xdt))
            ws = torch.empty(ws_bytes, device=x.device, dtype=torch.uint8)
        check(L.rc_eval_topk_bf16(_p(x), xdt, B, D, HW, _p(tb), K, _p(index_map), k, _p(out), _p(ws), ws_bytes, _stream(x)),
              "rc_eval_topk_bf16")
    else:
        raise RuntimeError(f"eval_topk: unknown precision {precision!r}")
    return out


def eval_topk_hist(x: torch.Tensor, t_norm: torch.Tensor, index_map: torch.Tensor, k: int, gt: torch.Tensor,
                   E_u8: torch.Tensor, cmap: torch.Tensor, hist: torch.Tensor, counters: torch.Tensor, t_bf16=None,
                   want_ids: bool = True):
    """Fused evaluation batch (model.py:164-173 + validate.py:88-139): top-k on the tensor cores with the five class
    histograms and three counters built from the ids while they are still in registers (added to ``hist`` /
    ``counters``, int64).  Returns the ids [B,k,H,W] when ``want_ids`` (the drop-in ``predict`` needs them), else None.
    Tensor-core path only (D % 64 == 0, D <= 512, HW % 8 == 0)."""
    _need_cuda(x, t_norm, index_map, gt, E_u8, cmap, hist, counters)
    x, B, D, HW = _emb3(x)
    if not topk_bf16_supported(D, HW):
        raise RuntimeError(f"eval_topk_hist: the fused kernel does not cover D={D}, HW={HW}; use eval_topk + eval_hist")
    K = t_norm.shape[0]
    k = min(k, K)
    C = cmap.numel()
    gt = gt.reshape(-1).to(torch.int64).contiguous()
    if gt.numel() != B * HW:
        raise RuntimeError("eval_topk_hist: gt must have one entry per pixel")
    if hist.dtype != torch.int64 or counters.dtype != torch.int64 or tuple(hist.shape) != (5, C) or counters.numel() != 3:
        raise RuntimeError("eval_topk_hist: hist must be int64 [5, C] and counters int64 [3]")
    out = torch.empty((B, k) + tuple(x.shape[2:]), device=x.device, dtype=torch.int64) if want_ids else None
    index_map = index_map.to(torch.int64).contiguous()
    tb = t_bf16 if t_bf16 is not None else text_to_bf16(t_norm)[0]
    L = _lib.lib()
    xdt = _dt(x)
    ws, ws_bytes = None, 0
    if xdt == RC_F32:
        ws_bytes = int(L.rc_infonce_workspace_bytes(B, D, HW, K, xdt))
        ws = torch.empty(ws_bytes, device=x.device, dtype=torch.uint8)
    check(L.rc_eval_topk_hist_bf16(_p(x), xdt, B, D, HW, _p(tb), K, _p(index_map), k, _p(out), _p(gt), _p(E_u8.contiguous()),
                                   _p(cmap.to(torch.int64).contiguous()), C, _p(hist), _p(counters), _p(ws), ws_bytes,
                                   _stream(x)), "rc_eval_topk_hist_bf16")
    return out


def eval_hist(gt: torch.Tensor, topk: torch.Tensor, E_u8: torch.Tensor, cmap: torch.Tensor,
              hist: Optional[torch.Tensor] = None, counters: Optional[torch.Tensor] = None):
    """One batch of validate.py:88-139 as five class histograms + three counters (int64, added to)."""
    _need_cuda(gt, topk, E_u8, cmap)
    B, k = topk.shape[0], topk.shape[1]
    HW = topk[0, 0].numel()
    C = cmap.numel()
    gt = gt.reshape(-1).to(torch.int64).contiguous()
    topk = topk.to(torch.int64).contiguous()
    if gt.numel() != B * HW:
        raise RuntimeError("eval_hist: gt and topk disagree on the number of pixels")
    if hist is None:
        hist = torch.zeros(5, C, device=gt.device, dtype=torch.int64)
    if counters is None:
        counters = torch.zeros(3, device=gt.device, dtype=torch.int64)
    check(_lib.lib().rc_eval_hist(_p(gt), _p(topk), B, HW, k, _p(E_u8), _p(cmap), C, _p(hist), _p(counters),
                                  _stream(gt)), "rc_eval_hist")
    return hist, counters


def eval_fold(batch_hist: torch.Tensor, batch_index: int, acc: torch.Tensor, first_seen: torch.Tensor) -> None:
    C = batch_hist.shape[1]
    check(_lib.lib().rc_eval_fold(_p(batch_hist), C, int(batch_index), _p(acc), _p(first_seen), _stream(acc)),
          "rc_eval_fold")


def debug_umma_gemm(a: torch.Tensor, b: torch.Tensor, variant: int) -> torch.Tensor:
    """Bring-up check of the TMA + tcgen05 + TMEM path: C[128,N] = A B^T in bf16 -> f32."""
    _need_cuda(a, b)
    N, Kd = b.shape
    c = torch.empty(128, N, device=a.device, dtype=torch.float32)
    check(_lib.lib().rc_debug_umma_gemm(_p(a.contiguous()), _p(b.contiguous()), N, Kd, variant, _p(c), _stream(a)),
          "rc_debug_umma_gemm")
    return c


def debug_umma_gemm_2sm(a: torch.Tensor, b: torch.Tensor) -> torch.Tensor:
    """Bring-up check of the CTA-pair (cta_group::2) path: C[256,N] = A[256,Kd] B[N,Kd]^T."""
    _need_cuda(a, b)
    N, Kd = b.shape
    c = torch.empty(256, N, device=a.device, dtype=torch.float32)
    check(_lib.lib().rc_debug_umma_gemm_2sm(_p(a.contiguous()), _p(b.contiguous()), N, Kd, _p(c), _stream(a)),
          "rc_debug_umma_gemm_2sm")
    return c
