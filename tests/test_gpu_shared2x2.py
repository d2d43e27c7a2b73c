"""GPU parity of the shared-embedding (2x2 block) path, SURVEY 8(f)-1: `rc_infonce_bf16_rep4` and
`compute_loss_shared2x2` against the oracle, against the fixtures produced by the reference decoder tail +
`compute_loss` (tests/golden/make_golden_up2.py), and against the full-resolution kernel on the upsampled
tensor.  bf16 tensor-core path: 2e-2 max-relative on gradients, 2e-3 on loss / lse (BASELINE.json)."""
import os
import random
from unittest import mock

import numpy as np
import pytest
import torch

from oracle import rangeclip_oracle as O

pytestmark = pytest.mark.gpu

BF16_MAXREL = 2e-2


def dev():
    return torch.device("cuda:0")


def maxrel(a, b):
    a = np.asarray(a, dtype=np.float64)
    b = np.asarray(b, dtype=np.float64)
    return float(np.abs(a - b).max() / max(np.abs(b).max(), 1e-30))


def _rep4_case(B, D, h, w, K, seed, tau=0.07):
    g = torch.Generator().manual_seed(seed)
    x = (torch.randn(B, D, h, w, generator=g) * (0.5 + torch.rand(B, 1, h, w, generator=g))).to(torch.bfloat16).float()
    t = torch.nn.functional.normalize(torch.randn(K, D, generator=g), dim=1).to(torch.bfloat16).float()
    M = B * h * w
    base = torch.randint(0, K, (M, 1), generator=g, dtype=torch.int32)
    y4 = base.repeat(1, 4)                                   # mostly one label per block ...
    mixed = torch.rand(M, 4, generator=g) < 0.3              # ... with mixed blocks, duplicates and ignored pixels
    y4 = torch.where(mixed, torch.randint(0, K, (M, 4), generator=g, dtype=torch.int32), y4)
    y4[torch.rand(M, 4, generator=g) < 0.15] = -1
    y4[torch.rand(M, generator=g) < 0.05] = -1               # blocks without any target
    w4 = torch.randint(0, 4, (M, 4), generator=g).float()
    return x, t, y4, w4, 1.0 / tau


def _oracle_rep4(x, t, y4, w4, inv_tau):
    B, D, h, w = x.shape
    r = O.infonce_dense_rep(x.permute(0, 2, 3, 1).reshape(-1, D), t, y4, w4, inv_tau)
    r["dx4"] = r["dx"].reshape(B, h, w, D).permute(0, 3, 1, 2)
    return r


@pytest.mark.parametrize("B,D,h,w,K,xdtype,tau", [
    (2, 512, 16, 16, 256, torch.bfloat16, 0.07), (2, 256, 16, 24, 100, torch.bfloat16, 0.07),
    (1, 512, 20, 20, 200, torch.float32, 0.07), (5, 512, 9, 8, 130, torch.bfloat16, 0.07),
    (2, 256, 5, 8, 33, torch.bfloat16, 0.07), (2, 512, 16, 16, 256, torch.bfloat16, 0.02),
    (1, 512, 8, 8, 2, torch.bfloat16, 0.07)])
def test_infonce_rep4_tensor_core(B, D, h, w, K, xdtype, tau):
    from rangeclip_b200 import ops
    x, t, y4, w4, inv_tau = _rep4_case(B, D, h, w, K, seed=B * 91 + D + K + h, tau=tau)
    ref = _oracle_rep4(x, t, y4, w4, inv_tau)
    r = ops.infonce_raw(x.to(dev()).to(xdtype), t.to(dev()), y4.to(dev()), w4.to(dev()), inv_tau, True, False, "bf16", rep=4)
    torch.cuda.synchronize()
    loss = float(r["loss_sum"] / r["w_sum"])
    tol = 3e-3 if tau < 0.029 else 2e-3
    assert float(r["w_sum"]) == float(ref["wsum"])
    assert abs(loss - float(ref["loss"])) <= tol * abs(float(ref["loss"])) + 1e-6, (loss, float(ref["loss"]))
    assert maxrel(r["lse"].cpu(), ref["lse"]) < tol
    assert maxrel(r["dx"].float().cpu(), ref["dx4"]) < BF16_MAXREL, maxrel(r["dx"].float().cpu(), ref["dx4"])
    assert abs(float(r["dlogtau"]) - float(ref["dlogtau"])) <= BF16_MAXREL * abs(float(ref["dlogtau"])) + 1e-7
    # forward-only launch gives the same loss
    r2 = ops.infonce_raw(x.to(dev()).to(xdtype), t.to(dev()), y4.to(dev()), w4.to(dev()), inv_tau, False, False, "bf16", rep=4)
    assert abs(float(r2["loss_sum"] / r2["w_sum"]) - loss) <= 1e-6 * abs(loss) + 1e-9


def test_infonce_rep4_dtext():
    from rangeclip_b200 import ops
    x, t, y4, w4, inv_tau = _rep4_case(2, 512, 16, 16, 200, seed=77)
    ref = _oracle_rep4(x, t, y4, w4, inv_tau)
    r = ops.infonce_raw(x.to(dev()).to(torch.bfloat16), t.to(dev()), y4.to(dev()), w4.to(dev()), inv_tau, True, True, "bf16", rep=4)
    assert maxrel(r["dt"].cpu(), ref["dt"]) < BF16_MAXREL
    assert maxrel(r["dx"].float().cpu(), ref["dx4"]) < BF16_MAXREL


def test_rep4_equals_full_resolution_kernel():
    """Size-independent property at 128 x 128 shared rows (= one 256 x 256 map): the four-target launch on the
    distinct rows agrees with the one-target launch on the nearest-upsampled tensor -- same loss, and the row
    gradient is the sum of the block's four pixel gradients."""
    from rangeclip_b200 import ops
    from rangeclip_b200.losses import group_2x2
    g = torch.Generator(device="cuda").manual_seed(5)
    B, D, h, w, K = 1, 512, 128, 128, 256
    H, W = 2 * h, 2 * w
    e = torch.randn(B, D, h, w, device=dev(), generator=g).to(torch.bfloat16)
    t = torch.nn.functional.normalize(torch.randn(K, D, device=dev(), generator=g), dim=1)
    seg = torch.randint(0, K, (B, H // 8, W // 8), device=dev(), generator=g).repeat_interleave(8, 1).repeat_interleave(8, 2)
    seg = torch.roll(seg, (3, 5), (1, 2)).to(torch.int32)            # label borders cut through embedding blocks
    wt = torch.randint(0, 3, (B, H, W), device=dev(), generator=g).float()
    x_hi = e.repeat_interleave(2, 2).repeat_interleave(2, 3).contiguous()
    r_hi = ops.infonce_raw(x_hi, t, seg.reshape(B, -1), wt.reshape(B, -1), 1 / 0.07, True, False, "bf16")
    r_lo = ops.infonce_raw(e, t, group_2x2(seg), group_2x2(wt), 1 / 0.07, True, False, "bf16", rep=4)
    assert float(r_hi["w_sum"]) == float(r_lo["w_sum"])
    l_hi, l_lo = float(r_hi["loss_sum"] / r_hi["w_sum"]), float(r_lo["loss_sum"] / r_lo["w_sum"])
    assert abs(l_hi - l_lo) <= 1e-5 * abs(l_hi)
    dsum = r_hi["dx"].float().view(B, D, h, 2, w, 2).sum(dim=(3, 5))
    assert maxrel(r_lo["dx"].float().cpu(), dsum.cpu()) < BF16_MAXREL
    assert abs(float(r_hi["dlogtau"]) - float(r_lo["dlogtau"])) <= 1e-3 * abs(float(r_hi["dlogtau"]))


class _Model(torch.nn.Module):
    def __init__(self):
        super().__init__()
        self.log_temperature_text = torch.nn.Parameter(torch.log(torch.tensor(0.07)))
        self.log_temperature_image = torch.nn.Parameter(torch.log(torch.tensor(0.1)))


@pytest.mark.parametrize("case", ["a", "b"])
def test_compute_loss_shared2x2_golden(golden_dir, case):
    """Reference: decoder tail (decoder.py:112-116) + DepthUNet.compute_loss, gradient w.r.t. the decoder output."""
    import rangeclip_b200 as R
    g = np.load(os.path.join(golden_dir, f"up2_{case}.npz"))
    model = _Model().to(dev())
    E = torch.tensor(g["E"]).to(dev()).requires_grad_(True)
    C = g["text"].shape[0]
    sets = {"medium": {i: [int(v) for v in g["medium"][i]] for i in range(C)},
            "hard": {i: [int(v) for v in g["hard"][i]] for i in range(C)}}
    rand_idx = torch.tensor(g["rand_idx"]).to(dev())
    seed = int(g["seed"])
    np.random.seed(seed); torch.manual_seed(seed); random.seed(seed)
    with mock.patch("torch.randint", lambda *a, **k: rand_idx):
        total, info = R.compute_loss_shared2x2(model, E, torch.tensor(g["seg"]).to(dev()), torch.tensor(g["text"]).to(dev()),
                                               sets, None, None, W_smooth=float(g["W_smooth"]),
                                               percent_image_sampling=float(g["pct_sampling"]),
                                               k_distractors=int(g["k_distractors"]))
    total.backward()
    assert abs(info["text_contrastive_loss"] - float(g["text_loss"])) <= 2e-3 * abs(float(g["text_loss"]))
    assert abs(info["smoothness_loss"] - float(g["smooth_loss"])) <= 1e-5 * abs(float(g["smooth_loss"])) + 1e-12
    assert abs(float(total) - float(g["total"])) <= 2e-3 * abs(float(g["total"]))
    assert maxrel(E.grad.cpu(), g["dE"]) < BF16_MAXREL, maxrel(E.grad.cpu(), g["dE"])
    assert abs(float(model.log_temperature_text.grad) - float(g["dlogtau_text"])) <= BF16_MAXREL * abs(float(g["dlogtau_text"]))


def test_rep4_rejects_unsupported_shapes():
    from rangeclip_b200 import ops
    x = torch.zeros(1, 128, 8, 8, device=dev(), dtype=torch.bfloat16)
    t = torch.nn.functional.normalize(torch.randn(4, 128, device=dev()), dim=1)
    y4 = torch.zeros(64, 4, device=dev(), dtype=torch.int32)
    with pytest.raises(RuntimeError):
        ops.infonce_raw(x, t, y4, y4.float(), 10.0, True, False, "bf16", rep=4)


def test_area_pooling_shared2x2():
    """dataloader.py:286-304 on the decoder-tail output (decoder.py:113-114) against four accumulating pooling passes
    over the half-resolution rows, forward and gradient w.r.t. the decoder output."""
    import rangeclip_b200 as R
    g = torch.Generator().manual_seed(21)
    B, D, h, w = 3, 64, 12, 10
    H, W = 2 * h, 2 * w
    E = torch.randn(B, D, h, w, generator=g) * (0.5 + torch.rand(B, 1, h, w, generator=g))
    seg = torch.randint(0, 7, (B, H // 3 + 1, W // 3 + 1), generator=g).repeat_interleave(3, 1).repeat_interleave(3, 2)[:, :H, :W].contiguous()
    items, labels = [], []
    for b in range(B):
        for lab in torch.unique(seg[b]).tolist():
            items.append(b); labels.append(lab)
    items += [0, 1]; labels += [labels[0], 99]            # a duplicate and an absent label
    Er = E.double().requires_grad_(True)
    ref = O.area_pool_per_image(O.decoder_tail(Er, (H, W)), seg, items, labels)
    up = torch.randn(ref.shape, generator=g, dtype=torch.float64)
    (ref * up).sum().backward()
    Eg = E.to(dev()).requires_grad_(True)
    out = R.pool_objects_per_image(Eg, seg.to(dev()), items, labels, differentiable=True, shared2x2=True)
    (out * up.float().to(dev())).sum().backward()
    assert maxrel(out.detach().cpu(), ref.detach()) < 1e-5
    assert float(out[-1].abs().sum()) == 0.0
    assert maxrel(Eg.grad.cpu(), Er.grad) < 1e-5
    # the no-grad drop-in form returns the same values
    out2 = R.pool_objects_per_image(E.to(dev()), seg.to(dev()), items, labels, shared2x2=True)
    assert torch.equal(out2, out.detach())


def test_normalize_rows_handover_between_pooling_and_loss():
    """The no_grad normalisation of the area pooling is handed to the next call on the same data (one kernel instead of two per
    step) -- and only then: an in-place change of the tensor (version counter) or another tensor gets a fresh result."""
    from rangeclip_b200 import ops
    g = torch.Generator().manual_seed(4)
    x = torch.randn(2, 256, 8, 16, generator=g).to(torch.bfloat16).to(dev())
    n0 = _launches()
    with torch.no_grad():
        a = ops.normalize_rows(x.detach())
    xg = x.requires_grad_(True)
    b = ops.normalize_rows(xg)                              # same storage, same version: handed over
    assert _launches() - n0 == 1 and torch.equal(a, b) and b.requires_grad
    b.sum().backward()
    assert xg.grad is not None
    with torch.no_grad():
        a2 = ops.normalize_rows(x.detach())
        x.detach().mul_(2.0)                                # in-place change: the version counter moves
    c = ops.normalize_rows(x.detach())
    assert torch.allclose(c, a2, atol=1e-6) and c.data_ptr() != a2.data_ptr()
    with torch.no_grad():
        ops.normalize_rows(x.detach())
    y = torch.randn(2, 256, 8, 16, generator=g).to(torch.bfloat16).to(dev())
    d = ops.normalize_rows(y)
    assert torch.allclose(d, torch.nn.functional.normalize(y.float(), dim=1), atol=1e-6)


def _launches():
    from rangeclip_b200 import _lib
    return _lib.launch_count()


@pytest.mark.parametrize("dtype,shape", [(torch.float32, (2, 96, 8, 16)), (torch.bfloat16, (1, 256, 12, 24)), (torch.bfloat16, (2, 64, 1, 8)),
                                         (torch.float32, (1, 32, 5, 8))])
def test_smoothness_normalized_matches_the_two_step_path(dtype, shape):
    """smoothness(normalize(x)) with the fused backward (rc_tv_normalize_bwd) against PyTorch autograd of the same expression
    (F.normalize + l1 means, model.py:332-334 / decoder.py:114), including repeated values (sign(0) = 0) and zero rows."""
    from rangeclip_b200 import ops
    g = torch.Generator().manual_seed(sum(shape))
    x = torch.randn(shape, generator=g)
    x[..., 1::2] = x[..., 0::2]                     # horizontally repeated pixels, as after a nearest x2 upsample (quirk Q8)
    x[0, :, 0, 0] = 0
    x = x.to(dtype)
    B, D, H, W = shape
    xr = x.float().clone().requires_grad_(True)
    xh = torch.nn.functional.normalize(xr, p=2, dim=1)
    dh, dv = float(B * D * H * (W - 1)), float(B * D * (H - 1) * W)
    ref = (xh[..., :-1] - xh[..., 1:]).abs().sum() / dh
    if H > 1:
        ref = ref + (xh[:, :, :-1] - xh[:, :, 1:]).abs().sum() / dv
    (3.0 * ref).backward()
    xd = x.to(dev()).requires_grad_(True)
    if H > 1:
        out = ops.smoothness_normalized(xd)
    else:
        out = ops.smoothness_normalized(xd, denominators=(dh, 1.0))       # (H = 1: the vertical term is empty)
    (3.0 * out).backward()
    assert abs(float(out) - float(ref)) <= 1e-5 * abs(float(ref)) + 1e-9
    # the kernel's sign(0) = 0 is autograd's abs'(0) = 0; tolerance: fp32 path 1e-5, bf16 output 1e-2
    err = float((xd.grad.float().cpu() - xr.grad).abs().max() / xr.grad.abs().max())
    assert err < (1e-5 if dtype == torch.float32 else 1e-2), err
