"""Edge shapes of every op against the CPU oracle: tiny, odd, very wide and many-image inputs that the parity tests do
not enumerate (one tile smaller than a pixel group, K = 1 / 255 / 257, rows wider than a shared-memory tile, hundreds
of 8-pixel images, empty sample lists).  Tolerances as in test_gpu_parity.py."""
import numpy as np
import pytest
import torch

from oracle import rangeclip_oracle as O

pytestmark = pytest.mark.gpu


def dev():
    return torch.device("cuda:0")


def maxrel(a, b):
    a = np.asarray(a, dtype=np.float64)
    b = np.asarray(b, dtype=np.float64)
    return float(np.abs(a - b).max() / max(np.abs(b).max(), 1e-30))


@pytest.mark.parametrize("shape,dt", [((1, 1, 3, 2048), torch.bfloat16), ((1, 1, 2, 4096), torch.float32), ((1, 2, 1, 8), torch.float32),
                                      ((1, 2, 8, 1), torch.float32), ((1, 1, 300, 8), torch.bfloat16), ((2, 3, 7, 2040), torch.bfloat16)])
def test_smoothness_edge_shapes(shape, dt):
    from rangeclip_b200 import ops
    g = torch.Generator().manual_seed(sum(shape))
    x = torch.randn(shape, generator=g).to(dt)
    ref = O.smoothness(x.double())
    gref = O.smoothness_grad(x.float())
    xd = x.to(dev()).requires_grad_(True)
    loss = ops.smoothness(xd)
    loss.backward()
    if np.isfinite(float(ref)):
        assert abs(float(loss.detach()) - float(ref)) <= 1e-5 * abs(float(ref))
    else:                       # H == 1 or W == 1: l1_loss over an empty slice is NaN in the reference
        assert not np.isfinite(float(loss.detach()))
    assert maxrel(xd.grad.float().cpu(), gref) < (1e-6 if dt == torch.float32 else 1e-2)


@pytest.mark.parametrize("B,D,H,W,dt", [(1, 1, 1, 8, torch.float32), (1, 3, 2, 4, torch.bfloat16), (2, 65, 4, 64, torch.bfloat16),
                                        (1, 520, 2, 136, torch.float32), (1, 8, 1, 264, torch.bfloat16), (3, 64, 1, 520, torch.bfloat16)])
def test_pooling_edge_shapes(B, D, H, W, dt):
    import rangeclip_b200 as R
    g = torch.Generator().manual_seed(B + D + H + W)
    x = torch.randn(B, D, H, W, generator=g).to(dt).float()
    seg = torch.randint(0, 4, (B, H, W), generator=g)
    items = [b for b in range(B) for _ in range(4)]
    labels = [lab for _ in range(B) for lab in range(4)]
    ref = O.area_pool_per_image(x.double(), seg, items, labels)
    out = R.pool_objects_per_image(x.to(dev()).to(dt), seg.to(dev()), items, labels)
    assert maxrel(out.float().cpu(), ref) < (2e-6 if dt == torch.float32 else 5e-3)


@pytest.mark.parametrize("B,D,H,W,K,rep", [(1, 256, 1, 8, 1, 1), (1, 512, 1, 8, 256, 1), (300, 256, 1, 8, 5, 1), (1, 512, 3, 8, 255, 1),
                                           (2, 128, 1, 8, 2, 1), (1, 384, 2, 64, 65, 1), (1, 512, 1, 136, 129, 1),
                                           (1, 256, 1, 8, 3, 4), (130, 512, 1, 8, 256, 4), (1, 512, 5, 40, 77, 4)])
def test_infonce_tensor_core_edge_shapes(B, D, H, W, K, rep):
    from rangeclip_b200 import ops
    g = torch.Generator().manual_seed(B + D + H + W + K + rep)
    x = torch.randn(B, D, H, W, generator=g).to(torch.bfloat16).float()
    t = torch.nn.functional.normalize(torch.randn(K, D, generator=g), dim=1).to(torch.bfloat16).float()
    M = B * H * W
    y = torch.randint(-1, K, (M, rep), generator=g, dtype=torch.int32)
    w = torch.randint(0, 3, (M, rep), generator=g).float()
    rows = x.permute(0, 2, 3, 1).reshape(-1, D)
    ref = O.infonce_dense_rep(rows, t, y, w, 1 / 0.07) if rep == 4 else O.infonce_dense(rows, t, y.reshape(-1), w.reshape(-1), 1 / 0.07)
    r = ops.infonce_raw(x.to(dev()).to(torch.bfloat16), t.to(dev()), y.to(dev()), w.to(dev()), 1 / 0.07, True, False, "bf16", rep=rep)
    torch.cuda.synchronize()
    loss = float(r["loss_sum"] / r["w_sum"]) if float(r["w_sum"]) > 0 else 0.0
    assert abs(loss - float(ref["loss"])) <= 2e-3 * abs(float(ref["loss"])) + 1e-5
    dref = ref["dx"].reshape(B, H, W, D).permute(0, 3, 1, 2)
    if float(dref.abs().max()) > 1e-9:
        assert maxrel(r["dx"].float().cpu(), dref) < 2e-2
    else:                       # K = 1: the gradient is identically zero
        assert float(r["dx"].float().abs().max()) < 1e-6


@pytest.mark.parametrize("B,D,H,W,K,k", [(1, 64, 1, 8, 1, 1), (1, 512, 1, 8, 5, 5), (2, 128, 2, 8, 2000, 5), (1, 256, 1, 136, 257, 8),
                                         (70, 64, 1, 8, 300, 3),
                                         # an ODD number of 128-pixel tiles: the CTA-pair form scans a tile past the end and must not write it
                                         (1, 256, 1, 264, 600, 5), (5, 128, 1, 8, 1024, 1), (3, 512, 3, 128, 300, 5)])
def test_eval_topk_edge_shapes(B, D, H, W, K, k):
    from rangeclip_b200 import ops
    g = torch.Generator().manual_seed(B + D + W + K + k)
    x = torch.randn(B, D, H, W, generator=g).to(torch.bfloat16)
    t = torch.nn.functional.normalize(torch.randn(K, D, generator=g), dim=1).to(torch.bfloat16).float()
    out = ops.eval_topk(x.to(dev()), t.to(dev()), torch.arange(K, device=dev()), k, "bf16").cpu()
    logits = torch.einsum("bdn,cd->bcn", x.float().view(B, D, H * W), t)
    ref = logits.topk(k, dim=1)
    got = logits.gather(1, out.view(B, k, H * W))
    assert float((got - ref.values).abs().max()) < 1e-5        # same logits rank by rank (ids may differ on exact ties)


def test_sample_weights_without_draws():
    from rangeclip_b200 import ops
    g = torch.Generator().manual_seed(1)
    seg = torch.randint(0, 5, (2, 64), generator=g).to(dev())
    lm = (torch.arange(5, dtype=torch.int32) - 1).to(dev())
    w, y = ops.sample_weights(seg, None, lm)                          # dense mode: weight 1 on every foreground pixel
    assert float(w.sum()) == float((seg > 0).sum())
    w0, _ = ops.sample_weights(seg, torch.zeros(2, 0, dtype=torch.int64, device=dev()), lm)
    assert w0.shape == w.shape


def test_fused_topk_histograms_with_an_odd_tile_count():
    """rc_eval_topk_hist_bf16 in the CTA-pair form on 3 tiles (the pair's fourth tile lies past the end): the histograms and
    counters equal those of the returned ids (nothing from the phantom tile is counted)."""
    from rangeclip_b200 import ops
    g = torch.Generator().manual_seed(9)
    B, D, H, W, K, k, C = 1, 256, 3, 128, 600, 5, 600
    t = torch.nn.functional.normalize(torch.randn(K, D, generator=g), dim=1).to(dev())
    gt = torch.randint(0, C, (B, H, W), generator=g).to(dev())
    x = (t[gt].permute(0, 3, 1, 2) + 0.3 * torch.randn(B, D, H, W, generator=g).to(dev())).to(torch.bfloat16)
    E = torch.eye(C, dtype=torch.uint8, device=dev()); cmap = torch.arange(C, device=dev())
    hist = torch.zeros(5, C, dtype=torch.int64, device=dev()); cnt = torch.zeros(3, dtype=torch.int64, device=dev())
    ids = ops.eval_topk_hist(x, t, torch.arange(K, device=dev()), k, gt, E, cmap, hist, cnt)
    hist2, cnt2 = ops.eval_hist(gt, ids, E, cmap)
    assert torch.equal(hist, hist2) and torch.equal(cnt, cnt2) and int(cnt[2]) == B * H * W


@pytest.mark.parametrize("B,D,H,W", [(2, 256, 6, 32), (1, 512, 7, 24), (3, 128, 1, 8), (1, 256, 9, 264), (2, 384, 4, 8), (1, 256, 5, 512)])
def test_fused_prepass_smoothness_sums(B, D, H, W):
    """fp32 X: rc_infonce_prepass_tv (bf16 copy + row norms + smoothness sums from one read) against the separate pre-pass and
    rc_tv_fwd -- strips that end before a multiple of 4 rows, rows narrower / wider than a warp, units that do not fill a warp."""
    from rangeclip_b200 import ops
    dev = torch.device("cuda:0")
    g = torch.Generator(device=dev).manual_seed(B * 1000 + H * 10 + W)
    x = torch.randn(B, D, H, W, device=dev, generator=g)
    x[:, :, :, ::3] = x[:, :, :, 1::3][:, :, :, : x[:, :, :, ::3].shape[3]] if W > 8 else x[:, :, :, ::3]     # some exact zeros in the differences
    K = 40
    t = torch.nn.functional.normalize(torch.randn(K, D, device=dev, generator=g), dim=1)
    y = torch.randint(0, K, (B * H * W,), device=dev, generator=g, dtype=torch.int32)
    w = torch.rand(B * H * W, device=dev, generator=g)
    a = ops.infonce_raw(x, t, y, w, 1 / 0.07, True, False, "bf16", fuse_tv=True)
    b = ops.infonce_raw(x, t, y, w, 1 / 0.07, True, False, "bf16", fuse_tv=False)
    assert a["tv_sums"] is not None and b["tv_sums"] is None
    assert torch.equal(a["lse"], b["lse"]) and torch.allclose(a["dx"], b["dx"], rtol=1e-2, atol=1e-7)
    assert abs(float(a["loss_sum"]) - float(b["loss_sum"])) <= 1e-9 * abs(float(b["loss_sum"]))
    ref = ops.tv_sums(x)
    xd = x.double()
    exact = torch.stack([(xd[..., 1:] - xd[..., :-1]).abs().sum(), (xd[:, :, 1:] - xd[:, :, :-1]).abs().sum()])
    assert torch.allclose(a["tv_sums"], exact, rtol=2e-6, atol=1e-9), (a["tv_sums"], exact)
    assert torch.allclose(ref, exact, rtol=2e-6, atol=1e-9)
    # and through the operator: same losses, same gradient as the two-pass form
    xg = x.clone().requires_grad_(True)
    lt = torch.log(torch.tensor(0.07, device=dev))
    loss, smooth = ops.pixel_losses(xg, t, lt, y, w, "bf16")
    dh, dv = ops.tv_denominators(x.shape)
    want = (exact[0] / dh if dh > 0 else float("nan")) + (exact[1] / dv if dv > 0 else float("nan"))
    if dh > 0 and dv > 0:
        assert abs(float(smooth) - float(want)) <= 2e-6 * abs(float(want))
        (loss + 3.0 * smooth).backward()
        # the backward runs from the 4-bit difference signs kept by the pre-pass; the two-pass form reads x again
        scale = torch.tensor([3.0 / dh, 3.0 / dv], device=dev)
        one = torch.ones((), device=dev)
        want_g = ops.tv_backward(x, scale, b["dx"].view(x.shape), one)
        got_g = ops.tv_backward_codes(a["tv_codes"], scale, b["dx"].view(x.shape), one)
        assert torch.allclose(got_g, want_g, rtol=1e-6, atol=1e-10), float((got_g - want_g).abs().max())
        # (the operator's own launch may round single dX values to the neighbouring bf16 -- the weight sum is an atomic double
        # sum, exp(-log(tau)) is not 1 / tau to the last bit -- and the smoothness term may cancel most of such a value)
        ulp = 2.0 ** -7 * float(b["dx"].abs().max())
        assert torch.allclose(xg.grad, want_g, rtol=1e-2, atol=ulp), float((xg.grad - want_g).abs().max())
