"""Golden fixture for the shared-embedding (2x2 block) path, SURVEY 8(f)-1 / quirk Q8.

Run in the build container only:   python tests/golden/make_golden_up2.py

The UNMODIFIED reference decoder tail (utils/src/decoder.py:112-116: output_conv -> nearest
interpolate to the target shape -> L2 normalise) is run on a low-resolution tensor E (the decoder's
blocks and output_conv replaced by identities, so E plays the role of the output_conv result), the
result goes through the reference ``DepthUNet.compute_loss`` and the gradient is taken with respect
to E.  Stored: E, the inputs of compute_loss, the losses, dE and dlog(tau).
"""
import os
import random
from unittest import mock

import numpy as np
import torch

import make_golden as G                      # shims + reference imports (its __main__ is not run)
from utils.src.decoder import DepthDecoder   # noqa: E402  (sys.path set by make_golden)


def reference_tail(E, H, W):
    dec = DepthDecoder.__new__(DepthDecoder)
    torch.nn.Module.__init__(dec)
    dec.up_blocks = torch.nn.ModuleList([torch.nn.Identity()])
    dec.output_conv = torch.nn.Identity()
    return dec(E, [E], (H, W))               # decoder.py:98-116


def up2_case(name, seed, B=2, D=256, h=8, w=8, C=48, k_distractors=14, W_smooth=2e2, pct_sampling=0.7, blk=3):
    gen = torch.Generator().manual_seed(seed)
    H, W = 2 * h, 2 * w
    # bf16-exact, un-normalised low-resolution embeddings (the bf16 tensor-core path reads them unchanged)
    E = (torch.randn(B, D, h, w, generator=gen) * 1.7).to(torch.bfloat16).float().requires_grad_(True)
    # label blocks of odd size: 2x2 embedding blocks straddle label borders (mixed targets inside a block)
    gh, gw = -(-H // blk), -(-W // blk)
    pick = torch.randint(0, 9, (B, gh, gw), generator=gen)
    seg = pick.repeat_interleave(blk, 1).repeat_interleave(blk, 2)[:, :H, :W].contiguous()
    text = torch.randn(C, D, generator=gen)
    hard = {i: [int(v) for v in torch.randperm(C, generator=gen)[:6]] for i in range(C)}
    med = {i: [int(v) for v in torch.randperm(C, generator=gen)[:6]] for i in range(C)}
    sets = {"medium": med, "hard": hard}
    n = int(pct_sampling * H * W)
    rand_idx = torch.randint(0, H * W, (B, n), generator=gen)

    model = G.bare_model()
    uniq_calls = []
    real_unique = torch.unique

    def spy_unique(*a, **k):
        r = real_unique(*a, **k)
        uniq_calls.append(r)
        return r

    np.random.seed(seed)
    torch.manual_seed(seed)
    random.seed(seed)
    X = reference_tail(E, H, W)
    assert X.shape == (B, D, H, W)
    with mock.patch("torch.randint", lambda *a, **k: rand_idx), mock.patch("torch.unique", spy_unique):
        loss, info = model.compute_loss(X, seg, text, sets, None, None, W_text=1.0, W_image=0.5, W_smooth=W_smooth,
                                        percent_image_sampling=pct_sampling, k_distractors=k_distractors,
                                        pct_medium=0.0, pct_hard=0.75, pct_rand=0.25)
    contrast = uniq_calls[-1]
    loss.backward()
    mixed = (seg.view(B, h, 2, w, 2).permute(0, 1, 3, 2, 4).reshape(B, h * w, 4))
    n_mixed = int((mixed.min(-1).values != mixed.max(-1).values).sum())
    np.savez_compressed(
        os.path.join(G.OUT, f"up2_{name}.npz"),
        E=E.detach().numpy(), seg=seg.numpy(), text=text.numpy(),
        hard=np.array([hard[i] for i in range(C)]), medium=np.array([med[i] for i in range(C)]),
        rand_idx=rand_idx.numpy(), seed=seed, k_distractors=k_distractors, W_smooth=W_smooth,
        pct_sampling=pct_sampling, contrast=contrast.numpy(), total=loss.detach().numpy(),
        text_loss=info["text_contrastive_loss"], image_loss=info["image_contrastive_loss"],
        smooth_loss=info["smoothness_loss"], dE=E.grad.numpy(),
        dlogtau_text=model.log_temperature_text.grad.numpy(), n_mixed_blocks=n_mixed,
    )
    print(f"up2_{name}: K={len(contrast)} mixed blocks={n_mixed}/{B * h * w} total={float(loss):.6f} {info}")


if __name__ == "__main__":
    up2_case("a", seed=515)
    up2_case("b", seed=616, D=512, h=8, w=12, pct_sampling=1.0, W_smooth=0.0, blk=5)
