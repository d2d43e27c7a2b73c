"""Generate the golden fixtures in this directory by running the UNMODIFIED reference.

Run in the build container only (``/root/reference`` is not on the GPU box):

    python tests/golden/make_golden.py

The reference (jinryan/RangeCLIP) is imported from ``/root/reference`` with the shims
SURVEY.md section 8(c) lists (matplotlib stub, fake CLIP processor/encoder, pinned
``torch.randint``, a spy on ``torch.unique`` for the contrast set and a recording
``defaultdict`` for validate.py's four IoU dictionaries).  Everything written here is an
input the reference consumed or an output it produced; nothing is computed by this repo.
"""
import collections
import os
import random
import sys
import types
from unittest import mock

import numpy as np
import torch

REF_ROOT = os.environ.get("RANGECLIP_REFERENCE", "/root/reference")
OUT = os.path.dirname(os.path.abspath(__file__))

for m in ("matplotlib", "matplotlib.pyplot", "matplotlib.cm"):
    sys.modules[m] = types.ModuleType(m)
sys.modules["matplotlib"].pyplot = sys.modules["matplotlib.pyplot"]
sys.modules["matplotlib"].cm = sys.modules["matplotlib.cm"]
sys.path.insert(0, REF_ROOT)

from RangeCLIP.src.depth_segmentation_model.model import DepthUNet, masked_average_pooling  # noqa: E402
from RangeCLIP.src.depth_segmentation_model import validate as V  # noqa: E402
from RangeCLIP.src.depth_segmentation_model.dataloader import (  # noqa: E402
    prepare_image_contrast_data, build_equivalence_tensor, build_equivalence_class_map)
from utils.src.eval_utils import info_nce  # noqa: E402


def bare_model(tau_text=0.07, tau_image=0.1):
    """A DepthUNet without the backbone (only the loss / predict methods are exercised)."""
    m = DepthUNet.__new__(DepthUNet)
    torch.nn.Module.__init__(m)
    m.device = torch.device("cpu")
    m.log_temperature_text = torch.nn.Parameter(torch.log(torch.tensor(tau_text)))
    m.log_temperature_image = torch.nn.Parameter(torch.log(torch.tensor(tau_image)))
    return m


def unit(x, dim):
    return torch.nn.functional.normalize(x, dim=dim)


def block_seg(B, H, W, blk, labels, gen):
    """Per image a grid of blk x blk blocks, each with a label drawn from ``labels``."""
    gh, gw = H // blk, W // blk
    pick = torch.randint(0, len(labels), (B, gh, gw), generator=gen)
    lab = torch.as_tensor(labels)[pick]
    return lab.repeat_interleave(blk, 1).repeat_interleave(blk, 2).contiguous()


def loss_case(name, sim_form, seed, B=2, D=64, H=16, W=16, C=40, k_distractors=12,
              pcts=(0.0, 0.75, 0.25), n_obj=3, W_image=0.5, W_smooth=2e2, pct_sampling=0.7, bf16_exact=False):
    gen = torch.Generator().manual_seed(seed)
    X = unit(torch.randn(B, D, H, W, generator=gen), 1)
    if bf16_exact:      # values a bf16 tensor holds exactly: the tensor-core path then sees the reference's own inputs
        X = X.to(torch.bfloat16).float()
    X.requires_grad_(True)
    seg = block_seg(B, H, W, 4, list(range(0, 9)), gen)           # labels 0..8, 0 = background
    text = torch.randn(C, D, generator=gen)                        # un-normalised on purpose
    hard = {i: [int(v) for v in torch.randperm(C, generator=gen)[:6]] for i in range(C)}
    med = {i: [int(v) for v in torch.randperm(C, generator=gen)[:6]] for i in range(C)}
    if sim_form == "list":       # the loader's own format (dataloader.py:170-183) -> SURVEY Q3
        sets = {"medium": [med[i] for i in range(C)], "hard": [hard[i] for i in range(C)]}
    else:
        sets = {"medium": med, "hard": hard}
    area = torch.randn(n_obj, D, generator=gen) if n_obj else None
    img = torch.randn(n_obj, D, generator=gen) if n_obj else None
    n = int(pct_sampling * H * W)
    rand_idx = torch.randint(0, H * W, (B, n), generator=gen)

    model = bare_model()
    uniq_calls = []
    real_unique = torch.unique

    def spy_unique(*a, **k):
        r = real_unique(*a, **k)
        uniq_calls.append(r)
        return r

    np.random.seed(seed)
    torch.manual_seed(seed)
    random.seed(seed)
    with mock.patch("torch.randint", lambda *a, **k: rand_idx), mock.patch("torch.unique", spy_unique):
        loss, info = model.compute_loss(X, seg, text, sets, area, img, W_text=1.0, W_image=W_image,
                                        W_smooth=W_smooth, percent_image_sampling=pct_sampling,
                                        k_distractors=k_distractors, pct_medium=pcts[0],
                                        pct_hard=pcts[1], pct_rand=pcts[2])
    contrast = uniq_calls[-1]
    loss.backward()
    z = lambda t: np.zeros(0, np.float32) if t is None else t.detach().numpy()
    np.savez_compressed(
        os.path.join(OUT, f"loss_{name}.npz"),
        X=X.detach().numpy(), seg=seg.numpy(), text=text.numpy(),
        hard=np.array([hard[i] for i in range(C)]), medium=np.array([med[i] for i in range(C)]),
        sim_form=sim_form, area=z(area), img=z(img), rand_idx=rand_idx.numpy(), seed=seed,
        k_distractors=k_distractors, pcts=np.array(pcts), W_image=W_image, W_smooth=W_smooth,
        pct_sampling=pct_sampling, contrast=contrast.numpy(),
        total=loss.detach().numpy(), text_loss=info["text_contrastive_loss"],
        image_loss=info["image_contrastive_loss"], smooth_loss=info["smoothness_loss"],
        dX=X.grad.numpy(), dlogtau_text=model.log_temperature_text.grad.numpy(),
        dlogtau_image=(model.log_temperature_image.grad.numpy()
                       if model.log_temperature_image.grad is not None else np.zeros(())),
    )
    print(f"loss_{name}: K={len(contrast)} total={float(loss):.6f} {info}")


class FakeBatchFeature(dict):
    def to(self, device):
        return self


class FakeProcessor:
    """Stands in for CLIPProcessor: resizes every crop to 8x8."""

    def __call__(self, images, return_tensors="pt", padding=True, do_rescale=False):
        px = [torch.nn.functional.interpolate(im[None].float(), size=(8, 8), mode="bilinear",
                                              align_corners=False)[0] for im in images]
        return FakeBatchFeature(pixel_values=torch.stack(px))


class FakeEncoder:
    def __init__(self, D, seed):
        self.w = torch.randn(3 * 8 * 8, D, generator=torch.Generator().manual_seed(seed))

    def get_image_features(self, pixel_values):
        return pixel_values.reshape(pixel_values.shape[0], -1) @ self.w


def pool_case(seed=11, B=4, D=64, H=16, W=16):
    gen = torch.Generator().manual_seed(seed)
    X = unit(torch.randn(B, D, H, W, generator=gen), 1)
    seg = block_seg(B, H, W, 4, list(range(0, 7)), gen)
    image = torch.rand(B, 3, H, W, generator=gen)
    bbox = torch.tensor([[0, 0, 8, 8], [4, 2, 12, 16], [5, 5, 5, 9], [0, 0, 16, 16]])  # item 2 invalid
    label = torch.tensor([int(seg[0, 0, 0]), int(seg[1, 8, 8]), 3, 39])                # 39 absent -> zeros
    area, img = prepare_image_contrast_data(image, bbox, label, seg, X, FakeEncoder(D, seed),
                                            FakeProcessor(), torch.device("cpu"))
    objs = torch.tensor([1, 2, 5, 30])
    Xg = X.clone().requires_grad_(True)
    mp = masked_average_pooling(Xg, seg, objs)
    up = torch.randn(mp.shape, generator=gen)
    (mp * up).sum().backward()
    np.savez_compressed(os.path.join(OUT, "pool.npz"), X=X.numpy(), seg=seg.numpy(), image=image.numpy(),
                        bbox=bbox.numpy(), label=label.numpy(), area=area.numpy(), img=img.numpy(),
                        valid_items=np.array([0, 1, 3]), objs=objs.numpy(), mp=mp.detach().numpy(),
                        mp_upstream=up.numpy(), mp_dX=Xg.grad.numpy(), enc_seed=seed)
    print("pool:", area.shape, img.shape, float(mp.abs().sum()))


class FakeEnc(torch.nn.Module):
    def forward(self, d):
        return None, None, None


class FakeDec(torch.nn.Module):
    def __init__(self, emb):
        super().__init__()
        self.emb = emb

    def forward(self, final, feats, shape):
        return self.emb


def predict_case(seed=5, B=2, D=64, H=16, W=16, C=60, num_negatives=20, top_k=5):
    gen = torch.Generator().manual_seed(seed)
    text = torch.randn(C, D, generator=gen)
    seg = block_seg(B, H, W, 4, list(range(0, 10)), gen)
    emb = unit(text, 1)[seg].permute(0, 3, 1, 2) + 0.35 * torch.randn(B, D, H, W, generator=gen)
    model = bare_model()
    model.depth_encoder, model.depth_decoder = FakeEnc(), FakeDec(emb)
    random.seed(seed)
    topk, xn, temp = model.predict(torch.zeros(B, 1, H, W), text, seg, num_negatives=num_negatives, top_k=top_k)
    np.savez_compressed(os.path.join(OUT, "predict.npz"), emb=emb.numpy(), text=text.numpy(), seg=seg.numpy(),
                        seed=seed, num_negatives=num_negatives, top_k=top_k, topk=topk.numpy(),
                        xn=xn.numpy(), temperature=temp.detach().numpy())
    print("predict:", topk.shape, float((topk[:, 0] == seg).float().mean()))


def metrics_case(seed=9, C=48, B=2, H=16, W=16, k=5, n_batches=3):
    gen = torch.Generator().manual_seed(seed)
    # symmetric synonym pairs including NON-transitive chains a~b, b~c, a!~c
    eq = {i: {i} for i in range(C)}
    for a, b in [(3, 7), (7, 11), (11, 2), (20, 21), (21, 22), (5, 30), (0, 40), (33, 1)]:
        eq[a].add(b)
        eq[b].add(a)
    E = build_equivalence_tensor(eq, C)
    cmap = build_equivalence_class_map(E, torch.device("cpu"))
    batches, preds = [], []
    for b in range(n_batches):
        seg = block_seg(B, H, W, 4, list(range(0, 24)) if b < n_batches - 1 else [0, 3, 7, 11, 21, 40], gen)
        scores = torch.rand(B, C, H, W, generator=gen)
        scores.scatter_add_(1, seg[:, None], 0.55 * torch.ones(B, 1, H, W))
        topk = scores.topk(k, dim=1).indices
        batches.append({"depth": torch.zeros(B, 1, H, W), "image": torch.zeros(B, 3, H, W), "segmentation": seg,
                        "object_bbox": torch.zeros(B, 4, dtype=torch.long), "object_label": "x"})
        preds.append(topk)

    class FakeModel:
        def __init__(self):
            self.i = 0

        def eval(self):
            return self

        def predict(self, depth_maps, candidate_text_embeddings, segmentation, num_negatives, top_k):
            t = preds[self.i]
            self.i += 1
            return t, torch.zeros(B, 4, H, W), torch.tensor(0.07)

        def compute_loss(self, **kw):
            return torch.tensor(0.0), {"total_loss": 0.0}

    made = []

    class Rec(collections.defaultdict):
        def __init__(self, *a, **k):
            super().__init__(*a, **k)
            made.append(self)

    V.defaultdict = Rec
    V.log = lambda *a, **k: None
    with mock.patch("torch.cuda.empty_cache", lambda: None):
        best = V.validate_model(FakeModel(), None, None, None, [str(i) for i in range(C)], E, cmap, None,
                                dict(pct_medium=0.0, pct_hard=0.5, pct_rand=0.5), batches, 0,
                                {"step": -1, "loss": float("inf")}, torch.device("cpu"))
    names = ["intersection_top1", "union_top1", "intersection_topk", "union_topk"]
    dicts = {}
    for nm, d in zip(names, made[:4]):
        dicts[nm + "_keys"] = np.array(list(d.keys()), dtype=np.int64)      # insertion order (Q11)
        dicts[nm + "_vals"] = np.array(list(d.values()), dtype=np.int64)
    pairs = np.array([(a, b) for a in eq for b in eq[a] if a != b], dtype=np.int64)
    np.savez_compressed(os.path.join(OUT, "metrics.npz"), C=C, k=k, pairs=pairs, E=E.numpy(), cmap=cmap.numpy(),
                        seg=np.stack([b["segmentation"].numpy() for b in batches]),
                        topk=np.stack([p.numpy() for p in preds]),
                        mIoU_t1=best["mIoU_t1"], mIoU_tk=best["mIoU_tk"],
                        pixel_accuracy_t1=best["pixel_accuracy_t1"], pixel_accuracy_tk=best["pixel_accuracy_tk"],
                        **dicts)
    print("metrics:", {k_: best[k_] for k_ in ("mIoU_t1", "mIoU_tk", "pixel_accuracy_t1", "pixel_accuracy_tk")})


def infonce_numpy_case(seed=3, n=6, D=16):
    gen = torch.Generator().manual_seed(seed)
    a = unit(torch.randn(n, D, generator=gen), 1).numpy()
    b = unit(torch.randn(n, D, generator=gen), 1).numpy()
    np.savez_compressed(os.path.join(OUT, "info_nce_numpy.npz"), src=a, tgt=b, temperature=0.1,
                        loss=info_nce(a, b, temperature=0.1))


if __name__ == "__main__":
    loss_case("dict", "dict", seed=101)
    loss_case("list", "list", seed=202)                       # hard/medium silently unused
    loss_case("noimg", "dict", seed=303, n_obj=0)            # image branch: dummy * tau * 0 (Q14)
    loss_case("medium", "dict", seed=404, pcts=(0.25, 0.5, 0.25), k_distractors=20, n_obj=5)
    # D = 256 / 512: the shapes the DEFAULT (tcgen05) path of compute_loss takes; K ~ 40 contrast rows, dict-form sets
    loss_case("d256", "dict", seed=505, D=256, H=16, W=24, C=64, k_distractors=32, bf16_exact=True)
    loss_case("d512", "dict", seed=606, D=512, H=16, W=24, C=64, k_distractors=32, bf16_exact=True)
    pool_case()
    predict_case()
    metrics_case()
    infonce_numpy_case()
