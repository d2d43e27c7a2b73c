"""rc_clip_crops (dataloader.py:238-282): every object crop of a batch -> CLIP pixel values in one launch,
against fixtures from the real CLIPImageProcessor (transformers 5.5.0), against the oracle on random boxes, and through the
drop-in prepare_image_contrast_data (fused crops == the processor's own torchvision pipeline on the same GPU tensors)."""
import os

import numpy as np
import pytest
import torch

from oracle import rangeclip_oracle as O

pytestmark = pytest.mark.gpu
FP32_RTOL = 1e-5          # BASELINE.json north_star: fp32 path


def dev():
    return torch.device("cuda:0")


@pytest.mark.parametrize("tag", ["small", "default"])
def test_clip_crops_vs_real_processor_fixture(golden_dir, tag):
    from rangeclip_b200 import ops
    g = np.load(os.path.join(golden_dir, "crops.npz"))
    S, Sc = (int(v) for v in g[tag + "_cfg"])
    out = ops.clip_crops(torch.from_numpy(g["images"]).to(dev()), torch.from_numpy(g["boxes_" + tag]).to(dev()),
                         torch.from_numpy(g["index_" + tag]).to(dev()), S, Sc, g["mean"].tolist(), g["std"].tolist())
    ref = g["pixel_values_" + tag]
    assert tuple(out.shape) == ref.shape
    assert float(np.abs(out.cpu().numpy() - ref).max()) <= FP32_RTOL * float(np.abs(ref).max())


def test_clip_crops_random_boxes_vs_oracle():
    """Up- and down-scaling (antialias support > 4 taps), 1-pixel-wide boxes, full-image boxes, bf16 images, an invalid box."""
    from rangeclip_b200 import ops
    rng = np.random.default_rng(5)
    B, C, H, W = 3, 3, 150, 200
    images = rng.random((B, C, H, W), dtype=np.float32)
    boxes, index = [(0, 0, W, H), (5, 7, 6, 140), (0, 149, 200, 150), (17, 3, 190, 40)], [0, 1, 2, 1]
    for _ in range(8):
        x0, y0 = int(rng.integers(0, W - 2)), int(rng.integers(0, H - 2))
        boxes.append((x0, y0, int(rng.integers(x0 + 1, W + 1)), int(rng.integers(y0 + 1, H + 1))))
        index.append(int(rng.integers(0, B)))
    mean, std = [0.48145466, 0.4578275, 0.40821073], [0.26862954, 0.26130258, 0.27577711]
    S, Sc = 24, 20
    ref = O.clip_crops(images, boxes, index, S, Sc, mean, std)
    bx = torch.tensor(boxes, dtype=torch.int32, device=dev())
    ix = torch.tensor(index, dtype=torch.int32, device=dev())
    out = ops.clip_crops(torch.from_numpy(images).to(dev()), bx, ix, S, Sc, mean, std).cpu().numpy()
    assert float(np.abs(out - ref).max()) <= FP32_RTOL * float(np.abs(ref).max())
    img_bf = torch.from_numpy(images).to(torch.bfloat16)
    ref_bf = O.clip_crops(img_bf.float().numpy(), boxes, index, S, Sc, mean, std)
    out_bf = ops.clip_crops(img_bf.to(dev()), bx, ix, S, Sc, mean, std).cpu().numpy()
    assert float(np.abs(out_bf - ref_bf).max()) <= FP32_RTOL * float(np.abs(ref_bf).max())
    bad = torch.tensor([[10, 10, 300, 50]], dtype=torch.int32, device=dev())          # outside the image: a zero crop, no OOB read
    z = ops.clip_crops(torch.from_numpy(images).to(dev()), bad, torch.zeros(1, dtype=torch.int32, device=dev()), S, Sc, mean, std)
    assert float(z.abs().sum()) == 0.0


class _Encoder:
    """Stand-in for the frozen CLIP image tower (OUT of scope): a fixed projection of the pooled pixel values."""
    def __init__(self, D):
        g = torch.Generator().manual_seed(1)
        self.w = torch.randn(3 * 8 * 8, D, generator=g).to(dev())

    def get_image_features(self, pixel_values):
        return torch.nn.functional.adaptive_avg_pool2d(pixel_values, 8).flatten(1) @ self.w


class _ClipProcessor:
    """What transformers.CLIPProcessor does with the reference's call (train_util.py:126, dataloader.py:276): `padding` goes to
    the tokenizer (no text here), everything else to the image processor.  (The real class needs tokenizer files.)"""
    def __init__(self, image_processor):
        self.image_processor = image_processor

    def __call__(self, images=None, return_tensors=None, padding=None, **kw):
        return self.image_processor(images=images, return_tensors=return_tensors, **kw)


def test_prepare_image_contrast_data_fused_crops_equal_processor_pipeline():
    """The drop-in with fused_crops=True against fused_crops=False (the processor's own torchvision pipeline on the same
    CUDA tensors): same kept objects, same area embeddings, image embeddings within the fp32 tolerance."""
    from transformers import CLIPImageProcessor
    import rangeclip_b200 as R
    from rangeclip_b200 import pooling
    proc = _ClipProcessor(CLIPImageProcessor(size={"shortest_edge": 48}, crop_size={"height": 48, "width": 48}))
    assert pooling.fused_crop_config(proc) == (48, 48, list(proc.image_processor.image_mean), list(proc.image_processor.image_std))
    g = torch.Generator().manual_seed(2)
    B, D, H, W = 6, 64, 64, 96
    image = torch.rand(B, 3, H, W, generator=g).to(dev())
    seg = torch.randint(0, 5, (B, H, W), generator=g).to(dev())
    X = torch.randn(B, D, H, W, generator=g).to(dev())
    bbox = torch.tensor([[0, 0, 96, 64], [10, 5, 40, 60], [50, 50, 40, 60], [3, 3, 90, 20], [0, 0, 97, 64], [20, 10, 80, 50]])
    label = torch.tensor([1, 2, 3, 4, 1, 2])
    enc = _Encoder(D)
    a0, i0 = R.prepare_image_contrast_data(image, bbox, label, seg, X, enc, proc, dev(), fused_crops=False)
    a1, i1 = R.prepare_image_contrast_data(image, bbox, label, seg, X, enc, proc, dev(), fused_crops=True)
    a2, i2 = R.prepare_image_contrast_data(image, bbox, label, seg, X, enc, proc, dev())          # "auto" takes the kernel
    assert a0.shape == a1.shape == (4, D) and i0.shape == i1.shape          # items 2 (inverted box) and 4 (box past the edge) are skipped
    # (the pooled means are sums of float atomics: equal up to the summation order)
    assert float((a0 - a1).abs().max()) <= 1e-5 * float(a0.abs().max()) and float((a1 - a2).abs().max()) <= 1e-5 * float(a0.abs().max())
    assert torch.equal(i1, i2)
    assert float((i0 - i1).abs().max()) <= 1e-4 * float(i0.abs().max())
