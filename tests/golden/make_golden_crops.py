"""Golden vectors for the crop front half of prepare_image_contrast_data (dataloader.py:238-282): outputs of the REAL
`CLIPImageProcessor` (transformers 5.5.0, torchvision backend, CPU) called exactly as the reference calls it
(`clip_processor(images=crops, return_tensors="pt", padding=True, do_rescale=False)`, dataloader.py:276; the reference's
`clip_processor` is a `CLIPProcessor` (train_util.py:126), which hands `padding` to its tokenizer and everything else to this
image processor -- the tokenizer needs vocabulary files that are not available offline, so the image processor is called
directly) on slices
`image[:, ymin:ymax, xmin:xmax]` (dataloader.py:254).  Two processor configurations: the default CLIP one (224 / 224) on one
crop, and a small one (40 / 32) on crops that up-scale, down-scale (antialiasing active), are wider than tall and taller
than wide.  Run in the build container:  python tests/golden/make_golden_crops.py"""
import os

import numpy as np
import torch
import transformers
from transformers import CLIPImageProcessor

assert transformers.__version__.startswith("5.5"), transformers.__version__
g = torch.Generator().manual_seed(20261018)
images = torch.rand(3, 3, 96, 128, generator=g)
small = CLIPImageProcessor(size={"shortest_edge": 40}, crop_size={"height": 32, "width": 32})
default = CLIPImageProcessor()
assert small.backend == "torchvision" and default.backend == "torchvision"
boxes_small = [(0, 0, 128, 96), (10, 5, 31, 90), (3, 40, 120, 61), (50, 20, 58, 29), (0, 0, 40, 40), (64, 1, 127, 95), (7, 9, 107, 89)]
index_small = [0, 1, 2, 0, 1, 2, 1]
boxes_default = [(20, 10, 100, 90)]
index_default = [2]


def run(proc, boxes, index):
    crops = [images[b][:, y0:y1, x0:x1] for (x0, y0, x1, y1), b in zip(boxes, index)]
    return proc(images=crops, return_tensors="pt", do_rescale=False)["pixel_values"].numpy()


out = dict(images=images.numpy(), boxes_small=np.asarray(boxes_small, dtype=np.int32), index_small=np.asarray(index_small, dtype=np.int32),
           pixel_values_small=run(small, boxes_small, index_small), small_cfg=np.asarray([40, 32], dtype=np.int32),
           boxes_default=np.asarray(boxes_default, dtype=np.int32), index_default=np.asarray(index_default, dtype=np.int32),
           pixel_values_default=run(default, boxes_default, index_default), default_cfg=np.asarray([224, 224], dtype=np.int32),
           mean=np.asarray(default.image_mean, dtype=np.float32), std=np.asarray(default.image_std, dtype=np.float32))
path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "crops.npz")
np.savez_compressed(path, **out)
print(path, {k: v.shape for k, v in out.items()}, os.path.getsize(path))
