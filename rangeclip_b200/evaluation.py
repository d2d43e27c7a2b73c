"""Drop-ins for the evaluation hot path: the tail of ``DepthUNet.predict`` (model.py:144-173) and
the metric accumulation / finalisation of ``validate_model`` (validate.py:88-139, 194-214).

``MetricAccumulator`` keeps everything on the device as int64 histograms; one transfer at the end
reproduces the reference's four dictionaries (including their insertion order, which fixes the
float summation order of the mIoU) and therefore its four returned floats bit for bit.
Across ranks the histograms are summed with a single all-reduce (see distributed.py)."""
from __future__ import annotations

import random
from typing import Dict, Optional

import torch
import torch.nn.functional as F

from . import ops

INT32_MAX = 2**31 - 1


def build_reduced_candidates(segmentation: torch.Tensor, total_candidates: int, num_negatives: int):
    """GT labels + ``random.sample`` negatives, sorted (model.py:147-156; Python RNG, Q6)."""
    if segmentation is None:
        raise ValueError("segmentation must be provided for reduced-candidate prediction")
    gt = set(torch.unique(segmentation).tolist())
    pool = list(set(range(total_candidates)) - gt)
    negatives = random.sample(pool, min(num_negatives, len(pool)))
    return sorted(list(gt.union(negatives)))


def predict_from_embeddings(pixel_embeddings, candidate_text_embeddings, segmentation, num_negatives=300, top_k=5,
                            precision="fp32"):
    """model.py:144-173 given the decoder output: returns (topk ids [B,k,H,W] int64 in the ORIGINAL
    index space, L2-normalised embeddings).  The [B,Kr,HW] logits are never materialised."""
    total = candidate_text_embeddings.shape[0]
    reduced = build_reduced_candidates(segmentation, total, num_negatives)
    index_tensor = torch.tensor(reduced, device=pixel_embeddings.device)
    t_norm, _, _ = ops.text_prepare(candidate_text_embeddings, index_tensor, want_f32=True)
    topk = ops.eval_topk(pixel_embeddings, t_norm, index_tensor, min(top_k, len(reduced)), precision)
    return topk, F.normalize(pixel_embeddings, dim=1)


def predict(self, depth_maps, candidate_text_embeddings, segmentation, num_negatives=300, top_k=5):
    """Drop-in for ``DepthUNet.predict`` (model.py:119-175): backbone in PyTorch, tail on the kernels."""
    self.eval()
    B, _, H, W = depth_maps.shape
    with torch.no_grad():
        with torch.autocast("cuda", enabled=depth_maps.is_cuda):
            _, encoder_features, final_feature_map = self.depth_encoder(depth_maps)
            pixel_embeddings = self.depth_decoder(final_feature_map, encoder_features, (H, W))
        topk, pixel_embeddings = predict_from_embeddings(pixel_embeddings.float(), candidate_text_embeddings,
                                                         segmentation, num_negatives, top_k)
        return topk, pixel_embeddings, self.temperature_text


class MetricAccumulator:
    """Device-side state of validate.py:59-69 / 88-139."""

    def __init__(self, equivalence_tensor: torch.Tensor, equiv_class_map: torch.Tensor, device=None):
        device = device if device is not None else equiv_class_map.device
        self.E = equivalence_tensor.to(device=device, dtype=torch.uint8).contiguous()
        self.cmap = equiv_class_map.to(device=device, dtype=torch.int64).contiguous()
        self.C = int(self.cmap.numel())
        self.acc = torch.zeros(4, self.C, device=device, dtype=torch.int64)        # I1, U1, IK, UK
        self.counters = torch.zeros(3, device=device, dtype=torch.int64)           # correct1, correctk, total
        self.first_seen = torch.full((self.C,), INT32_MAX, device=device, dtype=torch.int32)
        self.n_batches = 0
        self._hist = torch.zeros(5, self.C, device=device, dtype=torch.int64)

    def update(self, segmentation: torch.Tensor, pred_topk: torch.Tensor, batch_index: Optional[int] = None) -> None:
        """One validation batch: segmentation [B,H,W], pred_topk [B,k,H,W] (no host sync)."""
        self._hist.zero_()
        ops.eval_hist(segmentation, pred_topk, self.E, self.cmap, self._hist, self.counters)
        ops.eval_fold(self._hist, self.n_batches if batch_index is None else batch_index, self.acc, self.first_seen)
        self.n_batches += 1

    def state(self) -> Dict[str, torch.Tensor]:
        return dict(acc=self.acc, counters=self.counters, first_seen=self.first_seen)

    def finalize(self, last_segmentation: torch.Tensor) -> Dict[str, object]:
        """validate.py:194-214: mIoU over labels present in the LAST batch's GT (Q11), averaged in the
        reference's dict insertion order (batch of first appearance, then label id); accuracies."""
        acc = self.acc.cpu().tolist()
        counters = self.counters.cpu().tolist()
        first = self.first_seen.cpu().tolist()
        valid = set(self.cmap[last_segmentation.reshape(-1).to(self.cmap.device)].tolist())
        order = sorted((fs, lab) for lab, fs in enumerate(first) if fs != INT32_MAX)
        dicts = {name: {} for name in ("intersection_top1", "union_top1", "intersection_topk", "union_topk")}
        for _, lab in order:
            dicts["intersection_top1"][lab] = acc[0][lab]
            dicts["union_top1"][lab] = acc[1][lab]
            dicts["intersection_topk"][lab] = acc[2][lab]
            dicts["union_topk"][lab] = acc[3][lab]

        def miou(inter, union):
            ious = [inter[lab] / union[lab] for lab in union if lab in valid and union[lab] > 0]
            return sum(ious) / len(ious) if ious else 0.0

        total = counters[2]
        return {
            "mIoU_t1": miou(dicts["intersection_top1"], dicts["union_top1"]),
            "mIoU_tk": miou(dicts["intersection_topk"], dicts["union_topk"]),
            "pixel_accuracy_t1": counters[0] / total if total > 0 else 0.0,
            "pixel_accuracy_tk": counters[1] / total if total > 0 else 0.0,
            "correct_pixels_top1": counters[0], "correct_pixels_topk": counters[1], "total_pixels": total,
            **dicts,
        }
