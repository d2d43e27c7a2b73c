"""Drop-ins for ``masked_average_pooling`` (model.py:15-56) and ``prepare_image_contrast_data``
(dataloader.py:205-305) on the segment-masked pooling kernel (one read of the embeddings instead
of one full-tensor ``torch.where`` temporary per object)."""
from __future__ import annotations

import torch

from . import ops


def masked_average_pooling(pixel_embeddings, segmentation_map, object_indices):
    """Batch-wide mean embedding per object index (model.py:15-56); zeros for absent indices.
    Differentiable w.r.t. ``pixel_embeddings`` like the reference function."""
    B, D, H, W = pixel_embeddings.shape
    device = pixel_embeddings.device
    object_indices = torch.as_tensor(object_indices, device=device, dtype=torch.long).reshape(-1)
    n = int(object_indices.numel())
    if n == 0:
        return torch.zeros((0, D), device=device, dtype=pixel_embeddings.dtype)
    seg = segmentation_map.to(device)
    # labels -> first slot holding that label; duplicate indices are filled by a row gather below
    C = int(max(int(object_indices.max()) + 1, 1))
    uniq, inverse = torch.unique(object_indices, return_inverse=True)
    lut = torch.full((C,), -1, dtype=torch.int32, device=device)
    ok = uniq >= 0
    lut[uniq[ok]] = torch.arange(uniq.numel(), device=device, dtype=torch.int32)[ok]
    pooled = ops.masked_pool(pixel_embeddings, seg, lut, False, int(uniq.numel()))
    return pooled[inverse]


def _subgrids_2x2(seg):
    """The four full-resolution label sub-grids [B, H/2, W/2] that sit on one half-resolution embedding map."""
    return [seg[:, a::2, b::2] for a in (0, 1) for b in (0, 1)]


def pool_objects_per_image(pixel_embeddings, segmentation, image_index, labels, differentiable=False, shared2x2=False):
    """area[i] = mean of pixel_embeddings[image_index[i]] over segmentation == labels[i]
    (dataloader.py:286-304); zeros where the mask is empty.  All objects in one kernel launch.

    ``shared2x2``: ``pixel_embeddings`` is the decoder's ``output_conv`` result [B,D,H/2,W/2] and the
    segmentation is full resolution: the mean over the nearest-upsampled, normalised tensor
    (decoder.py:113-114) is the count-weighted mean of the normalised half-resolution rows -- four
    accumulating passes over the small tensor, one per label sub-grid, instead of one over the large one."""
    if shared2x2:      # decoder.py:114
        hw_lo = pixel_embeddings.shape[2] * pixel_embeddings.shape[3]
        if hw_lo % 8 == 0 and pixel_embeddings.is_cuda:
            pixel_embeddings = ops.normalize_rows(pixel_embeddings if differentiable else pixel_embeddings.detach())
        else:
            pixel_embeddings = torch.nn.functional.normalize(pixel_embeddings.float(), p=2, dim=1)
    B, D = pixel_embeddings.shape[0], pixel_embeddings.shape[1]
    device = pixel_embeddings.device
    image_index = torch.as_tensor(image_index, device=device, dtype=torch.long).reshape(-1)
    labels = torch.as_tensor(labels, device=device, dtype=torch.long).reshape(-1)
    n = int(labels.numel())
    if n == 0:
        return torch.zeros((0, D), device=device, dtype=pixel_embeddings.dtype)
    seg = segmentation
    if seg.dim() == 4:
        seg = seg.squeeze(1)
    C = int(max(int(labels.max()) + 1, 1))
    key = image_index * C + labels.clamp_min(0)
    uniq, inverse = torch.unique(key, return_inverse=True)
    lut = torch.full((B * C,), -1, dtype=torch.int32, device=device)
    lut[uniq] = torch.arange(uniq.numel(), device=device, dtype=torch.int32)
    neg = labels < 0
    x = pixel_embeddings if differentiable else pixel_embeddings.detach()
    if shared2x2:
        if tuple(seg.shape[-2:]) != (2 * x.shape[2], 2 * x.shape[3]):
            raise RuntimeError(f"pool_objects_per_image(shared2x2): segmentation must be {2 * x.shape[2]}x{2 * x.shape[3]}")
        seg = _subgrids_2x2(seg)
    pooled = ops.masked_pool(x, seg, lut.view(B, C), True, int(uniq.numel()))[inverse]
    if bool(neg.any()):
        pooled = pooled.masked_fill(neg[:, None], 0)
    return pooled


def fused_crop_config(clip_processor):
    """(shortest_edge, crop_size, mean, std) when ``clip_processor`` is a torchvision-backend CLIP-style image processor
    whose pipeline rc_clip_crops reproduces (bicubic antialiased shortest-edge resize -> square centre crop -> normalise),
    else None.  The PIL backend rounds the resized image to uint8 (image_transforms.resize), so it is never replaced."""
    p = getattr(clip_processor, "image_processor", clip_processor)       # CLIPProcessor (train_util.py:126) wraps the image processor
    try:
        if getattr(p, "backend", None) != "torchvision":
            return None
        size, crop = p.size, p.crop_size
        short = size["shortest_edge"] if isinstance(size, dict) else getattr(size, "shortest_edge", None)
        longest = size.get("longest_edge") if isinstance(size, dict) else getattr(size, "longest_edge", None)
        ch = crop["height"] if isinstance(crop, dict) else getattr(crop, "height", None)
        cw = crop["width"] if isinstance(crop, dict) else getattr(crop, "width", None)
        resample = int(getattr(p.resample, "value", p.resample))
        ok = (p.do_resize and p.do_center_crop and p.do_normalize and short and not longest and ch and ch == cw and resample == 3
              and not getattr(p, "do_pad", False))
        if not ok:
            return None
        return int(short), int(ch), [float(v) for v in p.image_mean], [float(v) for v in p.image_std]
    except Exception:  # noqa: BLE001
        return None


@torch.no_grad()
def prepare_image_contrast_data(image_processed_batch, object_bbox_batch, object_label_batch, segmentation_batch,
                                pixel_embeddings_batch, clip_image_encoder, clip_processor, device, shared2x2=False,
                                fused_crops="auto"):
    """Area embeddings + CLIP crop embeddings for the image contrastive loss (dataloader.py:205-305).
    Validation, cropping and the CLIP call are the reference's host logic; the per-object masked
    means (dataloader.py:286-304) run as one pooling kernel.  ``@torch.no_grad`` as in the
    reference: the area embeddings are detached.  ``shared2x2``: ``pixel_embeddings_batch`` is the
    decoder's half-resolution ``output_conv`` result (see ``pool_objects_per_image``)."""
    if not isinstance(image_processed_batch, torch.Tensor) or image_processed_batch.dim() != 4:
        print("Warning: 'image_processed_batch' is not a 4D tensor.")
        return None, None
    if not isinstance(object_bbox_batch, torch.Tensor) or object_bbox_batch.dim() != 2 or object_bbox_batch.shape[1] != 4:
        print("Warning: 'object_bbox_batch' is not a [B, 4] tensor.")
        return None, None
    if not isinstance(object_label_batch, torch.Tensor):
        print("Warning: 'object_label_batch' is not a tensor.")
        return None, None
    B, _, H_proc, W_proc = image_processed_batch.shape
    if B == 0:
        return None, None
    boxes = object_bbox_batch.tolist()          # one transfer instead of 4*B .item() calls
    labels = object_label_batch.tolist()
    cfg = fused_crop_config(clip_processor) if fused_crops in ("auto", True) else None
    if fused_crops is True and cfg is None:
        raise RuntimeError("prepare_image_contrast_data(fused_crops=True): the processor is not a torchvision-backend CLIP-style "
                           "pipeline (bicubic shortest-edge resize, square centre crop, normalise)")
    fused = cfg is not None and image_processed_batch.is_cuda and image_processed_batch.shape[1] == len(cfg[2])
    crops, keep, keep_labels, keep_boxes = [], [], [], []
    for b in range(B):
        xmin, ymin, xmax, ymax = boxes[b]
        if xmax > xmin and ymax > ymin and xmin >= 0 and ymin >= 0 and xmax <= W_proc and ymax <= H_proc:
            if fused:             # the crop is never sliced out: rc_clip_crops reads the box straight from the image
                if int(ymax) > int(ymin) and int(xmax) > int(xmin):
                    keep.append(b)
                    keep_labels.append(int(labels[b]))
                    keep_boxes.append([int(xmin), int(ymin), int(xmax), int(ymax)])
                else:
                    print(f"Warning: Skipping item {b}, cropped processed tensor is empty for bbox [{xmin},{ymin},{xmax},{ymax}].")
                continue
            crop = image_processed_batch[b][:, int(ymin):int(ymax), int(xmin):int(xmax)]
            if crop.numel() > 0:
                crops.append(crop)
                keep.append(b)
                keep_labels.append(int(labels[b]))
            else:
                print(f"Warning: Skipping item {b}, cropped processed tensor is empty for bbox [{xmin},{ymin},{xmax},{ymax}].")
        else:
            print(f"Debug: Skipping item {b}, invalid bbox [{xmin},{ymin},{xmax},{ymax}] relative to processed dims ({H_proc}, {W_proc}) for label {labels[b]}.")
    if not keep:
        return None, None
    if fused:
        # dataloader.py:254,276 for every object in ONE launch: box -> bicubic antialiased resize -> centre crop -> normalise
        short, csize, mean, std = cfg
        bx = torch.tensor(keep_boxes, dtype=torch.int32).to(image_processed_batch.device, non_blocking=True)
        ix = torch.tensor(keep, dtype=torch.int32).to(image_processed_batch.device, non_blocking=True)
        image_inputs = {'pixel_values': ops.clip_crops(image_processed_batch, bx, ix, short, csize, mean, std).to(device)}
    else:
        try:
            image_inputs = clip_processor(images=crops, return_tensors="pt", padding=True, do_rescale=False).to(device)
        except Exception as e:  # same contract as dataloader.py:276-278
            print(f"Error during clip_processor processing cropped tensors: {e}")
            return None, None
    try:
        image_embeddings = clip_image_encoder.get_image_features(pixel_values=image_inputs['pixel_values'])
    except Exception as e:
        raise TypeError(f"CLIP image encoder failed: {e}. Ensure it's a compatible model.") from e
    area = pool_objects_per_image(pixel_embeddings_batch, segmentation_batch, keep, keep_labels, shared2x2=shared2x2)
    return area.to(image_embeddings.dtype), image_embeddings
