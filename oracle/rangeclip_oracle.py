"""CPU oracle for the DepthCLIP loss / evaluation hot path.

TEST INFRASTRUCTURE ONLY.  Nothing under ``rangeclip_b200/`` may import this
module: it is the checker the CUDA path is compared against (tests/,
``__graft_entry__.smoke()``) and the CPU baseline ``bench.py`` times
(``cpu_baseline`` / ``--impl reference``).  It is never shipped as, or routed
to from, the product path.

The reference (jinryan/RangeCLIP) is pure Python on stock PyTorch ops, so the
oracle is a restatement of the same arithmetic on CPU tensors (torch for the
floating-point terms, numpy for the integer metric accumulators).  Every
function cites the reference lines it follows (paths relative to the
reference root).  Parity of this restatement with the reference itself is
pinned by ``tests/golden/*.npz`` -- outputs of the unmodified reference
imported in the build container by ``tests/golden/make_golden.py`` -- and
checked in ``tests/test_oracle_golden.py``.
"""
from __future__ import annotations

import random as _pyrandom
from collections import defaultdict
from typing import Dict, List, Optional, Sequence, Tuple

import numpy as np
import torch
import torch.nn.functional as F

# ----------------------------------------------------------------------------
# (a1) pixel sampling  --  RangeCLIP/src/depth_segmentation_model/model.py:204-228
# ----------------------------------------------------------------------------

def num_text_samples(H: int, W: int, percent_image_sampling: float) -> int:
    """model.py:207-218 -- how many pixel indices are drawn per image."""
    hw = H * W
    if hw == 0:
        raise RuntimeError("Input dimensions H or W are zero.")  # model.py:214-215
    n = min(int(percent_image_sampling * H * W), hw)
    if n == 0 and hw > 0:
        n = hw
    return n


def sample_pixels(pixel_embeddings: torch.Tensor, target_indices: torch.Tensor,
                  rand_indices: torch.Tensor) -> Tuple[torch.Tensor, torch.Tensor]:
    """model.py:210-228 -- gather the sampled pixels (with replacement) and drop label 0."""
    B, D, H, W = pixel_embeddings.shape
    pred_flat = pixel_embeddings.reshape(B, D, -1)
    target_flat = target_indices.reshape(B, -1)
    pred = torch.gather(pred_flat, 2, rand_indices.unsqueeze(1).expand(-1, D, -1))
    lab = torch.gather(target_flat, 1, rand_indices)
    keep = lab > 0
    pred = pred.permute(0, 2, 1)[keep].reshape(-1, D)
    lab = lab[keep].reshape(-1)
    return pred, lab


def sampling_weights(target_indices: torch.Tensor, rand_indices: torch.Tensor) -> torch.Tensor:
    """Dense equivalent of model.py:220-228: multiplicity of each pixel among the sampled
    indices, zeroed on label 0.  Returns float32 [B, H*W]."""
    B = target_indices.shape[0]
    hw = target_indices[0].numel()
    w = torch.zeros(B, hw, dtype=torch.float32)
    w.scatter_add_(1, rand_indices.cpu(), torch.ones(rand_indices.shape, dtype=torch.float32))
    w = w * (target_indices.reshape(B, -1).cpu() > 0).to(torch.float32)
    return w


# ----------------------------------------------------------------------------
# (a2) contrast-set builder  --  model.py:231-270
# ----------------------------------------------------------------------------

def build_contrast_set(unique_labels: torch.Tensor, C: int, label_similarity_sets,
                       k_distractors: int, pct_medium: float, pct_hard: float,
                       pct_rand: float) -> torch.Tensor:
    """model.py:234-268.  Consumes ``np.random.choice`` (only when enough hard/medium
    candidates exist) and CPU ``torch.randperm`` exactly as the reference does;
    ``label in container`` keeps the list-vs-dict quirk."""
    assert abs(pct_medium + pct_hard + pct_rand - 1.0) < 1e-4, "Sum of text percentages must be 1."
    uniq = unique_labels.tolist()
    pool = set()
    n_medium = int(k_distractors * pct_medium)
    n_hard = int(k_distractors * pct_hard)
    n_rand = k_distractors - n_medium - n_hard
    if n_medium > 0:
        for lab in uniq:
            if lab in label_similarity_sets['medium']:
                pool.update(label_similarity_sets['medium'][lab])
    if n_hard > 0:
        for lab in uniq:
            if lab in label_similarity_sets['hard']:
                pool.update(label_similarity_sets['hard'][lab])
    pool = [d for d in list(pool) if d not in uniq]
    want = n_medium + n_hard
    picked = np.random.choice(pool, size=want, replace=False) if len(pool) >= want else pool
    picked = torch.tensor(picked, dtype=torch.long)
    all_idx = torch.arange(C)
    remaining = all_idx[~torch.isin(all_idx, torch.cat([unique_labels.cpu(), picked]))]
    if n_rand > 0 and len(remaining) > 0:
        rand_d = remaining[torch.randperm(len(remaining))[:n_rand]]
    else:
        rand_d = torch.tensor([], dtype=torch.long)
    return torch.unique(torch.cat([unique_labels.cpu(), picked, rand_d]))


# ----------------------------------------------------------------------------
# (f2) device-side contrast-set builder  --  the SET semantics of model.py:234-268 with the
# counter-based draws of rangeclip_b200/csrc/contrast.cu (the reference's NumPy / randperm
# streams cannot be reproduced on a GPU; SURVEY 8f-2: "parity then only statistical").
# Checker for rc_contrast_build: bit-exact for a given seed.
# ----------------------------------------------------------------------------

_M64 = (1 << 64) - 1


def _splitmix64(x: int) -> int:
    x = (x + 0x9E3779B97F4A7C15) & _M64
    x = ((x ^ (x >> 30)) * 0xBF58476D1CE4E5B9) & _M64
    x = ((x ^ (x >> 27)) * 0x94D049BB133111EB) & _M64
    return x ^ (x >> 31)


def contrast_draw_key(seed: int, phase: int, c: int) -> int:
    return _splitmix64((_splitmix64((seed ^ (phase << 56)) & _M64) + c) & _M64)


def contrast_build_device(counts, sim_off, sim_items, n_curriculum: int, n_rand: int, k_cap: int, seed: int,
                          include_label0: bool = False):
    """(label_map [C], contrast [k_cap] padded with -1, (K, flags, n_present, n_distractors)).
    present = labels >= 1 with a non-zero count (model.py:226,233); candidates = similarity lists of the present labels
    minus present (model.py:240-252); n_curriculum of them (all if fewer, model.py:254-259) and n_rand of the labels that
    are neither present nor chosen (model.py:261-266) are the ones with the smallest draw keys (ties of the top 44 key
    bits by label order); the contrast set is the sorted union (model.py:268), capped at k_cap rows."""
    counts = np.asarray(counts)
    C = counts.shape[0]
    present = [c for c in range(0 if include_label0 else 1, C) if counts[c] > 0]      # predict (model.py:147) keeps label 0
    pset = set(present)
    cand = set()
    if sim_off is not None and n_curriculum > 0:
        for c in present:
            for j in range(int(sim_off[c]), int(sim_off[c + 1])):
                d = int(sim_items[j])
                if 0 <= d < C and d not in pset:
                    cand.add(d)
    flags = 0
    keep_present = len(present)
    if keep_present > k_cap:
        keep_present, flags = k_cap, flags | 1
    take_cur = n_curriculum if len(cand) >= n_curriculum else len(cand)
    room = k_cap - keep_present
    if take_cur > room:
        take_cur, flags = room, flags | 2
    room -= take_cur
    n_free = C - len(present) - take_cur
    take_rand = max(0, min(n_rand, n_free))
    if take_rand > room:
        take_rand, flags = room, flags | 2

    def smallest(elig, n, phase):
        if n <= 0:
            return []
        if n >= len(elig):
            return list(elig)
        return sorted(elig, key=lambda c: (contrast_draw_key(seed, phase, c) >> 20, c))[:n]

    chosen = smallest(sorted(cand), take_cur, 1)
    cset = set(chosen)
    rand = smallest([c for c in range(C) if c not in pset and c not in cset], take_rand, 2)
    members = sorted(pset | cset | set(rand))
    label_map = np.full(C, -1, dtype=np.int32)
    contrast = np.full(k_cap, -1, dtype=np.int64)
    for r, c in enumerate(members):
        if r < k_cap:
            label_map[c] = r
            contrast[r] = c
    return label_map, contrast, (min(len(members), k_cap), flags, len(present), take_cur + take_rand)


# ----------------------------------------------------------------------------
# (f3) object crops for the CLIP image encoder  --  dataloader.py:254 (slice) + :276
# `clip_processor(images=crops, return_tensors="pt", padding=True, do_rescale=False)`.
# The processor is a THIRD-PARTY dependency (transformers 5.5.0, torchvision backend:
# image_processing_backends.py TorchvisionBackend._preprocess; tvF.resize -> ATen
# upsample_bicubic2d_aa).  Restated here from its published algorithm and pinned by
# tests/golden/crops.npz, which holds outputs of the real CLIPImageProcessor
# (tests/golden/make_golden_crops.py).
# ----------------------------------------------------------------------------

def _cubic_aa(x: np.float32) -> np.float32:
    """ATen aa bicubic filter, a = -0.5 (UpSampleKernel.cpp: HelperInterpCubic::aa_filter)."""
    f = np.float32
    a, x = f(-0.5), np.abs(f(x))
    if x < f(1.0):
        return ((a + f(2.0)) * x - (a + f(3.0))) * x * x + f(1.0)
    if x < f(2.0):
        return (((x - f(5.0)) * x + f(8.0)) * x - f(4.0)) * a
    return f(0.0)


def _aa_axis(in_size: int, out_size: int):
    """Per output index (lo, weights) of one axis: fp32 arithmetic in ATen's order (_compute_indices_weights_aa,
    align_corners=False): scale = in/out, support = 2*max(scale,1), centre = scale*(i+0.5), weights / their sum."""
    f = np.float32
    scale = f(in_size) / f(out_size)
    support = f(2.0) * scale if scale >= f(1.0) else f(2.0)
    invscale = f(1.0) / scale if scale >= f(1.0) else f(1.0)
    out = []
    for i in range(out_size):
        center = scale * (f(i) + f(0.5))
        lo = max(int(center - support + f(0.5)), 0)
        n = min(int(center + support + f(0.5)), in_size) - lo
        ws = np.array([_cubic_aa((f(j + lo) - center + f(0.5)) * invscale) for j in range(n)], dtype=np.float32)
        tot = f(0.0)
        for w in ws:
            tot = f(tot + w)
        out.append((lo, ws / tot))
    return out


def clip_crops(images: np.ndarray, boxes, image_index, shortest_edge: int, crop_size: int, mean, std) -> np.ndarray:
    """[n, C, crop, crop] f32 CLIP pixel values of the boxes (xmin, ymin, xmax, ymax) of images [B, C, H, W]:
    slice (dataloader.py:254) -> resize shortest edge to `shortest_edge`, longest to int(S*long/short), bicubic antialiased
    (horizontal pass, then vertical) -> centre crop at int((r - crop) / 2.0) -> (v - mean) / std."""
    images = np.asarray(images, dtype=np.float32)
    out = []
    for (x0, y0, x1, y1), b in zip(boxes, image_index):
        crop = images[b][:, y0:y1, x0:x1]
        C, h, w = crop.shape
        if w <= h:
            rw, rh = shortest_edge, int(shortest_edge * h / w)
        else:
            rh, rw = shortest_edge, int(shortest_edge * w / h)
        ax, ay = _aa_axis(w, rw), _aa_axis(h, rh)
        top, left = int((rh - crop_size) / 2.0), int((rw - crop_size) / 2.0)
        res = np.zeros((C, crop_size, crop_size), dtype=np.float32)
        for oy in range(crop_size):
            ylo, wy = ay[top + oy]
            rows = crop[:, ylo:ylo + len(wy), :]                                   # [C, ny, w]
            for ox in range(crop_size):
                xlo, wx = ax[left + ox]
                hsum = (rows[:, :, xlo:xlo + len(wx)] * wx[None, None, :]).sum(axis=2, dtype=np.float32)
                res[:, oy, ox] = (hsum * wy[None, :]).sum(axis=1, dtype=np.float32)
        m = np.asarray(mean, dtype=np.float32)[:, None, None]
        sd = np.asarray(std, dtype=np.float32)[:, None, None]
        out.append((res - m) / sd)
    return np.stack(out).astype(np.float32)


# ----------------------------------------------------------------------------
# (a3) logits + cross-entropy  --  model.py:272-291
# ----------------------------------------------------------------------------

def text_infonce_sampled(pixel_embeddings, target_indices, candidate_text_embeddings,
                         rand_indices, contrast_indices, log_temperature):
    """model.py:222-228,272-291 on given sample indices / contrast set (autograd-capable)."""
    C = candidate_text_embeddings.shape[0]
    pred, lab = sample_pixels(pixel_embeddings, target_indices, rand_indices)
    if pred.numel() == 0 or len(contrast_indices) <= 1:
        return torch.tensor(0.0)                                    # model.py:295-301
    t = F.normalize(candidate_text_embeddings[contrast_indices], dim=1)
    x = F.normalize(pred, dim=1)
    mapping = torch.full((C,), -1, dtype=torch.long)
    mapping[contrast_indices] = torch.arange(contrast_indices.shape[0])
    y = mapping[lab]
    ok = y != -1
    if not bool(ok.all()):                                          # model.py:280-284
        y, x = y[ok], x[ok]
    if len(y) == 0:
        return torch.tensor(0.0)
    logits = x @ t.T
    logits = logits / torch.exp(log_temperature)
    return F.cross_entropy(logits, y)


def infonce_dense(x_rows: torch.Tensor, t_norm: torch.Tensor, y: torch.Tensor, w: torch.Tensor,
                  inv_tau: float, dtype=torch.float64):
    """Closed-form weighted InfoNCE over rows (the form the CUDA kernels implement; SURVEY
    Q1/Q2): loss = sum_p w_p (lse_p - z_{p,y_p}) / sum_p w_p with z = normalize(x) t^T / tau.

    x_rows [M, D] raw (un-normalised) rows, t_norm [K, D] already L2-normalised text rows,
    y [M] int (-1 = ignore), w [M] weights.  Returns dict(loss, lse, dx, dt, dlogtau) where
    the gradients are for d loss (upstream = 1); dt is w.r.t. the *normalised* text rows.
    Follows model.py:273 (normalize, eps 1e-12), :289-291 (matmul, /tau, mean CE)."""
    x = x_rows.to(dtype)
    t = t_norm.to(dtype)
    w = w.to(dtype) * (y >= 0).to(dtype)
    yy = y.clamp(min=0).long()
    nrm = x.norm(dim=1, keepdim=True).clamp_min(1e-12)
    xh = x / nrm
    z = (xh @ t.T) * inv_tau
    lse = torch.logsumexp(z, dim=1)
    wsum = w.sum()
    zy = z.gather(1, yy[:, None])[:, 0]
    if float(wsum) == 0.0:
        zero = torch.zeros((), dtype=dtype)
        return dict(loss=zero, lse=lse, dx=torch.zeros_like(x), dt=torch.zeros_like(t),
                    dlogtau=zero, wsum=wsum)
    loss = (w * (lse - zy)).sum() / wsum
    p = torch.softmax(z, dim=1)
    p[torch.arange(len(yy)), yy] -= 1.0
    dz = p * (w / wsum)[:, None]
    dxh = (dz @ t) * inv_tau
    dx = (dxh - xh * (xh * dxh).sum(1, keepdim=True)) / nrm
    dt = (dz.T @ xh) * inv_tau
    dlogtau = -(dz * z).sum()
    return dict(loss=loss, lse=lse, dx=dx, dt=dt, dlogtau=dlogtau, wsum=wsum)


# ----------------------------------------------------------------------------
# (a5) area pooling  --  dataloader.py:286-304 and model.py:15-56
# ----------------------------------------------------------------------------

def area_pool_per_image(pixel_embeddings, segmentation, image_index: Sequence[int],
                        labels: Sequence[int]) -> torch.Tensor:
    """dataloader.py:287-304 -- mean embedding over seg[b_i] == l_i, zeros when empty."""
    D = pixel_embeddings.shape[1]
    out = torch.zeros(len(labels), D, dtype=pixel_embeddings.dtype)
    for i, (b, lab) in enumerate(zip(image_index, labels)):
        emb = pixel_embeddings[b]
        seg = segmentation[b]
        if seg.dim() == 3:
            seg = seg.squeeze(0)
        mask = (seg == lab).unsqueeze(0)
        cnt = mask.sum()
        if cnt > 0:
            out[i] = torch.where(mask.expand_as(emb), emb, torch.zeros_like(emb)).sum(dim=(1, 2)) / cnt
    return out


def masked_average_pooling(pixel_embeddings, segmentation_map, object_indices) -> torch.Tensor:
    """model.py:15-56 -- batch-wide mean per label (differentiable variant)."""
    B, D, H, W = pixel_embeddings.shape
    out = torch.zeros((len(object_indices), D), dtype=pixel_embeddings.dtype)
    seg = segmentation_map.unsqueeze(1)
    rows = []
    for i, obj in enumerate(object_indices):
        mask = seg == obj
        if bool(mask.any()):
            s = torch.where(mask.expand_as(pixel_embeddings), pixel_embeddings,
                            torch.zeros_like(pixel_embeddings)).sum(dim=(0, 2, 3))
            rows.append(s / mask.sum())
        else:
            rows.append(out[i])
    return torch.stack(rows) if rows else out


# ----------------------------------------------------------------------------
# (a6) image InfoNCE  --  model.py:304-326
# ----------------------------------------------------------------------------

def image_infonce(area_embeddings, image_embeddings, log_temperature):
    """model.py:307-321 (n > 1 branch)."""
    n = area_embeddings.shape[0]
    a = F.normalize(area_embeddings, dim=1)
    g = F.normalize(image_embeddings, dim=1)
    logits = (a @ g.T) / torch.exp(log_temperature)
    return F.cross_entropy(logits, torch.arange(n))


def info_nce_numpy(src: np.ndarray, tgt: np.ndarray, temperature: float = 0.1) -> float:
    """utils/src/eval_utils.py:3-31 -- the reference's own numpy restatement (diagonal
    positives); used as an independent known-answer cross-check of image_infonce."""
    sim = np.dot(src, tgt.T) / temperature
    e = np.exp(sim)
    p = e / e.sum(axis=1, keepdims=True)
    return float(np.mean(-np.log(np.diag(p))))


# ----------------------------------------------------------------------------
# (a7) smoothness  --  model.py:329-334
# ----------------------------------------------------------------------------

def smoothness(pixel_embeddings):
    """model.py:332-334 -- two L1 means with different denominators."""
    x = pixel_embeddings
    tv_h = F.l1_loss(x[:, :, :, :-1], x[:, :, :, 1:])
    tv_v = F.l1_loss(x[:, :, :-1, :], x[:, :, 1:, :])
    return tv_h + tv_v


def smoothness_grad(pixel_embeddings: torch.Tensor) -> torch.Tensor:
    """Closed-form d smoothness / dX with sign(0) = 0, fp64."""
    x = pixel_embeddings.to(torch.float64)
    B, D, H, W = x.shape
    g = torch.zeros_like(x)
    if W > 1:
        s = torch.sign(x[..., :, :-1] - x[..., :, 1:]) / (B * D * H * (W - 1))
        g[..., :, :-1] += s
        g[..., :, 1:] -= s
    if H > 1:
        s = torch.sign(x[..., :-1, :] - x[..., 1:, :]) / (B * D * (H - 1) * W)
        g[..., :-1, :] += s
        g[..., 1:, :] -= s
    return g


# ----------------------------------------------------------------------------
# (a1-a8) full hybrid loss  --  model.py:178-355
# ----------------------------------------------------------------------------

def compute_loss(pixel_embeddings, target_indices, candidate_text_embeddings, label_similarity_sets,
                 area_embeddings, image_embeddings, log_temperature_text, log_temperature_image,
                 W_text=1.0, W_image=0.5, W_smooth=2e2, percent_image_sampling=0.7,
                 k_distractors=50, pct_medium=0.0, pct_hard=0.75, pct_rand=0.25,
                 rand_indices: Optional[torch.Tensor] = None):
    """Restatement of DepthUNet.compute_loss (model.py:178-355) on CPU tensors.  RNG
    streams are consumed in the reference's order: torch.randint (unless ``rand_indices``
    is injected), np.random.choice, torch.randperm."""
    text_loss = torch.tensor(0.0)
    contrast = None
    if W_text > 0:
        B, D, H, W = pixel_embeddings.shape
        C = candidate_text_embeddings.shape[0]
        n = num_text_samples(H, W, percent_image_sampling)
        if rand_indices is None:
            rand_indices = torch.randint(0, H * W, (B, n))
        pred, lab = sample_pixels(pixel_embeddings, target_indices, rand_indices)
        if pred.numel() > 0 and lab.numel() > 0:
            contrast = build_contrast_set(torch.unique(lab), C, label_similarity_sets,
                                          k_distractors, pct_medium, pct_hard, pct_rand)
            if len(contrast) > 1:
                text_loss = text_infonce_sampled(pixel_embeddings, target_indices,
                                                 candidate_text_embeddings, rand_indices,
                                                 contrast, log_temperature_text)
    image_loss = torch.tensor(0.0)
    if area_embeddings is not None and image_embeddings is not None and area_embeddings.shape[0] > 1:
        image_loss = image_infonce(area_embeddings, image_embeddings, log_temperature_image)
    elif W_image > 0:
        image_loss = torch.tensor(1.0) * torch.exp(log_temperature_image) * 0.0   # model.py:325-326
    smooth = torch.tensor(0.0)
    if W_smooth > 0:
        smooth = smoothness(pixel_embeddings)
    total = W_text * text_loss + W_image * image_loss + W_smooth * smooth
    info = {
        'total_loss': float(total.detach()),
        'text_contrastive_loss': float(text_loss.detach()) if W_text > 0 else 0,
        'image_contrastive_loss': float(image_loss.detach()) if W_image > 0 else 0,
        'smoothness_loss': float(smooth.detach()) if W_smooth > 0 else 0,
        'temperature_text': float(torch.exp(log_temperature_text.detach())),
        'temperature_image': float(torch.exp(log_temperature_image.detach())),
        'W_text': W_text, 'W_image': W_image, 'W_smooth': W_smooth,
    }
    return total, info, contrast


# ----------------------------------------------------------------------------
# (f1) decoder tail + shared-embedding form  --  utils/src/decoder.py:112-116 (quirk Q8)
# ----------------------------------------------------------------------------

def decoder_tail(output_conv_result: torch.Tensor, target_shape) -> torch.Tensor:
    """decoder.py:113-114: nearest interpolation to the target shape, then L2 normalisation over
    the channels.  With target = 2x the input every 2x2 block shares one embedding."""
    out = F.interpolate(output_conv_result, size=target_shape, mode='nearest')
    return F.normalize(out, p=2, dim=1)


def group_2x2(t: torch.Tensor) -> torch.Tensor:
    """[B, 2h, 2w] -> [B, h*w, 4]: the four full-resolution entries of every embedding block."""
    B, H, W = t.shape
    return t.reshape(B, H // 2, 2, W // 2, 2).permute(0, 1, 3, 2, 4).reshape(B, (H // 2) * (W // 2), 4)


def infonce_dense_rep(x_rows: torch.Tensor, t_norm: torch.Tensor, y: torch.Tensor, w: torch.Tensor,
                      inv_tau: float, dtype=torch.float64):
    """Weighted InfoNCE where row q carries R targets (y, w are [M, R]): the reference loss
    (model.py:272-291) on the nearest-upsampled tensor, written on the distinct rows:
    loss = sum_q sum_j w_qj (lse_q - z[q, y_qj]) / sum w; dx = gradient w.r.t. the shared row."""
    x = x_rows.to(dtype)
    t = t_norm.to(dtype)
    w = w.to(dtype) * (y >= 0).to(dtype)
    yy = y.clamp(min=0).long()
    nrm = x.norm(dim=1, keepdim=True).clamp_min(1e-12)
    xh = x / nrm
    z = (xh @ t.T) * inv_tau
    lse = torch.logsumexp(z, dim=1)
    wsum = w.sum()
    wrow = w.sum(1)
    loss = ((wrow * lse).sum() - (w * z.gather(1, yy)).sum()) / wsum
    dz = torch.softmax(z, dim=1) * wrow[:, None]
    dz.scatter_add_(1, yy, -w)
    dz = dz / wsum
    dxh = (dz @ t) * inv_tau
    dx = (dxh - xh * (xh * dxh).sum(1, keepdim=True)) / nrm
    dt = (dz.T @ xh) * inv_tau
    dlogtau = -(dz * z).sum()
    return dict(loss=loss, lse=lse, dx=dx, dt=dt, dlogtau=dlogtau, wsum=wsum)


# ----------------------------------------------------------------------------
# (a9) predict tail  --  model.py:144-173
# ----------------------------------------------------------------------------

def build_candidate_set(segmentation: torch.Tensor, total_candidates: int, num_negatives: int) -> List[int]:
    """model.py:147-156 -- GT labels plus ``random.sample`` negatives, sorted."""
    gt = set(torch.unique(segmentation).tolist())
    pool = list(set(range(total_candidates)) - gt)
    neg = _pyrandom.sample(pool, min(num_negatives, len(pool)))
    return sorted(list(gt.union(neg)))


def predict_tail(pixel_embeddings, candidate_text_embeddings, reduced_indices, top_k=5):
    """model.py:144,159-173 -- normalise, logits against the reduced set, top-k, map to global
    ids.  Returns (topk [B,k,H,W] int64, logits [B,Kr,HW], normalised embeddings)."""
    B, D, H, W = pixel_embeddings.shape
    x = F.normalize(pixel_embeddings, dim=1)
    idx = torch.as_tensor(reduced_indices, dtype=torch.long)
    t = F.normalize(candidate_text_embeddings[idx], dim=1)
    logits = torch.einsum('bdn,cd->bcn', x.view(B, D, H * W), t)
    k = min(top_k, logits.shape[1])
    red = logits.topk(k, dim=1).indices.view(B, k, H, W)
    return idx[red], logits, x


# ----------------------------------------------------------------------------
# (a10-a11) metric accumulation / finalisation  --  validate.py:88-139, 194-214
# ----------------------------------------------------------------------------

class MetricState:
    """The accumulators validate.py:59-69 creates."""

    def __init__(self):
        self.intersection_top1: Dict[int, int] = defaultdict(int)
        self.union_top1: Dict[int, int] = defaultdict(int)
        self.intersection_topk: Dict[int, int] = defaultdict(int)
        self.union_topk: Dict[int, int] = defaultdict(int)
        self.correct_top1 = 0
        self.correct_topk = 0
        self.total = 0


def metrics_accumulate(state: MetricState, gt: np.ndarray, topk: np.ndarray,
                       E: np.ndarray, cmap: np.ndarray) -> None:
    """validate.py:88-139 for one batch, in numpy integer arithmetic.
    gt [N] int64, topk [N,k] int64 (column 0 = top-1), E [C,C] bool, cmap [C] int64.
    Keeps the reference quirks: oracle_pred starts from RAW top-1 ids (Q9); per-batch label
    set = unique(gt_equiv U pred_equiv_top1) (Q10); label 0 counts (Q12)."""
    gt = np.asarray(gt, dtype=np.int64).reshape(-1)
    topk = np.asarray(topk, dtype=np.int64).reshape(len(gt), -1)
    top1 = topk[:, 0]
    state.correct_top1 += int(E[gt, top1].sum())                       # validate.py:96-97
    state.total += int(gt.size)                                         # :98
    state.correct_topk += int(E[gt[:, None], topk].any(axis=1).sum())   # :101-103
    ge = cmap[gt]                                                       # :106
    p1 = cmap[top1]                                                     # :107
    labels = np.unique(np.concatenate([ge, p1]))                        # :108 (sorted)
    for lab in labels.tolist():                                         # :110-115
        pm, gm = p1 == lab, ge == lab
        state.intersection_top1[lab] += int(np.logical_and(pm, gm).sum())
        state.union_top1[lab] += int(np.logical_or(pm, gm).sum())
    tke = cmap[topk]                                                    # :119
    oracle_pred = top1.copy()                                           # :122
    for lab in labels.tolist():                                         # :123-131
        hit = (ge == lab) & (tke == lab).any(axis=1)
        oracle_pred[hit] = lab
    for lab in labels.tolist():                                         # :134-139
        pm, gm = oracle_pred == lab, ge == lab
        state.intersection_topk[lab] += int(np.logical_and(pm, gm).sum())
        state.union_topk[lab] += int(np.logical_or(pm, gm).sum())


def metrics_finalize(state: MetricState, last_gt: np.ndarray, cmap: np.ndarray) -> Dict[str, float]:
    """validate.py:194-214 -- mIoU over labels present in the LAST batch's GT, averaged in
    dict insertion order (Q11); accuracies."""
    valid = set(cmap[np.asarray(last_gt, dtype=np.int64).reshape(-1)].tolist())

    def miou(inter, union):
        ious = [inter[lab] / union[lab] for lab in union if lab in valid and union[lab] > 0]
        return sum(ious) / len(ious) if ious else 0.0

    tot = state.total
    return {
        'mIoU_t1': miou(state.intersection_top1, state.union_top1),
        'mIoU_tk': miou(state.intersection_topk, state.union_topk),
        'pixel_accuracy_t1': state.correct_top1 / tot if tot > 0 else 0.0,
        'pixel_accuracy_tk': state.correct_topk / tot if tot > 0 else 0.0,
    }


def topk_metrics_numpy(gt_flat, topk_flat, equivalence_dict):
    """benchmark/segclip.py:78-138 -- the reference's second, independent numpy statement of
    the metrics for ONE sample (class-mapped oracle_pred, no last-batch filter).  Agrees with
    metrics_accumulate/finalize only when the equivalence relation is transitive and every
    label maps to its class minimum; used as a cross-check under exactly that condition."""
    gt_flat = np.asarray(gt_flat).reshape(-1)
    topk_flat = np.asarray(topk_flat).reshape(len(gt_flat), -1)
    top1 = topk_flat[:, 0]
    eq = lambda g: equivalence_dict.get(int(g), {int(g)})
    c1 = np.array([int(p) in eq(g) for p, g in zip(top1, gt_flat)])
    ck = np.array([any(int(p) in eq(g) for p in row) for row, g in zip(topk_flat, gt_flat)])
    ge = np.array([min(eq(g)) for g in gt_flat])
    p1 = np.array([min(eq(p)) for p in top1])

    def miou(pred):
        ious = []
        for lab in np.unique(np.concatenate([ge, pred])):
            u = np.logical_or(ge == lab, pred == lab).sum()
            if u > 0:
                ious.append(np.logical_and(ge == lab, pred == lab).sum() / u)
        return float(np.mean(ious)) if ious else 0.0

    tke = np.array([[min(eq(p)) for p in row] for row in topk_flat])
    oracle_pred = p1.copy()
    for i, (g, row) in enumerate(zip(ge, tke)):
        if g in row:
            oracle_pred[i] = g
    return float(c1.mean()), miou(p1), float(ck.mean()), miou(oracle_pred)


# ----------------------------------------------------------------------------
# (a13) equivalence tables  --  dataloader.py:159-165, 191-202
# ----------------------------------------------------------------------------

def build_equivalence_tensor(equivalence_dict: Dict[int, set], num_classes: int) -> np.ndarray:
    """dataloader.py:159-165."""
    E = np.zeros((num_classes, num_classes), dtype=bool)
    for g, eqs in equivalence_dict.items():
        for p in eqs:
            E[g, p] = True
    return E


def build_equivalence_class_map(E: np.ndarray) -> np.ndarray:
    """dataloader.py:191-202 -- row-minimum representative (NOT transitive, Q9)."""
    C = E.shape[0]
    cmap = np.arange(C, dtype=np.int64)
    for i in range(C):
        nz = np.nonzero(E[i])[0]
        if len(nz) > 0:
            cmap[i] = nz.min()
    return cmap
