"""Pin the CPU oracle (oracle/rangeclip_oracle.py) against outputs of the unmodified
reference (tests/golden/*.npz, written by tests/golden/make_golden.py)."""
import os
import random

import numpy as np
import pytest
import torch

from oracle import rangeclip_oracle as O

LOSS_CASES = ["dict", "list", "noimg", "medium"]


def _load(golden_dir, name):
    return np.load(os.path.join(golden_dir, name), allow_pickle=False)


def _sets(g):
    C = g["text"].shape[0]
    hard = {i: [int(v) for v in g["hard"][i]] for i in range(C)}
    med = {i: [int(v) for v in g["medium"][i]] for i in range(C)}
    if str(g["sim_form"]) == "list":
        return {"medium": [med[i] for i in range(C)], "hard": [hard[i] for i in range(C)]}
    return {"medium": med, "hard": hard}


@pytest.mark.parametrize("case", LOSS_CASES)
def test_compute_loss_matches_reference(golden_dir, case):
    g = _load(golden_dir, f"loss_{case}.npz")
    X = torch.tensor(g["X"], requires_grad=True)
    lt = torch.log(torch.tensor(0.07)).requires_grad_(True)
    li = torch.log(torch.tensor(0.1)).requires_grad_(True)
    area = torch.tensor(g["area"]) if g["area"].size else None
    img = torch.tensor(g["img"]) if g["img"].size else None
    seed = int(g["seed"])
    np.random.seed(seed); torch.manual_seed(seed); random.seed(seed)
    total, info, contrast = O.compute_loss(
        X, torch.tensor(g["seg"]), torch.tensor(g["text"]), _sets(g), area, img, lt, li,
        W_image=float(g["W_image"]), W_smooth=float(g["W_smooth"]),
        percent_image_sampling=float(g["pct_sampling"]), k_distractors=int(g["k_distractors"]),
        pct_medium=float(g["pcts"][0]), pct_hard=float(g["pcts"][1]), pct_rand=float(g["pcts"][2]),
        rand_indices=torch.tensor(g["rand_idx"]))
    # the contrast set consumes np.random / torch RNG streams exactly like the reference
    assert np.array_equal(contrast.numpy(), g["contrast"])
    total.backward()
    assert np.allclose(total.detach().numpy(), g["total"], rtol=1e-6, atol=0)
    assert np.isclose(info["text_contrastive_loss"], float(g["text_loss"]), rtol=1e-6)
    assert np.isclose(info["image_contrastive_loss"], float(g["image_loss"]), rtol=1e-6, atol=1e-12)
    assert np.isclose(info["smoothness_loss"], float(g["smooth_loss"]), rtol=1e-6)
    assert np.allclose(X.grad.numpy(), g["dX"], rtol=1e-5, atol=1e-9)
    assert np.allclose(lt.grad.numpy(), g["dlogtau_text"], rtol=1e-5)
    if img is not None:
        assert np.allclose(li.grad.numpy(), g["dlogtau_image"], rtol=1e-5)


@pytest.mark.parametrize("case", LOSS_CASES)
def test_dense_weighted_form_equals_sampled_reference(golden_dir, case):
    """the multiplicity-weighted dense form the CUDA kernels implement reproduces
    the reference's sampled-with-replacement loss and its gradients."""
    g = _load(golden_dir, f"loss_{case}.npz")
    X = torch.tensor(g["X"]); seg = torch.tensor(g["seg"]); text = torch.tensor(g["text"])
    B, D, H, W = X.shape
    contrast = torch.tensor(g["contrast"])
    w = O.sampling_weights(seg, torch.tensor(g["rand_idx"])).reshape(-1)
    mapping = torch.full((text.shape[0],), -1, dtype=torch.long)
    mapping[contrast] = torch.arange(len(contrast))
    y = mapping[seg.reshape(-1)]
    y = torch.where(seg.reshape(-1) > 0, y, torch.full_like(y, -1))
    rows = X.permute(0, 2, 3, 1).reshape(-1, D)
    t = torch.nn.functional.normalize(text[contrast].double(), dim=1)
    r = O.infonce_dense(rows, t, y, w, 1.0 / 0.07)
    assert np.isclose(float(r["loss"]), float(g["text_loss"]), rtol=2e-6)
    # isolate the text part of the reference gradient: subtract smoothness and check direction
    dX_text = r["dx"].reshape(B, H, W, D).permute(0, 3, 1, 2)
    dX_smooth = O.smoothness_grad(X) * float(g["W_smooth"])
    assert np.allclose((dX_text + dX_smooth).numpy(), g["dX"], rtol=2e-4, atol=2e-7)
    assert np.isclose(float(r["dlogtau"]), float(g["dlogtau_text"]), rtol=1e-5)


def test_pooling_matches_reference(golden_dir):
    g = _load(golden_dir, "pool.npz")
    X = torch.tensor(g["X"]); seg = torch.tensor(g["seg"])
    items = g["valid_items"].tolist()
    labels = [int(g["label"][i]) for i in items]
    area = O.area_pool_per_image(X, seg, items, labels)
    assert np.array_equal(area.numpy(), g["area"])       # same op order -> bit-equal
    assert np.all(g["area"][2] == 0)                     # absent label -> zeros (dataloader.py:300)
    Xg = X.clone().requires_grad_(True)
    mp = O.masked_average_pooling(Xg, seg, torch.tensor(g["objs"]))
    assert np.allclose(mp.detach().numpy(), g["mp"], rtol=1e-6, atol=1e-9)
    (mp * torch.tensor(g["mp_upstream"])).sum().backward()
    assert np.allclose(Xg.grad.numpy(), g["mp_dX"], rtol=1e-6, atol=1e-10)


def test_predict_tail_matches_reference(golden_dir):
    g = _load(golden_dir, "predict.npz")
    seg = torch.tensor(g["seg"]); text = torch.tensor(g["text"])
    random.seed(int(g["seed"]))
    reduced = O.build_candidate_set(seg, text.shape[0], int(g["num_negatives"]))
    topk, logits, xn = O.predict_tail(torch.tensor(g["emb"]), text, reduced, int(g["top_k"]))
    assert np.array_equal(topk.numpy(), g["topk"])
    assert np.allclose(xn.numpy(), g["xn"], rtol=1e-6, atol=1e-8)


def _run_metrics(g):
    st = O.MetricState()
    E, cmap = g["E"], g["cmap"]
    for seg, topk in zip(g["seg"], g["topk"]):
        k = topk.shape[1]
        O.metrics_accumulate(st, seg.reshape(-1), np.transpose(topk, (0, 2, 3, 1)).reshape(-1, k), E, cmap)
    return st, O.metrics_finalize(st, g["seg"][-1], cmap)


def test_metrics_match_reference_bit_exact(golden_dir):
    g = _load(golden_dir, "metrics.npz")
    st, fin = _run_metrics(g)
    for nm in ["intersection_top1", "union_top1", "intersection_topk", "union_topk"]:
        d = getattr(st, nm)
        assert list(d.keys()) == g[nm + "_keys"].tolist(), nm      # insertion order (Q11)
        assert list(d.values()) == g[nm + "_vals"].tolist(), nm
    for key in ["mIoU_t1", "mIoU_tk", "pixel_accuracy_t1", "pixel_accuracy_tk"]:
        assert fin[key] == float(g[key]), key                        # bit-exact floats


def test_equivalence_tables_match_reference(golden_dir):
    g = _load(golden_dir, "metrics.npz")
    C = int(g["C"])
    eq = {i: {i} for i in range(C)}
    for a, b in g["pairs"]:
        eq[int(a)].add(int(b))
    E = O.build_equivalence_tensor(eq, C)
    assert np.array_equal(E, g["E"])
    cmap = O.build_equivalence_class_map(E)
    assert np.array_equal(cmap, g["cmap"])
    # the chain 3~7~11~2 is non-transitive under the row-minimum map (Q9)
    assert cmap[3] == 3 and cmap[7] == 3 and cmap[11] == 2


def test_info_nce_numpy_known_answer(golden_dir):
    g = _load(golden_dir, "info_nce_numpy.npz")
    assert np.isclose(O.info_nce_numpy(g["src"], g["tgt"], float(g["temperature"])), float(g["loss"]), rtol=1e-7)
    li = torch.log(torch.tensor(float(g["temperature"])))
    mine = O.image_infonce(torch.tensor(g["src"]), torch.tensor(g["tgt"]), li)
    assert np.isclose(float(mine), float(g["loss"]), rtol=1e-5)


def test_segclip_restatement_agrees_when_transitive():
    """benchmark/segclip.py:60-140 agrees with validate.py only for a transitive relation."""
    rng = np.random.default_rng(0)
    C, N, k = 12, 400, 3
    eq = {i: {i} for i in range(C)}
    for a, b in [(1, 2), (5, 6)]:
        eq[a].add(b); eq[b].add(a)
    E = O.build_equivalence_tensor(eq, C); cmap = O.build_equivalence_class_map(E)
    gt = rng.integers(0, C, N)
    topk = np.stack([rng.permutation(C)[:k] for _ in range(N)])
    st = O.MetricState(); O.metrics_accumulate(st, gt, topk, E, cmap)
    fin = O.metrics_finalize(st, gt, cmap)
    a1, m1, ak, mk = O.topk_metrics_numpy(gt, topk, eq)
    assert np.isclose(a1, fin["pixel_accuracy_t1"]) and np.isclose(ak, fin["pixel_accuracy_tk"])
    assert np.isclose(m1, fin["mIoU_t1"])


# ---------------------------------------------------------------------------------------------
# shared-embedding (2x2 block) form, SURVEY 8(f)-1: fixtures from the reference decoder tail +
# compute_loss (tests/golden/make_golden_up2.py)
# ---------------------------------------------------------------------------------------------

@pytest.mark.parametrize("case", ["a", "b"])
def test_decoder_tail_and_loss_match_reference(golden_dir, case):
    g = _load(golden_dir, f"up2_{case}.npz")
    E = torch.tensor(g["E"], requires_grad=True)
    seg = torch.tensor(g["seg"])
    lt = torch.log(torch.tensor(0.07)).requires_grad_(True)
    li = torch.log(torch.tensor(0.1)).requires_grad_(True)
    C = g["text"].shape[0]
    sets = {"medium": {i: [int(v) for v in g["medium"][i]] for i in range(C)},
            "hard": {i: [int(v) for v in g["hard"][i]] for i in range(C)}}
    seed = int(g["seed"])
    np.random.seed(seed); torch.manual_seed(seed); random.seed(seed)
    X = O.decoder_tail(E, seg.shape[1:])
    total, info, contrast = O.compute_loss(X, seg, torch.tensor(g["text"]), sets, None, None, lt, li,
                                           W_smooth=float(g["W_smooth"]), percent_image_sampling=float(g["pct_sampling"]),
                                           k_distractors=int(g["k_distractors"]), rand_indices=torch.tensor(g["rand_idx"]))
    assert np.array_equal(contrast.numpy(), g["contrast"])
    total.backward()
    assert np.allclose(total.detach().numpy(), g["total"], rtol=1e-6)
    assert np.allclose(E.grad.numpy(), g["dE"], rtol=1e-5, atol=1e-9)
    assert np.allclose(lt.grad.numpy(), g["dlogtau_text"], rtol=1e-5)


@pytest.mark.parametrize("case", ["a", "b"])
def test_shared_embedding_form_equals_reference(golden_dir, case):
    """Quirk Q8: the loss on the nearest-upsampled tensor equals the four-target form on the distinct
    low-resolution rows, and the gradient w.r.t. the decoder output is that form's row gradient plus
    the smoothness term taken at low resolution (each low-res difference appears twice)."""
    g = _load(golden_dir, f"up2_{case}.npz")
    E = torch.tensor(g["E"]).double()
    seg = torch.tensor(g["seg"]); text = torch.tensor(g["text"])
    B, D, h, w_ = E.shape
    H, W = 2 * h, 2 * w_
    assert int(g["n_mixed_blocks"]) > 0          # blocks with several different targets are exercised
    contrast = torch.tensor(g["contrast"])
    wt = O.sampling_weights(seg, torch.tensor(g["rand_idx"]))
    mapping = torch.full((text.shape[0],), -1, dtype=torch.long)
    mapping[contrast] = torch.arange(len(contrast))
    y = torch.where(seg > 0, mapping[seg], torch.full_like(seg, -1))
    y4 = O.group_2x2(y).reshape(-1, 4)
    w4 = O.group_2x2(wt.reshape(B, H, W)).reshape(-1, 4)
    rows = E.permute(0, 2, 3, 1).reshape(-1, D)
    t = torch.nn.functional.normalize(text[contrast].double(), dim=1)
    r = O.infonce_dense_rep(rows, t, y4, w4, 1.0 / 0.07)
    assert np.isclose(float(r["loss"]), float(g["text_loss"]), rtol=2e-6)
    assert np.isclose(float(r["dlogtau"]), float(g["dlogtau_text"]), rtol=1e-5)
    dE = r["dx"].reshape(B, h, w_, D).permute(0, 3, 1, 2)
    Ws = float(g["W_smooth"])
    if Ws > 0:
        El = E.clone().requires_grad_(True)
        N = torch.nn.functional.normalize(El, dim=1)
        sm = 2 * (N[..., :, :-1] - N[..., :, 1:]).abs().sum() / (B * D * H * (W - 1)) \
            + 2 * (N[..., :-1, :] - N[..., 1:, :]).abs().sum() / (B * D * (H - 1) * W)
        assert np.isclose(float(sm.detach()), float(g["smooth_loss"]), rtol=1e-6)
        (Ws * sm).backward()
        dE = dE + El.grad
    assert np.allclose(dE.numpy(), g["dE"], rtol=2e-4, atol=2e-7)


def test_clip_crops_oracle_matches_the_real_clip_processor(golden_dir):
    """oracle.clip_crops (slice -> antialiased bicubic shortest-edge resize -> centre crop -> normalise, ATen's fp32 order)
    against pixel values produced by the real CLIPImageProcessor of transformers 5.5.0 (tests/golden/make_golden_crops.py;
    dataloader.py:254,276).  fp32 path: 1e-5 relative to the value range."""
    g = np.load(os.path.join(golden_dir, "crops.npz"))
    for tag in ("small", "default"):
        S, Sc = (int(v) for v in g[tag + "_cfg"])
        ref = g["pixel_values_" + tag]
        out = O.clip_crops(g["images"], g["boxes_" + tag].tolist(), g["index_" + tag].tolist(), S, Sc, g["mean"], g["std"])
        assert out.shape == ref.shape
        assert float(np.abs(out - ref).max()) <= 1e-5 * float(np.abs(ref).max())
