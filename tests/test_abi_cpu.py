"""CPU-side checks: the C-ABI library builds, loads and exports every symbol the header declares;
host-side logic (contrast-set builder, candidate-set builder) matches the reference's golden
outputs; ops refuse CPU tensors instead of silently falling back."""
import ctypes
import os
import random
import re

import numpy as np
import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_library_exports_every_declared_symbol():
    from rangeclip_b200 import _lib
    if not os.path.exists(_lib.LIB_PATH):
        _lib.build()
    hdr = open(os.path.join(ROOT, "include", "rangeclip_b200.h")).read()
    shipped, bringup = hdr.split("#ifdef RC_BRINGUP")[0], hdr.split("#ifdef RC_BRINGUP")[1].split("#endif /* RC_BRINGUP */")[0]
    declared = set(re.findall(r"\b(rc_[a-z0-9_]+)\s*\(", shipped))
    declared_bringup = set(re.findall(r"\b(rc_[a-z0-9_]+)\s*\(", bringup))
    assert declared == set(_lib.PROTOTYPES)
    assert declared_bringup == set(_lib.BRINGUP_PROTOTYPES)
    lib = ctypes.CDLL(_lib.LIB_PATH)
    for name in declared:
        assert hasattr(lib, name), name
    for name in declared_bringup:                   # the shipped library carries no bring-up entry point ...
        assert not hasattr(lib, name), name
    blib = ctypes.CDLL(_lib.BRINGUP_LIB_PATH)       # ... the bring-up build of the same sources carries both sets
    for name in declared | declared_bringup:
        assert hasattr(blib, name), name
    import subprocess
    strings = subprocess.run(["strings", _lib.LIB_PATH], capture_output=True, text=True).stdout
    assert "RANGECLIP_B200_" not in strings        # no environment switch is compiled into the shipped library
    assert _lib.lib().rc_abi_version() == 1
    assert _lib.launch_count() >= 0


def test_argument_validation_without_gpu():
    from rangeclip_b200 import _lib
    L = _lib.lib()
    assert L.rc_eval_hist(None, None, 1, 1, 1, None, None, 1, None, None, None) == -1
    assert b"null pointer" in L.rc_last_error()
    assert L.rc_infonce_workspace_bytes(2, 512, 1024, 256, _lib.RC_F32) > 2 * 512 * 1024 * 2


def test_ops_refuse_cpu_tensors():
    from rangeclip_b200 import ops
    with pytest.raises(RuntimeError):
        ops.tv_sums(torch.zeros(1, 1, 4, 4))
    with pytest.raises(RuntimeError):
        ops.eval_hist(torch.zeros(4, dtype=torch.long), torch.zeros(1, 1, 2, 2, dtype=torch.long),
                      torch.eye(3, dtype=torch.uint8), torch.arange(3))


def test_shared_embedding_and_fused_eval_entry_points_validate_on_cpu():
    """The round's added entry points fail loudly without a GPU and reject malformed arguments before any launch."""
    import rangeclip_b200 as R
    from rangeclip_b200 import _lib, ops
    L = _lib.lib()
    assert L.rc_infonce_bf16_rep4(None, _lib.RC_BF16, 1, 512, 64, None, None, 4, None, None, 1.0, None, None, None, None, None,
                                  None, None, None, None, 0, 0, None) == -1
    assert L.rc_eval_topk_hist_bf16(None, _lib.RC_BF16, 1, 512, 64, None, 4, None, 1, None, None, None, None, 4, None, None,
                                    None, 0, None) == -1
    assert _lib.bringup_lib().rc_debug_max_active_clusters(0, 640, 1024) == -1            # bad cluster size
    with pytest.raises(RuntimeError):                                      # CPU tensors
        ops.infonce_raw(torch.zeros(1, 256, 2, 4), torch.zeros(3, 256), torch.zeros(8, 4, dtype=torch.int32), torch.zeros(8, 4),
                        10.0, True, False, "bf16", rep=4)
    with pytest.raises(RuntimeError):
        ops.eval_topk_hist(torch.zeros(1, 64, 2, 4), torch.zeros(3, 64), torch.arange(3), 1, torch.zeros(1, 2, 4, dtype=torch.long),
                           torch.eye(3, dtype=torch.uint8), torch.arange(3), torch.zeros(5, 3, dtype=torch.long),
                           torch.zeros(3, dtype=torch.long))

    class M:
        log_temperature_text = torch.tensor(0.0)
        log_temperature_image = torch.tensor(0.0)
    with pytest.raises(RuntimeError):                                      # targets must be twice the embedding resolution
        R.compute_loss_shared2x2(M(), torch.zeros(1, 256, 4, 4), torch.zeros(1, 9, 8, dtype=torch.long), torch.zeros(5, 256),
                                 {"medium": {}, "hard": {}}, None, None)
    assert R.losses.group_2x2(torch.arange(16).view(1, 4, 4)).tolist() == [[[0, 1, 4, 5], [2, 3, 6, 7], [8, 9, 12, 13], [10, 11, 14, 15]]]


@pytest.mark.parametrize("case", ["dict", "list", "noimg", "medium"])
def test_contrast_set_builder_matches_reference(golden_dir, case):
    from rangeclip_b200 import build_contrast_indices
    from oracle import rangeclip_oracle as O
    g = np.load(os.path.join(golden_dir, f"loss_{case}.npz"))
    C = g["text"].shape[0]
    hard = {i: [int(v) for v in g["hard"][i]] for i in range(C)}
    med = {i: [int(v) for v in g["medium"][i]] for i in range(C)}
    sets = {"medium": med, "hard": hard}
    if str(g["sim_form"]) == "list":
        sets = {"medium": [med[i] for i in range(C)], "hard": [hard[i] for i in range(C)]}
    _, lab = O.sample_pixels(torch.tensor(g["X"]), torch.tensor(g["seg"]), torch.tensor(g["rand_idx"]))
    seed = int(g["seed"])
    np.random.seed(seed); torch.manual_seed(seed); random.seed(seed)
    got = build_contrast_indices(torch.unique(lab), C, sets, int(g["k_distractors"]), float(g["pcts"][0]),
                                 float(g["pcts"][1]), float(g["pcts"][2]), torch.device("cpu"))
    assert np.array_equal(got.numpy(), g["contrast"])


def test_candidate_set_builder_matches_reference(golden_dir):
    from rangeclip_b200 import build_reduced_candidates
    from oracle import rangeclip_oracle as O
    g = np.load(os.path.join(golden_dir, "predict.npz"))
    seg = torch.tensor(g["seg"])
    random.seed(int(g["seed"]))
    mine = build_reduced_candidates(seg, g["text"].shape[0], int(g["num_negatives"]))
    random.seed(int(g["seed"]))
    assert mine == O.build_candidate_set(seg, g["text"].shape[0], int(g["num_negatives"]))
    assert set(np.unique(g["topk"]).tolist()) <= set(mine)


def test_device_contrast_builder_oracle_properties():
    """oracle.contrast_build_device (the checker of rc_contrast_build): the SET semantics of model.py:234-268 -- present labels
    always in, curriculum distractors only from the similarity lists of present labels, random ones from the rest, sorted
    unique output, deterministic in the seed, capped at k_cap -- and the CSR form of the similarity tables (quirk Q3)."""
    import numpy as np
    from oracle import rangeclip_oracle as O
    from rangeclip_b200 import losses
    rng = np.random.default_rng(0)
    C = 200
    sets = {"hard": {c: [int(v) for v in rng.choice(C, 5, replace=False)] for c in range(0, C, 2)},
            "medium": {c: [int(v) for v in rng.choice(C, 5, replace=False)] for c in range(C)}}
    off, items = losses.similarity_csr(sets, C, False, True, "cpu")
    off, items = off.numpy(), items.numpy()
    for c in range(C):
        assert list(items[off[c]:off[c + 1]]) == (sets["hard"][c] if c % 2 == 0 else [])
    off_l, _ = losses.similarity_csr({"hard": [[1, 2], [3]], "medium": [[4]]}, C, True, True, "cpu")       # list form: never matches (Q3)
    assert int(off_l[-1]) == 0
    counts = np.zeros(C, dtype=np.int32)
    present = rng.choice(np.arange(1, C), 25, replace=False)
    counts[present] = 3
    counts[0] = 100
    lm, con, (K, flags, n_present, n_dis) = O.contrast_build_device(counts, off, items, 20, 10, 256, 99)
    members = con[:K].tolist()
    assert flags == 0 and n_present == 25 and n_dis == 30 and K == 55 and list(con[K:]) == [-1] * (256 - K)
    assert members == sorted(set(members)) and set(present.tolist()) <= set(members)
    pool = set(v for c in present.tolist() for v in sets["hard"].get(c, [])) - set(present.tolist())
    assert len(set(members) & pool) >= 20
    assert [lm[c] for c in members] == list(range(K)) and (lm >= 0).sum() == K
    again = O.contrast_build_device(counts, off, items, 20, 10, 256, 99)
    other = O.contrast_build_device(counts, off, items, 20, 10, 256, 100)
    assert np.array_equal(again[1], con) and not np.array_equal(other[1], con)
    _, con_c, (Kc, flags_c, _, n_dis_c) = O.contrast_build_device(counts, off, items, 20, 10, 40, 99)
    assert Kc == 40 and flags_c == 2 and n_dis_c == 15 and set(present.tolist()) <= set(con_c[:Kc].tolist())
    _, _, (Ko, flags_o, _, n_dis_o) = O.contrast_build_device(counts, off, items, 20, 10, 16, 99)
    assert Ko == 16 and flags_o & 1 and n_dis_o == 0


def test_fp32_single_read_entry_points_validate_on_cpu():
    """rc_infonce_prepass_tv / rc_tv_bwd_codes reject null pointers and rows that are not whole 8-pixel groups before any launch."""
    import ctypes
    from rangeclip_b200 import _lib, ops
    L = _lib.lib()
    buf = (ctypes.c_uint8 * 4096)()
    p = ctypes.addressof(buf)
    assert L.rc_infonce_prepass_tv(None, 1, 4, 2, 8, None, 0, None, None, None) == -1
    assert b"null pointer" in L.rc_last_error()
    assert L.rc_infonce_prepass_tv(p, 1, 4, 2, 12, p, 4096, p, None, None) != 0
    assert b"multiple of 8" in L.rc_last_error()
    assert L.rc_tv_bwd_codes(None, 1, 2, 8, None, None, _lib.RC_BF16, None, None, None) == -1
    assert L.rc_tv_bwd_codes(p, 1, 2, 12, p, None, _lib.RC_BF16, None, p, None) != 0
    assert b"multiple of 8" in L.rc_last_error()
    with pytest.raises(RuntimeError):                                      # CPU tensors
        ops.tv_backward_codes(torch.zeros(1, 1, 2, 1, dtype=torch.int32), torch.zeros(2))
