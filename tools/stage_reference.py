"""Stage the UNMODIFIED reference files the full-training-step benchmark needs (BASELINE configs[0] / configs[2]: the
ResNet-18-UNet+ASPP backbone and the reference DepthUNet) under ``baseline/_ref/``.

``/root/reference`` does not exist on the GPU box; ``baseline/_ref/`` is git-ignored (no reference source enters the
history) but NOT gpurun-ignored, so the staged copy travels with the snapshot.  Run in the build container:

    python tools/stage_reference.py            # also called by __graft_entry__.build() when /root/reference exists

Nothing in the product (``rangeclip_b200/``) imports these files; only ``bench.py --workload full_step`` (the backbone
stays the reference's own PyTorch code) and its reference arm do."""
import os
import shutil
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = os.environ.get("RANGECLIP_REFERENCE", "/root/reference")
DST = os.path.join(ROOT, "baseline", "_ref")
FILES = [
    "utils/src/__init__.py", "utils/src/encoder.py", "utils/src/decoder.py", "utils/src/networks.py", "utils/src/net_utils.py",
    "utils/src/log_utils.py",
    "RangeCLIP/src/depth_segmentation_model/model.py", "RangeCLIP/src/depth_segmentation_model/dataloader.py",
    "RangeCLIP/src/depth_segmentation_model/validate.py", "RangeCLIP/src/depth_segmentation_model/datasets.py",
]


def stage() -> bool:
    if not os.path.isdir(REF):
        return False
    for rel in FILES:
        src = os.path.join(REF, rel)
        dst = os.path.join(DST, rel)
        os.makedirs(os.path.dirname(dst), exist_ok=True)
        shutil.copyfile(src, dst)
    return True


def import_reference_model():
    """DepthUNet of the staged reference (matplotlib stubbed: absent in the image and only used for visualisation)."""
    import types
    for m in ("matplotlib", "matplotlib.pyplot", "matplotlib.cm"):
        if m not in sys.modules:
            sys.modules[m] = types.ModuleType(m)
    sys.modules["matplotlib"].pyplot = sys.modules["matplotlib.pyplot"]
    sys.modules["matplotlib"].cm = sys.modules["matplotlib.cm"]
    if not os.path.isdir(DST):
        raise RuntimeError("baseline/_ref is missing: run tools/stage_reference.py in the build container")
    if DST not in sys.path:
        sys.path.insert(0, DST)
    from RangeCLIP.src.depth_segmentation_model.model import DepthUNet
    return DepthUNet


if __name__ == "__main__":
    print("staged" if stage() else f"{REF} not found: nothing staged", "->", DST)
