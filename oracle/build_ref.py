"""Recipe for oracle/_ref/: the reference's OWN Python modules of the hot path, staged (unmodified) from where they lie
under /root/reference so that bench.py's reference arm and the tests can run the real reference implementation on the GPU
box, where /root/reference does not exist.

    python oracle/build_ref.py            (also called by __graft_entry__.build() when /root/reference is present)

oracle/_ref/ is git-ignored (no reference source enters the history) but NOT gpurun-ignored, so it travels with the
snapshot exactly like the built .so files.  Nothing here is product code: only tests/, smoke() and bench.py's
cpu_baseline / --impl reference legs may use it (through oracle/ref_harness.py).

Staged files (byte-identical copies, SHA-256 recorded in oracle/_ref/MANIFEST.json):
    models/{__init__,layers,miniViT,unet_adaptive_bins}.py, loss.py, ExternalInfoLoaders/{__init__,SemanticsLoader,
    InstanceSegmentationLoader}.py
Not staged: pytorch3d (third party, absent: ref_harness injects oracle.adabins_oracle.chamfer_distance, the restatement of
its v0.6.1 algorithm, below the reference's own BinsChamferLoss wrapper) and geffnet (torch.hub, network: the
geffnet-shaped random-init backbone of the product package is passed to the reference's UnetAdaptiveBins constructor).
"""
import hashlib
import json
import os
import shutil
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
REF = os.environ.get("MDE_REFERENCE_ROOT", "/root/reference")
DST = os.path.join(HERE, "_ref")
FILES = ["models/__init__.py", "models/layers.py", "models/miniViT.py", "models/unet_adaptive_bins.py", "loss.py",
         "ExternalInfoLoaders/__init__.py", "ExternalInfoLoaders/SemanticsLoader.py",
         "ExternalInfoLoaders/InstanceSegmentationLoader.py"]


def build(quiet=False):
    """Returns True if oracle/_ref/ is (now) populated, False if the reference tree is not available here."""
    if not os.path.isdir(REF):
        return os.path.exists(os.path.join(DST, "MANIFEST.json"))
    manifest = {}
    for rel in FILES:
        src, dst = os.path.join(REF, rel), os.path.join(DST, rel)
        os.makedirs(os.path.dirname(dst), exist_ok=True)
        shutil.copyfile(src, dst)
        with open(dst, "rb") as f:
            manifest[rel] = hashlib.sha256(f.read()).hexdigest()
    with open(os.path.join(DST, "MANIFEST.json"), "w") as f:
        json.dump({"source": REF, "files": manifest}, f, indent=1)
    if not quiet:
        print(f"oracle/_ref: staged {len(FILES)} reference files from {REF}")
    return True


if __name__ == "__main__":
    sys.exit(0 if build() else 1)
