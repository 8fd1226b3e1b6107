"""Copies the UNMODIFIED reference model package into baseline/_ref/ so that it travels to the GPU box.

    python tools/install_reference.py            (build container only; __graft_entry__.build() calls it)

The reference (Jiang-Muyun/PointNet12) is plain Python without packaging metadata, so "installing" it is copying
`model/*.py` and the shipped checkpoint byte for byte; baseline/_ref/ is git-ignored (no reference source enters the
history) but not gpurun-ignored.  bench.py --impl reference imports it from there and runs `model.utils.load_pointnet`
+ the eval forward on the host CPU (CUDA hidden from that process), which is the reference's own PyTorch-CPU path.
A MANIFEST with the sha256 of every copied file is written next to them; bench.py verifies it before timing."""
import hashlib
import json
import os
import shutil
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = os.environ.get("PN_REFERENCE", "/root/reference")
DST = os.path.join(ROOT, "baseline", "_ref")
FILES = ["model/pointnet_util.py", "model/pointnet2.py", "model/pointnet.py", "model/utils.py", "model/chamfer.py",
         "checkpoints/pointnet2-inview-0.55884-0001.pth"]


def sha(path):
    h = hashlib.sha256()
    with open(path, "rb") as f:
        h.update(f.read())
    return h.hexdigest()


def install() -> bool:
    if not os.path.isdir(os.path.join(REF, "model")):
        return False
    manifest = {}
    for rel in FILES:
        src, dst = os.path.join(REF, rel), os.path.join(DST, rel)
        os.makedirs(os.path.dirname(dst), exist_ok=True)
        shutil.copyfile(src, dst)
        manifest[rel] = sha(dst)
    with open(os.path.join(DST, "MANIFEST.json"), "w") as f:
        json.dump({"source": "Jiang-Muyun/PointNet12 (unmodified copies)", "sha256": manifest}, f, indent=1)
    return True


def verify() -> bool:
    path = os.path.join(DST, "MANIFEST.json")
    if not os.path.exists(path):
        return False
    with open(path) as f:
        manifest = json.load(f)["sha256"]
    return all(os.path.exists(os.path.join(DST, rel)) and sha(os.path.join(DST, rel)) == h for rel, h in manifest.items())


if __name__ == "__main__":
    ok = install()
    print("installed" if ok else f"{REF} not found: nothing installed", "->", DST, "verified:", verify())
    sys.exit(0 if ok else 1)
