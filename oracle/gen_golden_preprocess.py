"""Generate tests/golden/preprocess_scans.npz with the REFERENCE'S OWN data path (build container only):

    python oracle/gen_golden_preprocess.py

data_utils/kitti_utils.py and data_utils/SemKITTI_Loader.py are imported unmodified from /root/reference; the modules
they import but never touch on this path and that are absent here (cv2, redis) are stubbed in sys.modules.  Two synthetic
raw scans (pointnet12_b200.synthetic.raw_scan, seeds 7000 / 7001) are written as .bin / .label files in the dataset's
directory layout under a temporary root; Semantic_KITTI_Utils(root, 'inview').get() reads and filters them, then
pcd_normalize, pcd_jitter and np.random.choice run under np.random.seed(70 + i) exactly as __getitem__ calls them.
Stored: the filtered points' indices, the draws (noise, choice) and the resulting (pcd, label) for train and eval.
"""
import os
import sys
import tempfile
import types

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = os.environ.get("PN_REFERENCE", "/root/reference")
sys.path.insert(0, ROOT)
from pointnet12_b200 import synthetic as syn  # noqa: E402

M, NPOINTS = 20000, 3000


def main():
    for name in ("cv2", "redis"):
        sys.modules.setdefault(name, types.ModuleType(name))
    sys.path.insert(0, REF)
    from data_utils.kitti_utils import Semantic_KITTI_Utils
    from data_utils.SemKITTI_Loader import pcd_jitter, pcd_normalize

    root = tempfile.mkdtemp(prefix="pn12_kitti_")
    os.makedirs(os.path.join(root, "sequences/00/velodyne"))
    os.makedirs(os.path.join(root, "sequences/00/labels"))
    scans = [syn.raw_scan(M, 7000 + i) for i in range(2)]
    for i, (p, l) in enumerate(scans):
        p.tofile(os.path.join(root, f"sequences/00/velodyne/{i:06d}.bin"))
        l.tofile(os.path.join(root, f"sequences/00/labels/{i:06d}.label"))
    utils = Semantic_KITTI_Utils(root, "inview")
    assert {int(k): int(v) for k, v in utils.learning_map.items()} == syn.SEMANTIC_KITTI_LEARNING_MAP
    arrays = {}
    for i, (p, l) in enumerate(scans):
        pts, lab = utils.get("00", i)
        # indices of the kept points: the filter keeps file order, so match row by row
        kept, j = [], 0
        for q in range(M):
            if j < len(pts) and np.array_equal(p[q], pts[j]) and utils.learning_map[int(l[q] & 0xFFFF)] == lab[j] + 1:
                kept.append(q)
                j += 1
        assert j == len(pts)
        np.random.seed(70 + i)
        pcd = pcd_jitter(pcd_normalize(pts))
        choice = np.random.choice(pcd.shape[0], NPOINTS, replace=True)
        np.random.seed(70 + i)                       # the same draws again, recorded
        noise = np.clip(0.01 * np.random.randn(*pts.shape), -0.05, 0.05).astype(pts.dtype)
        choice2 = np.random.choice(pts.shape[0], NPOINTS, replace=True)
        assert np.array_equal(choice, choice2) and np.array_equal(noise + pcd_normalize(pts), pcd)
        arrays[f"kept{i}"] = np.array(kept, dtype=np.int32)
        arrays[f"noise{i}"] = noise
        arrays[f"choice{i}"] = choice.astype(np.int32)
        arrays[f"train_pcd{i}"] = pcd[choice]
        arrays[f"eval_pcd{i}"] = pcd_normalize(pts)[choice]
        arrays[f"label{i}"] = lab[choice].astype(np.int8)
        arrays[f"checksum{i}"] = np.array(syn.checksum(p))
        print(f"scan {i}: {len(pts)} of {M} points kept")
    path = os.path.join(ROOT, "tests", "golden", "preprocess_scans.npz")
    np.savez_compressed(path, **arrays)
    print(f"wrote {path}: {os.path.getsize(path) / 1e6:.2f} MB")


if __name__ == "__main__":
    main()
