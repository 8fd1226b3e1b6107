"""CPU oracle of the scan pre-processing path.  TEST INFRASTRUCTURE ONLY.

numpy restatement, statement by statement, of Semantic_KITTI_Utils.get (data_utils/kitti_utils.py:183-227),
points_basic_filter / hv_in_range / box_in_range (:238-280) and SemKITTI_Loader.__getitem__
(data_utils/SemKITTI_Loader.py:91-115 with pcd_normalize :23-30 and pcd_jitter :17-21).  Parity status: PINNED --
tests/golden/preprocess_scans.npz was produced by the reference's own functions (oracle/gen_golden_preprocess.py imports
data_utils/kitti_utils.py and SemKITTI_Loader.py with the absent cv2 / redis modules stubbed, points them at synthetic
.bin / .label files and records what get(), pcd_normalize, pcd_jitter and np.random.choice return).
The random draws (jitter noise, choice) are INPUTS here, as they are for the CUDA path's parity mode.
"""
import numpy as np

H_FOV, V_FOV = (-40, 40), (-20, 20)        # kitti_utils.py:222


def scan_filter(points: np.ndarray, raw_label: np.ndarray, learning_map: dict, inview: bool = True):
    """-> (kept indices into the scan in file order, class labels 0..18 of the kept points)."""
    label = raw_label & 0xFFFF                                                            # :205
    label = np.array([learning_map[int(x)] for x in label], dtype=np.int32)              # :213
    keep = label != 0                                                                      # :216
    if inview:
        x, y, z = points[:, 0], points[:, 1], points[:, 2]
        d = np.sqrt(x ** 2 + y ** 2 + z ** 2)                                              # :270
        h = np.logical_and(np.arctan2(y, x) > (-H_FOV[1] * np.pi / 180), np.arctan2(y, x) < (-H_FOV[0] * np.pi / 180))
        v = np.logical_and(np.arctan2(z, d) < (V_FOV[1] * np.pi / 180), np.arctan2(z, d) > (V_FOV[0] * np.pi / 180))
        box = np.logical_and.reduce((x > -10000, x < 10000, y > -10000, y < 10000, z > -10000, z < 10000,
                                     d > -10000, d < 10000))                               # set_filter defaults, :229-235
        keep = keep & h & v & box
    kept = np.nonzero(keep)[0]
    return kept.astype(np.int64), (label[kept] - 1).astype(np.int64)                       # :218


def pcd_normalize(pcd: np.ndarray) -> np.ndarray:
    pcd = pcd.copy()
    pcd[:, 0] = pcd[:, 0] / 70
    pcd[:, 1] = pcd[:, 1] / 70
    pcd[:, 2] = pcd[:, 2] / 3
    pcd[:, 3] = (pcd[:, 3] - 0.5) * 2
    return np.clip(pcd, -1, 1)


def scan_sample(points, raw_label, learning_map, npoints, choice, noise=None, inview=True):
    """One __getitem__: -> (pcd [npoints, 4] float32, label [npoints] int64).
    choice [npoints] = np.random.choice(length, npoints, replace=True); noise [length, 4] float32 =
    clip(sigma * randn(length, 4), -clip, clip).astype(float32) (train) or None (eval)."""
    kept, label = scan_filter(points, raw_label, learning_map, inview)
    pcd = pcd_normalize(points[kept].astype(np.float32))
    if noise is not None:
        jittered = np.asarray(noise, dtype=np.float32).copy()
        jittered += pcd                                                                    # SemKITTI_Loader.py:20
        pcd = jittered
    return pcd[choice], label[choice]
