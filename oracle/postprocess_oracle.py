"""CPU oracle of the motion post-processing (TEST INFRASTRUCTURE ONLY: imported by tests/ and nothing else).

Restates reference Diffusion_Stage/tools/visualization.py:20-26 (smooth_motion) and :107-120 (vis_motion's reshape,
`motions[i] *= window`, smooth_motion(motion, kernel=19)) with the reference's own third-party dependency,
scipy.signal.savgol_filter (scipy is unpinned in the reference, DS/requirements.txt; this image has scipy 1.x).
Pinned by tests/test_oracle_golden.py::test_postprocess_oracle_* against closed-form properties of the filter
(polynomials of degree <= order are reproduced exactly, including at the edges) -- the reference ships no fixture
for this function and tools/visualization.py is not importable here (moviepy, mmcv missing).
"""
import numpy as np
from scipy.signal import savgol_filter


def smooth_motion(kp_pred: np.ndarray, kernel: int = 11, order: int = 5) -> np.ndarray:
    """visualization.py:20-26 -- in place, per keypoint i and axis j, along frames."""
    for i in range(kp_pred.shape[1]):
        for j in range(2):
            kp_pred[:, i, j] = savgol_filter(kp_pred[:, i, j], kernel, order)
    return kp_pred


def vis_motion_keypoints(motions: np.ndarray, window: float = 600, kernel: int = 19) -> np.ndarray:
    """visualization.py:107-120 up to the smoothed pixel keypoints: (num_conductor, T, 26) -> (num_conductor, T, 13, 2)."""
    motions = np.array(motions)
    motions = motions.reshape([motions.shape[0], motions.shape[1], 13, 2])
    out = []
    for i in range(len(motions)):
        motions[i] *= window
        out.append(smooth_motion(motions[i], kernel=kernel))
    return np.stack(out)
