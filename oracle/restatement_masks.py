"""TEST INFRASTRUCTURE ONLY - numpy restatement of the reference's gradient mask and world-normal helpers
(SURVEY.md §8(f) rank 4), used to check csrc/masks.cu.  Pinned against the reference functions themselves by
tests/test_masks_oracle.py (golden: tests/golden/ref_mask_cases.npz, made by oracle/make_golden.py).

The reference computes the mask with torch float32 convolutions whose summation order is a library detail, so the
restatement works in float64 and also returns the quantities that are thresholded: a pixel whose value lies within
a small band of its threshold is a tie and is excluded from exact comparisons (band stated in the tests)."""

from __future__ import annotations

import numpy as np
from scipy.ndimage import correlate, correlate1d


def gaussian_taps(edge_sigma: float) -> np.ndarray:
    """initilizer.py:262-273 (float32 torch ops in the reference; float64 here)."""
    k = int(2 * edge_sigma * 3) + 1
    if k % 2 == 0:
        k += 1
    x = np.arange(k, dtype=np.float64) - k // 2
    g = np.exp(-0.5 * (x / edge_sigma) ** 2)
    return g / g.sum()


def gradient_mask(depth, normal=None, depth_threshold=0.2, normal_threshold=0.3, edge_sigma=1.0):
    """initilizer.py:236-328.  Returns (mask, relative depth gradient, normal gradient magnitude)."""
    d = depth.astype(np.float64)
    smooth = d
    if edge_sigma > 0:
        t = gaussian_taps(edge_sigma)
        smooth = correlate1d(correlate1d(d, t, axis=1, mode="constant"), t, axis=0, mode="constant")  # kernel_x then kernel_y
    sx = np.array([[-1, 0, 1], [-2, 0, 2], [-1, 0, 1]], np.float64)
    sy = np.array([[-1, -2, -1], [0, 0, 0], [1, 2, 1]], np.float64)
    dx, dy = correlate(smooth, sx, mode="constant"), correlate(smooth, sy, mode="constant")
    with np.errstate(divide="ignore", invalid="ignore"):
        rel = np.sqrt(dx**2 + dy**2) / (smooth + 1e-6)
    mask = rel > depth_threshold
    nmag = None
    if normal is not None:
        n = normal.astype(np.float64)
        acc = np.zeros(depth.shape)
        for ch in range(3):
            gy, gx = np.gradient(n[..., ch])
            acc += gx**2 + gy**2
        nmag = np.sqrt(acc)
        mask = mask | (nmag > normal_threshold)
    return mask, rel, nmag


def transform_normals(normal_map, cam_from_world, mask):
    """visualizer.py:346-376."""
    R = np.asarray(cam_from_world)[:3, :3]
    n = normal_map.reshape(-1, 3)[mask.reshape(-1) > 0]
    w = (R.T @ n.T).T
    return w / (np.linalg.norm(w, axis=1, keepdims=True) + 1e-8)
