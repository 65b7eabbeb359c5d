"""Gradient-mask restatement (oracle/restatement_masks.py) against the reference function's own outputs
(tests/golden/ref_mask_cases.npz from initilizer.py:236-328).  The reference thresholds float32 convolution results
whose summation order is a library detail: pixels within BAND (relative) of a threshold are ties."""

import numpy as np
import pytest

from oracle import restatement_masks as M

CASES = {"default": {}, "tight": dict(depth_threshold=0.02, normal_threshold=0.1, edge_sigma=2.0),
         "no_blur": dict(edge_sigma=0.0, depth_threshold=0.05)}
BAND = 2e-5


def ties(rel, nmag, kw):
    dt, nt = kw.get("depth_threshold", 0.2), kw.get("normal_threshold", 0.3)
    t = np.abs(rel - dt) <= BAND * np.maximum(np.abs(rel), dt) + 1e-7
    if nmag is not None:
        t |= np.abs(nmag - nt) <= BAND * max(nt, 1.0)
    return t | ~np.isfinite(rel)


@pytest.mark.parametrize("name", sorted(CASES))
def test_mask_restatement_matches_golden(golden_dir, name):
    g = np.load(golden_dir / "ref_mask_cases.npz")
    for v in range(2):
        for with_normal in (1, 0):
            mask, rel, nmag = M.gradient_mask(g["mono_depth"][v], g["normal"][v] if with_normal else None, **CASES[name])
            ref = g[f"{name}/{v}/{with_normal}"]
            tie = ties(rel, nmag, CASES[name])
            assert tie.mean() < 0.01 and 0 < ref.mean() < 1
            assert np.array_equal(mask[~tie], ref[~tie])


def test_transform_normals_definition():
    rng = np.random.default_rng(0)
    n = rng.normal(size=(6, 7, 3)).astype(np.float32)
    q = rng.normal(size=4)
    from depthdensifier_b200.colmap_io import quat_to_rotmat

    R = quat_to_rotmat(q)
    pose = np.hstack([R, rng.normal(size=(3, 1))])
    mask = rng.random((6, 7)) > 0.3
    w = M.transform_normals(n, pose, mask)
    assert w.shape == (int(mask.sum()), 3) and np.allclose(np.linalg.norm(w, axis=1), 1.0, atol=1e-6)
    assert np.allclose(w @ R.T, n[mask] / np.linalg.norm(n[mask], axis=1, keepdims=True), atol=1e-6)  # R n_world = n_cam
