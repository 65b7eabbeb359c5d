"""csrc/masks.cu against the reference function's outputs (golden) and the restatement."""

import numpy as np
import pytest
import torch

from oracle import restatement_masks as M

from test_masks_oracle import CASES, ties

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("name", sorted(CASES))
def test_gradient_mask_matches_reference(lib_built, golden_dir, name):
    from depthdensifier_b200.edge_masks import compute_depth_normal_gradient_mask

    g = np.load(golden_dir / "ref_mask_cases.npz")
    for v in range(2):
        for with_normal in (1, 0):
            depth, normal = torch.from_numpy(g["mono_depth"][v]), torch.from_numpy(g["normal"][v])
            got = compute_depth_normal_gradient_mask(depth.cuda(), normal.cuda() if with_normal else None, **CASES[name])
            assert got.is_cuda and got.dtype == torch.bool
            _, rel, nmag = M.gradient_mask(g["mono_depth"][v], g["normal"][v] if with_normal else None, **CASES[name])
            tie = ties(rel, nmag, CASES[name])
            ref = g[f"{name}/{v}/{with_normal}"]
            assert np.array_equal(got.cpu().numpy()[~tie], ref[~tie])
    # channel-first normals and CPU tensors are accepted like in the reference
    got2 = compute_depth_normal_gradient_mask(depth, normal.permute(2, 0, 1), **CASES[name])
    assert not got2.is_cuda and np.array_equal(got2.numpy(), compute_depth_normal_gradient_mask(depth.cuda(), normal.cuda(), **CASES[name]).cpu().numpy())


def test_transform_normals(lib_built, golden_dir):
    from depthdensifier_b200.edge_masks import transform_normals

    g = np.load(golden_dir / "ref_pchip_cases.npz")
    for pose in (g["cam_from_world"][0], np.vstack([g["cam_from_world"][1], [0, 0, 0, 1]]), g["cam_from_world"][1][:, :3]):
        got = transform_normals(g["normal"][0], pose, g["mask"][0])
        ref = M.transform_normals(g["normal"][0], pose, g["mask"][0])
        assert got.dtype == np.float64 and got.shape == ref.shape
        assert np.abs(got - ref).max() < 1e-14
    assert transform_normals(g["normal"][0], g["cam_from_world"][0], np.zeros_like(g["mask"][0])).shape == (0, 3)
