"""FastPCHIPRefiner on the GPU (csrc/pchip.cu through the C ABI) against the reference's outputs
(tests/golden/ref_pchip_cases.npz) and, stage by stage, against the restatement."""

import numpy as np
import pytest
import torch

from oracle import restatement_pchip as P

from test_pchip_oracle import CASES, case_args

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("name", sorted(CASES))
def test_fast_pchip_refiner_matches_reference(lib_built, golden_dir, name):
    from depthdensifier_b200.fast_pchip_refiner import FastPCHIPRefiner

    g = np.load(golden_dir / "ref_pchip_cases.npz")
    for v in range(2):
        args, kw = case_args(g, name, v)
        res = FastPCHIPRefiner(verbose=0, **kw).refine_depth(**args)
        ref = g[f"{name}/{v}/refined"]
        assert res["num_iterations"] == int(g[f"{name}/{v}/iters"]) and res["scale"] == float(g[f"{name}/{v}/scale"])
        assert set(res) == {"refined_depth", "scale", "energy_history", "num_iterations", "used_normals"}
        got = np.asarray(res["refined_depth"], np.float32)
        assert got.shape == ref.shape
        assert np.array_equal(got, ref), f"max abs diff {np.abs(got - ref).max()} at {np.argwhere(got != ref)[:3]}"
        if res["num_iterations"] == 0 and res["scale"] == 1.0:
            assert res["refined_depth"] is args["depth_map"]  # early returns alias the input (fast_pchip_refiner.py:457-463)


def test_edge_mask_bit_exact(lib_built, golden_dir):
    from depthdensifier_b200.fast_pchip_refiner import FastPCHIPRefiner

    g = np.load(golden_dir / "ref_pchip_cases.npz")
    for v in range(2):
        depth, normal, mask, rgb = g["mono_depth"][v], g["normal"][v], g["mask"][v], g["rgb"][v]
        for cfg, use_mask, use_normal, use_rgb in ((P.PchipConfig(), True, True, False), (P.PchipConfig(edge_threshold=0.02), False, False, False),
                                                   (P.PchipConfig(use_image_edges=True, image_edge_threshold=12.0), True, True, True),
                                                   (P.PchipConfig(edge_sigma=0.8), True, True, False)):
            ref = (P.detect_image_edges(rgb, mask if use_mask else None, cfg) if use_rgb
                   else P.detect_depth_edges(depth, mask if use_mask else None, normal if use_normal else None, cfg))
            r = FastPCHIPRefiner(verbose=0, edge_threshold=cfg.edge_threshold, edge_sigma=cfg.edge_sigma,
                                 use_image_edges=cfg.use_image_edges, image_edge_threshold=cfg.image_edge_threshold)
            dev = r.device
            got = r.detect_edges(torch.from_numpy(depth).to(dev), torch.from_numpy(mask.astype(np.uint8)).to(dev) if use_mask else None,
                                 torch.from_numpy(normal).to(dev) if use_normal else None, rgb if use_rgb else None)
            assert np.array_equal(got.cpu().numpy().astype(bool), ref)
            assert 0 < ref.mean() < 1


def test_refine_depth_from_colmap(lib_built, golden_dir):
    from depthdensifier_b200.colmap_io import Camera, Image, Point3D, rotmat_to_quat
    from depthdensifier_b200.fast_pchip_refiner import refine_depth_from_colmap

    g = np.load(golden_dir / "ref_pchip_cases.npz")
    v = 0
    lo, hi = int(g["sparse_offsets"][v]), int(g["sparse_offsets"][v + 1])
    pose = g["cam_from_world"][v]
    pts = {i + 1: Point3D(p) for i, p in enumerate(g["sparse_xyz"][lo:hi])}
    image = Image(1, rotmat_to_quat(pose[:, :3]), pose[:, 3], 1, "a.png", np.zeros((hi - lo, 2)), np.arange(1, hi - lo + 1))
    cam = Camera(1, "PINHOLE", g["mono_depth"].shape[2], g["mono_depth"].shape[1], g["intrinsics"][v])
    res = refine_depth_from_colmap(g["mono_depth"][v].copy(), g["normal"][v], image, cam, pts, verbose=0)
    ref = g[f"mask_none_no_normal/{v}/refined"]  # same call without a mask; normals only change the edge mask
    assert res["num_iterations"] == 1 and res["refined_depth"].shape == ref.shape
    assert np.isfinite(res["refined_depth"]).all() and abs(res["scale"] - float(g[f"default/{v}/scale"])) < 0.05
