"""Pins the oracle (oracle/restatement.py) to the reference: bit-identical to the reference's own
code where /root/reference exists, and to the committed golden vectors everywhere."""

import contextlib
import io

import numpy as np
import pytest

from depthdensifier_b200.hashperm import hash_perm
from depthdensifier_b200.neighbours import all_views_table
from depthdensifier_b200.synthetic import SceneConfig, make_scene
from oracle import restatement as R
from oracle.run_reference import reference_available, run_reference_main, run_reference_refiner


def _densify_from_npz(g, **kw):
    V = g["mono_depth"].shape[0]
    return R.densify(
        g["mono_depth"], g["normal"], g["mask"], g["rgb"], g["sparse_xyz"], g["sparse_offsets"], g["cam_from_world"],
        g["intrinsics"], all_views_table(V), int(g["vote_threshold"]), depth_threshold=float(g["depth_threshold"]), **kw
    )


def test_golden_main_allviews(golden_dir):
    g = np.load(golden_dir / "ref_main_allviews.npz")
    out = _densify_from_npz(g, align=R.AlignConfig(adaptive_correspondences=False))
    assert np.array_equal(out["refined"], g["ref_refined"])
    assert np.array_equal(out["points"], g["ref_points"])
    assert np.array_equal(out["colors"], g["ref_colors"])
    assert np.array_equal(out["normals"], g["ref_normals"])
    assert np.array_equal(out["votes"], g["ref_votes"].astype(np.int64))
    assert np.array_equal(out["keep"], g["ref_keep"])
    assert np.array_equal(out["counts_per_view"], g["ref_counts_per_view"])
    assert np.array_equal(out["points"][out["keep"]], g["ref_kept_points"])
    assert np.array_equal(out["colors"][out["keep"]], g["ref_kept_colors"])
    assert 0 < (~out["keep"]).sum() < len(out["keep"])  # the filter does real work on this scene


def test_golden_main_subsample_and_skipped_view(golden_dir):
    g = np.load(golden_dir / "ref_main_subsample.npz")
    out = _densify_from_npz(g, align=R.AlignConfig(), randperm=lambda n: hash_perm(n, 0))
    ids = g["ref_refined_view_ids"]
    assert list(out["active_views"]) == list(ids) and 2 not in ids  # view 2 has no sparse points
    assert np.array_equal(out["refined"][ids], g["ref_refined"])
    assert np.array_equal(out["points"], g["ref_points"])
    assert np.array_equal(out["colors"], g["ref_colors"])
    assert np.array_equal(out["votes"], g["ref_votes"].astype(np.int64))
    assert np.array_equal(out["keep"], g["ref_keep"])


def test_golden_refiner_cases(golden_dir):
    g = np.load(golden_dir / "ref_refiner_cases.npz")
    specs = {
        "default_hashperm": (R.AlignConfig(), True, True),
        "no_subsample": (R.AlignConfig(adaptive_correspondences=False), False, True),
        "skip_smoothing": (R.AlignConfig(adaptive_correspondences=False, skip_smoothing=True), False, True),
        "not_robust": (R.AlignConfig(adaptive_correspondences=False, robust=False), False, True),
        "mask_none": (R.AlignConfig(adaptive_correspondences=False), False, False),
        "too_few": (R.AlignConfig(min_correspondences=5000), False, True),
        "wide_margin": (R.AlignConfig(adaptive_correspondences=False, edge_margin=30, outlier_threshold=1.0), False, True),
    }
    for name, (cfg, perm, use_mask) in specs.items():
        for v in range(3):
            lo, hi = int(g["sparse_offsets"][v]), int(g["sparse_offsets"][v + 1])
            r = R.refine_view(
                g["mono_depth"][v].copy(), g["sparse_xyz"][lo:hi], g["cam_from_world"][v], R.kmatrix(g["intrinsics"][v]),
                g["mask"][v] if use_mask else None, cfg, randperm=(lambda n: hash_perm(n, 0)) if perm else None,
            )
            assert np.array_equal(np.asarray(r["refined_depth"], np.float32), g[f"{name}/{v}/refined"]), (name, v)
            assert r["num_correspondences"] == int(g[f"{name}/{v}/num"]), (name, v)
            assert r.get("outliers_removed", -1) == int(g[f"{name}/{v}/removed"]), (name, v)
            assert r["scale_factor"] == float(g[f"{name}/{v}/scale"]), (name, v)
    assert int(g["too_few/0/removed"]) == -1  # early-return dict has no outliers_removed key


@pytest.mark.skipif(not reference_available(), reason="/root/reference not present (GPU box)")
def test_restatement_equals_unmodified_reference_main():
    sc = make_scene(SceneConfig(n_views=5, width=112, height=80, n_sparse=600, seed=11))
    with contextlib.redirect_stdout(io.StringIO()), contextlib.redirect_stderr(io.StringIO()):
        ref = run_reference_main(sc, vote_threshold=3, depth_threshold=0.8, randperm=lambda n: hash_perm(n, 5))
    out = R.densify(
        sc.mono_depth.numpy(), sc.normal.numpy(), sc.mask.numpy(), sc.rgb.numpy(), sc.sparse_xyz.numpy(),
        sc.sparse_offsets.numpy(), sc.cam_from_world.numpy(), sc.intrinsics.numpy(), all_views_table(5), 3,
        depth_threshold=0.8, randperm=lambda n: hash_perm(n, 5),
    )
    assert np.array_equal(out["refined"], ref["refined"])
    assert np.array_equal(out["points"], ref["points"])
    assert np.array_equal(out["colors"], ref["colors"])
    assert np.array_equal(out["votes"], ref["votes"])
    assert np.array_equal(out["keep"], ref["keep"])


@pytest.mark.skipif(not reference_available(), reason="/root/reference not present (GPU box)")
def test_restatement_equals_reference_refiner_stride_and_fallbacks():
    sc = make_scene(SceneConfig(n_views=2, width=96, height=64, n_sparse=300, seed=5))
    a = sc
    for kw in (dict(), dict(skip_smoothing=True), dict(min_correspondences=10_000), dict(edge_margin=40)):
        for v in range(2):
            lo, hi = int(a.sparse_offsets[v]), int(a.sparse_offsets[v + 1])
            args = (a.mono_depth[v].numpy().copy(), a.sparse_xyz[lo:hi].numpy(), a.cam_from_world[v].numpy(),
                    R.kmatrix(a.intrinsics[v].numpy()), a.mask[v].numpy())
            ref = run_reference_refiner(*args, adaptive_correspondences=False, **kw)
            cfg = R.AlignConfig(adaptive_correspondences=False, **kw)
            mine = R.refine_view(*args, cfg)
            assert np.array_equal(np.asarray(ref["refined_depth"]), np.asarray(mine["refined_depth"]))
            assert {k: ref[k] for k in ref if k != "refined_depth"} == {k: mine[k] for k in mine if k != "refined_depth"}


def test_explicit_bilinear_matches_grid_sample():
    import torch
    import torch.nn.functional as F

    rng = np.random.default_rng(0)
    d = rng.uniform(0.5, 5, (37, 53)).astype(np.float32)
    u = rng.uniform(0, 52, 500).astype(np.float32)
    v = rng.uniform(0, 36, 500).astype(np.float32)
    grid = torch.stack([torch.from_numpy(u) / 52 * 2 - 1, torch.from_numpy(v) / 36 * 2 - 1], -1)[None, None]
    ref = F.grid_sample(torch.from_numpy(d)[None, None], grid, mode="bilinear", padding_mode="zeros", align_corners=True).squeeze().numpy()
    mine = R.bilinear_align_corners(d, u, v)
    assert (mine == ref).mean() > 0.999  # same operation order as ATen's CPU kernel
    np.testing.assert_allclose(mine, ref, rtol=3e-6, atol=1e-6)


def test_voxel_fuse_definition():
    rng = np.random.default_rng(1)
    xyz = rng.uniform(-1, 1, (5000, 3)).astype(np.float32)
    rgb = rng.integers(0, 256, (5000, 3)).astype(np.uint8)
    k, m, c, n = R.voxel_fuse(xyz, rgb, 0.05)
    assert (np.diff(k.astype(np.int64)) > 0).all() and n.sum() == 5000
    o = R.voxel_origin(xyz, 0.05)
    keys = R.voxel_keys(xyz, 0.05, o)
    j = int(np.argmax(n))
    sel = keys == k[j]
    np.testing.assert_allclose(m[j], xyz[sel].astype(np.float64).mean(0), rtol=1e-6)
    assert np.array_equal(c[j], np.floor(rgb[sel].astype(np.float64).mean(0) + 0.5).astype(np.uint8))
