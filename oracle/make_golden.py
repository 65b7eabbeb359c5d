"""TEST INFRASTRUCTURE ONLY - regenerate tests/golden/*.npz from the reference's own code.

Run in the build container (needs /root/reference):  python -m oracle.make_golden
The fixtures hold the inputs AND the outputs of the unmodified reference so that machines without
/root/reference (the GPU box) can still pin the oracle and the CUDA path to it.
"""

from __future__ import annotations

import contextlib
import io
from pathlib import Path

import numpy as np

from depthdensifier_b200.hashperm import hash_perm
from depthdensifier_b200.synthetic import SceneConfig, make_scene
from oracle.restatement import kmatrix
from oracle.run_reference import run_reference_main, run_reference_pchip, run_reference_refiner

GOLDEN = Path(__file__).resolve().parent.parent / "tests" / "golden"


def scene_arrays(sc):
    return dict(
        mono_depth=sc.mono_depth.numpy(), normal=sc.normal.numpy(), mask=sc.mask.numpy(), rgb=sc.rgb.numpy(),
        sparse_xyz=sc.sparse_xyz.numpy(), sparse_offsets=sc.sparse_offsets.numpy(),
        cam_from_world=sc.cam_from_world.numpy(), intrinsics=sc.intrinsics.numpy(),
    )


def main():
    GOLDEN.mkdir(parents=True, exist_ok=True)
    # (1) whole main(), all-views semantics (K = V), no random subsample
    sc = make_scene(SceneConfig(n_views=6, width=96, height=72, n_sparse=384, seed=0))
    with contextlib.redirect_stdout(io.StringIO()), contextlib.redirect_stderr(io.StringIO()):
        ref = run_reference_main(sc, vote_threshold=2, adaptive_correspondences=False)
    np.savez_compressed(
        GOLDEN / "ref_main_allviews.npz", vote_threshold=2, depth_threshold=0.7, **scene_arrays(sc),
        ref_refined=ref["refined"], ref_points=ref["points"], ref_colors=ref["colors"], ref_normals=ref["normals"],
        ref_votes=ref["votes"].astype(np.int16), ref_keep=ref["keep"], ref_counts_per_view=ref["counts_per_view"],
        ref_kept_points=ref["kept_points"], ref_kept_colors=ref["kept_colors"],
    )
    print("ref_main_allviews:", ref["points"].shape, "kept", int(ref["keep"].sum()), "vote hist", np.bincount(ref["votes"]))

    # (2) main() with the 500-subsample driven by the hash permutation and one view without sparse points
    sc2 = make_scene(SceneConfig(n_views=5, width=96, height=72, n_sparse=900, seed=7))
    off = sc2.sparse_offsets.numpy().copy()
    # drop the sparse points of view 2 (scripts/test.py:136-137 skips such views)
    lo, hi = int(off[2]), int(off[3])
    keep_rows = np.ones(off[-1], bool)
    keep_rows[lo:hi] = False
    import torch
    sc2.sparse_xyz = sc2.sparse_xyz[torch.from_numpy(keep_rows)]
    off[3:] -= hi - lo
    sc2.sparse_offsets = torch.from_numpy(off)
    with contextlib.redirect_stdout(io.StringIO()), contextlib.redirect_stderr(io.StringIO()):
        ref2 = run_reference_main(sc2, vote_threshold=2, randperm=lambda n: hash_perm(n, 0))
    np.savez_compressed(
        GOLDEN / "ref_main_subsample.npz", vote_threshold=2, depth_threshold=0.7, **scene_arrays(sc2),
        ref_refined=ref2["refined"], ref_refined_view_ids=ref2["refined_view_ids"], ref_points=ref2["points"],
        ref_colors=ref2["colors"], ref_votes=ref2["votes"].astype(np.int16), ref_keep=ref2["keep"],
        ref_counts_per_view=ref2["counts_per_view"],
    )
    print("ref_main_subsample:", ref2["points"].shape, "kept", int(ref2["keep"].sum()), "views", ref2["refined_view_ids"])

    # (3) DepthRefiner.refine_depth cases
    sc3 = make_scene(SceneConfig(n_views=3, width=160, height=120, n_sparse=700, seed=3))
    a = scene_arrays(sc3)
    cases = {}
    specs = {
        "default_hashperm": dict(kw={}, perm=True, mask=True),
        "no_subsample": dict(kw=dict(adaptive_correspondences=False), perm=False, mask=True),
        "skip_smoothing": dict(kw=dict(adaptive_correspondences=False, skip_smoothing=True), perm=False, mask=True),
        "not_robust": dict(kw=dict(adaptive_correspondences=False, robust=False), perm=False, mask=True),
        "mask_none": dict(kw=dict(adaptive_correspondences=False), perm=False, mask=False),
        "too_few": dict(kw=dict(min_correspondences=5000), perm=False, mask=True),
        "wide_margin": dict(kw=dict(adaptive_correspondences=False, edge_margin=30, outlier_threshold=1.0), perm=False, mask=True),
    }
    for name, sp in specs.items():
        for v in range(3):
            lo, hi = int(a["sparse_offsets"][v]), int(a["sparse_offsets"][v + 1])
            r = run_reference_refiner(
                a["mono_depth"][v].copy(), a["sparse_xyz"][lo:hi], a["cam_from_world"][v], kmatrix(a["intrinsics"][v]),
                a["mask"][v] if sp["mask"] else None, randperm=(lambda n: hash_perm(n, 0)) if sp["perm"] else None, **sp["kw"],
            )
            cases[f"{name}/{v}/refined"] = np.asarray(r["refined_depth"], dtype=np.float32)
            cases[f"{name}/{v}/num"] = np.int64(r["num_correspondences"])
            cases[f"{name}/{v}/removed"] = np.int64(r.get("outliers_removed", -1))
            cases[f"{name}/{v}/scale"] = np.float64(r["scale_factor"])
    np.savez_compressed(GOLDEN / "ref_refiner_cases.npz", **a, **cases)
    print("ref_refiner_cases:", len(cases) // 4, "cases")

    # (4) FastPCHIPRefiner.refine_depth cases (fast_pchip_refiner.py:386-548)
    sc4 = make_scene(SceneConfig(n_views=2, width=176, height=128, n_sparse=900, seed=9))
    b = scene_arrays(sc4)
    pc = {}
    pspecs = {
        "default": dict(kw={}, mask=True, rgb=False, normal=True),
        "mask_none_no_normal": dict(kw={}, mask=False, rgb=False, normal=False),
        "image_edges": dict(kw=dict(use_image_edges=True, image_edge_threshold=12.0), mask=True, rgb=True, normal=True),
        "not_robust_tight": dict(kw=dict(robust=False, edge_threshold=0.02, edge_margin=5), mask=True, rgb=False, normal=True),
        "too_few": dict(kw=dict(min_correspondences=100000), mask=True, rgb=False, normal=True),
        "too_few_after_outliers": dict(kw=dict(outlier_threshold=1e-9, min_correspondences=50), mask=True, rgb=False, normal=True),
    }
    for name, sp in pspecs.items():
        for v in range(2):
            lo, hi = int(b["sparse_offsets"][v]), int(b["sparse_offsets"][v + 1])
            r = run_reference_pchip(
                b["mono_depth"][v].copy(), b["normal"][v] if sp["normal"] else None, b["sparse_xyz"][lo:hi], b["cam_from_world"][v],
                kmatrix(b["intrinsics"][v]), b["mask"][v] if sp["mask"] else None, rgb_image=b["rgb"][v] if sp["rgb"] else None, **sp["kw"])
            pc[f"{name}/{v}/refined"] = np.asarray(r["refined_depth"], dtype=np.float32)
            pc[f"{name}/{v}/scale"] = np.float64(r["scale"])
            pc[f"{name}/{v}/iters"] = np.int64(r["num_iterations"])
    np.savez_compressed(GOLDEN / "ref_pchip_cases.npz", **b, **pc)
    print("ref_pchip_cases:", len(pc) // 3, "cases")

    # (5) compute_depth_normal_gradient_mask (initilizer.py:236-328) on the same maps
    import importlib.util

    import torch
    from oracle.run_reference import REFERENCE_ROOT

    spec = importlib.util.spec_from_file_location("ddn_reference_init", REFERENCE_ROOT / "src" / "depthdensifier" / "initilizer.py")
    init = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(init)
    mc = {"mono_depth": b["mono_depth"], "normal": b["normal"]}
    for name, kw in {"default": {}, "tight": dict(depth_threshold=0.02, normal_threshold=0.1, edge_sigma=2.0),
                     "no_blur": dict(edge_sigma=0.0, depth_threshold=0.05)}.items():
        for v in range(2):
            for with_normal in (True, False):
                m = init.compute_depth_normal_gradient_mask(torch.from_numpy(b["mono_depth"][v]),
                                                            torch.from_numpy(b["normal"][v]) if with_normal else None, **kw)
                mc[f"{name}/{v}/{int(with_normal)}"] = m.numpy()
    np.savez_compressed(GOLDEN / "ref_mask_cases.npz", **mc)
    print("ref_mask_cases:", len(mc) - 2, "cases")


if __name__ == "__main__":
    main()
