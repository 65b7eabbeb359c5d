"""TEST INFRASTRUCTURE ONLY - run the reference's UNMODIFIED code on a synthetic scene.

Runs only where ``/root/reference`` exists (the build container); the GPU box uses the committed
fixtures under ``tests/golden/`` that ``oracle/make_golden.py`` produced with this module.

* ``run_reference_main``: executes ``/root/reference/scripts/test.py:main`` (lines 95-370) under the
  stand-in ``pycolmap``/``moge`` modules of ``oracle/standins.py`` and captures its locals
  (points, colours, normals, ``floater_votes``, keep mask, per-view refined depth) with a profile
  hook at function return, so the source file is imported as it lies and never edited.
* ``run_reference_refiner``: calls the reference ``DepthRefiner.refine_depth``
  (``src/depthdensifier/depth_refiner.py:207-328``) on CPU float32.
"""

from __future__ import annotations

import importlib.util
import os
import sys
import tempfile
from pathlib import Path

import numpy as np

REFERENCE_ROOT = Path(os.environ.get("DDN_REFERENCE_ROOT", "/root/reference"))


def reference_available() -> bool:
    return (REFERENCE_ROOT / "scripts" / "test.py").is_file()


def _import_reference_script():
    from . import standins

    standins.install()
    src = str(REFERENCE_ROOT / "src")
    if src not in sys.path:
        sys.path.insert(0, src)
    spec = importlib.util.spec_from_file_location("ddn_reference_script", REFERENCE_ROOT / "scripts" / "test.py")
    mod = importlib.util.module_from_spec(spec)
    sys.modules["ddn_reference_script"] = mod
    spec.loader.exec_module(mod)
    return mod


def import_reference_refiner():
    src = str(REFERENCE_ROOT / "src")
    if src not in sys.path:
        sys.path.insert(0, src)
    # Import the module file directly: the package __init__ is fine too, but this keeps the
    # import independent of any same-named installed package.
    spec = importlib.util.spec_from_file_location(
        "ddn_reference_refiner", REFERENCE_ROOT / "src" / "depthdensifier" / "depth_refiner.py"
    )
    mod = importlib.util.module_from_spec(spec)
    sys.modules["ddn_reference_refiner"] = mod
    spec.loader.exec_module(mod)
    return mod


def run_reference_refiner(depth, points3D, cam_from_world, K, mask, randperm=None, **refiner_kwargs):
    """Reference stage 1 on CPU fp32. ``randperm`` optionally replaces ``torch.randperm`` (the
    reference's subsample at depth_refiner.py:304 is generator specific)."""
    import torch

    mod = import_reference_refiner()
    refiner_kwargs.setdefault("use_fp16", False)
    ref = mod.DepthRefiner(**refiner_kwargs)
    ref.device = torch.device("cpu")
    ref.dtype = torch.float32
    orig = torch.randperm
    if randperm is not None:
        torch.randperm = lambda n, device=None, **kw: torch.as_tensor(randperm(n))
    try:
        return ref.refine_depth(
            depth_map=depth, normal_map=None, points3D=points3D, cam_from_world=cam_from_world, K=K, mask=mask
        )
    finally:
        torch.randperm = orig


def run_reference_main(scene, downsample_density=1, vote_threshold=5, depth_threshold=0.7, randperm=None, **refiner_kwargs):
    """Execute the reference ``main`` on ``scene`` (a host ``depthdensifier_b200.synthetic.Scene``)."""
    import torch
    from PIL import Image as PILImage

    from . import standins

    mod = _import_reference_script()
    V = scene.n_views
    poses = scene.cam_from_world.cpu().numpy()
    intr = scene.intrinsics.cpu().numpy()
    H, W = scene.cfg.height, scene.cfg.width
    off = scene.sparse_offsets.cpu().numpy()
    sparse = scene.sparse_xyz.cpu().numpy()

    tmp = tempfile.mkdtemp(prefix="ddn_ref_")
    rec_key = os.path.join(tmp, "sparse")
    rec = standins.Reconstruction()
    standins.Reconstruction._REGISTRY[rec_key] = rec
    standins.FakeMoGeModel.queue = []
    for pid in range(sparse.shape[0]):
        rec.points3D[pid] = standins.Point3D(sparse[pid])
    for v in range(V):
        rec.cameras[v + 1] = standins.Camera(v + 1, W, H, intr[v])
        ids = list(range(int(off[v]), int(off[v + 1])))
        rec.images[v + 1] = standins.Image(v + 1, f"view_{v:04d}.png", v + 1, poses[v, :, :3], poses[v, :, 3], ids)
        PILImage.fromarray(scene.rgb[v].cpu().numpy()).save(os.path.join(tmp, f"view_{v:04d}.png"))
        if len(ids) > 0:
            standins.FakeMoGeModel.queue.append(
                (
                    scene.mono_depth[v].cpu().numpy().copy(),
                    scene.normal[v].cpu().numpy().copy(),
                    scene.mask[v].cpu().numpy().copy(),
                )
            )

    cfg = mod.ScriptConfig()
    cfg.paths.recon_path = Path(rec_key)
    cfg.paths.image_dir = Path(tmp)
    cfg.paths.output_model_dir = Path(tmp) / "out"
    cfg.processing.downsample_density = downsample_density
    cfg.filtering.vote_threshold = vote_threshold
    cfg.filtering.depth_threshold = depth_threshold
    refiner_kwargs.setdefault("use_fp16", False)
    for k, val in refiner_kwargs.items():
        setattr(cfg.refiner, k, val)

    captured = {}

    def hook(frame, event, arg):
        if event == "return" and frame.f_code.co_name == "main" and frame.f_code.co_filename.endswith("test.py"):
            loc = frame.f_locals
            for name in ("floater_votes", "points_to_keep_mask", "final_normals", "all_dense_points", "all_dense_colors"):
                if name in loc:
                    captured[name] = loc[name]
            captured["refined"] = {k: d["refined_depth"] for k, d in loc.get("cached_refinement_data", {}).items()}

    orig_randperm = torch.randperm
    orig_avail = torch.cuda.is_available
    torch.cuda.is_available = lambda: False  # the reference has no device argument (depth_refiner.py:85)
    if randperm is not None:
        torch.randperm = lambda n, device=None, **kw: torch.as_tensor(randperm(n))
    sys.setprofile(hook)
    try:
        mod.main(cfg)
    finally:
        sys.setprofile(None)
        torch.randperm = orig_randperm
        torch.cuda.is_available = orig_avail
        standins.Reconstruction._REGISTRY.pop(rec_key, None)

    pts_all = np.concatenate(captured["all_dense_points"], 0) if captured.get("all_dense_points") else np.zeros((0, 3))
    cols_all = (
        np.concatenate(captured["all_dense_colors"], 0) if captured.get("all_dense_colors") else np.zeros((0, 3), np.uint8)
    )
    counts = np.array([len(p) for p in captured.get("all_dense_points", [])], dtype=np.int64)
    out = {
        "points": pts_all,  # [N,3] f64 before filtering
        "colors": cols_all,
        "normals": captured.get("final_normals"),
        "votes": np.asarray(captured.get("floater_votes")),
        "keep": np.asarray(captured.get("points_to_keep_mask")),
        "counts_per_view": counts,
        "refined": np.stack([captured["refined"][k] for k in sorted(captured["refined"])]) if captured["refined"] else None,
        "refined_view_ids": np.array(sorted(captured["refined"]), dtype=np.int64) - 1,
        "kept_points": np.array(rec.added_xyz, dtype=np.float64).reshape(-1, 3),
        "kept_colors": np.array(rec.added_rgb, dtype=np.uint8).reshape(-1, 3),
    }
    return out


def import_reference_pchip():
    """The reference's fast_pchip_refiner module (it imports pycolmap at module top -> stand-in)."""
    from . import standins

    standins.install()
    src = str(REFERENCE_ROOT / "src")
    if src not in sys.path:
        sys.path.insert(0, src)
    import importlib

    return importlib.import_module("depthdensifier.fast_pchip_refiner")


def run_reference_pchip(depth, normal, points3D, cam_from_world, K, mask, rgb_image=None, **kwargs):
    """Reference FastPCHIPRefiner.refine_depth (fast_pchip_refiner.py:386-548) on CPU."""
    mod = import_reference_pchip()
    kwargs.setdefault("verbose", 0)
    ref = mod.FastPCHIPRefiner(**kwargs)
    ref.device = "cpu"
    return ref.refine_depth(depth, normal, points3D, cam_from_world, K, mask=mask, rgb_image=rgb_image)
