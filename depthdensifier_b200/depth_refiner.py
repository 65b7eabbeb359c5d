"""Drop-in ``DepthRefiner`` / ``RefinerConfig`` (reference: src/depthdensifier/depth_refiner.py:16-328).

Same constructor, same ``refine_depth`` signature, same result dictionary; the work runs in the
sm_100a kernels of libddn_b200.so (csrc/align.cu) in float32.  Differences, all deliberate:

* ``use_fp16`` is accepted and ignored: the kernels always compute in float32 (>= the reference's
  precision; the reference silently switches to FP16 on a GPU, depth_refiner.py:86).
* the <=500 subsample uses the deterministic hash permutation of ``hashperm.py`` instead of
  ``torch.randperm`` (depth_refiner.py:304), seeded by ``subsample_seed``.
* there is no CPU path: without a CUDA device ``refine_depth`` raises.
"""

from __future__ import annotations

from dataclasses import dataclass
from typing import Any

import numpy as np
import torch

from . import ops


@dataclass
class RefinerConfig:
    """Configuration for DepthRefiner parameters (field for field depth_refiner.py:16-31)."""

    min_correspondences: int = 50
    edge_margin: int = 10
    robust: bool = True
    outlier_threshold: float = 2.5
    use_fp16: bool = True
    skip_smoothing: bool = False
    adaptive_correspondences: bool = True
    verbose: int = 0


class DepthRefiner:
    """Per-view alignment of a monocular depth map to COLMAP sparse points on a B200."""

    def __init__(
        self,
        config: RefinerConfig | None = None,
        min_correspondences: int | None = None,
        edge_margin: int | None = None,
        robust: bool | None = None,
        outlier_threshold: float | None = None,
        use_fp16: bool | None = None,
        skip_smoothing: bool | None = None,
        adaptive_correspondences: bool | None = None,
        verbose: int | None = None,
        *,
        align_mode: str = "pwl",
        subsample_seed: int = 0,
        device: str | torch.device | None = None,
    ):
        config = config or RefinerConfig()

        def pick(v, d):
            return v if v is not None else d

        self.min_correspondences = pick(min_correspondences, config.min_correspondences)
        self.edge_margin = pick(edge_margin, config.edge_margin)
        self.robust = pick(robust, config.robust)
        self.outlier_threshold = pick(outlier_threshold, config.outlier_threshold)
        self.use_fp16 = pick(use_fp16, config.use_fp16)
        self.skip_smoothing = pick(skip_smoothing, config.skip_smoothing)
        self.adaptive_correspondences = pick(adaptive_correspondences, config.adaptive_correspondences)
        self.verbose = pick(verbose, config.verbose)
        self.align_mode = align_mode
        self.subsample_seed = subsample_seed
        self.device = torch.device(device) if device is not None else torch.device("cuda")
        self.dtype = torch.float32
        if self.verbose > 0:
            print("[DepthRefiner] Using CUDA backend (libddn_b200, sm_100a) with FP32")

    def options(self, **over) -> ops.AlignOptions:
        o = ops.AlignOptions(
            min_correspondences=self.min_correspondences,
            edge_margin=self.edge_margin,
            robust=self.robust,
            outlier_threshold=self.outlier_threshold,
            skip_smoothing=self.skip_smoothing,
            adaptive_correspondences=self.adaptive_correspondences,
            align_mode=self.align_mode,
            subsample_seed=self.subsample_seed,
        )
        for k, v in over.items():
            setattr(o, k, v)
        return o

    def refine_depth(
        self,
        depth_map: np.ndarray,
        normal_map: np.ndarray | None,
        points3D: np.ndarray,
        cam_from_world: np.ndarray,
        K: np.ndarray,
        mask: np.ndarray | None = None,
        **kwargs,
    ) -> dict[str, Any]:
        """Same contract as depth_refiner.py:207-328.  ``cam_from_world`` is 3x4 (the reference
        appends the homogeneous row itself, :98); ``normal_map`` is ignored as in the reference."""
        if not torch.cuda.is_available():
            raise ops.DDNError("DepthRefiner needs a CUDA device: depthdensifier_b200 has no CPU fallback")
        dev = self.device
        depth = torch.from_numpy(np.ascontiguousarray(depth_map, dtype=np.float32)).to(dev)[None]
        m = None
        if mask is not None:
            m = torch.from_numpy(np.ascontiguousarray(mask).astype(bool)).to(dev)[None]
        pts = np.ascontiguousarray(points3D, dtype=np.float64).reshape(-1, 3)
        pose = torch.from_numpy(np.ascontiguousarray(cam_from_world, dtype=np.float64)[:3, :4].copy()).to(dev)[None]
        kmat = torch.from_numpy(np.ascontiguousarray(K, dtype=np.float64)).to(dev)[None]
        sparse = torch.from_numpy(pts).to(dev)
        offsets = torch.tensor([0, pts.shape[0]], dtype=torch.int64, device=dev)
        refined, stats = ops.align_views(depth, m, pose.contiguous(), kmat.contiguous(), sparse, offsets, max(pts.shape[0], 1), self.options())
        st = ops.decode_stats(stats)[0]
        if st["status"] == ops.STATUS_REFINED:
            if self.verbose > 0:
                print(f"[DepthRefiner] Refined using {st['num_correspondences']} correspondences")
            return {
                "refined_depth": refined[0].cpu().numpy(),
                "num_correspondences": st["num_correspondences"],
                "outliers_removed": st["outliers_removed"],
                "scale_factor": st["scale_factor"],
            }
        # early-return paths hand back the input array itself (depth_refiner.py:259,278,285,299)
        if self.verbose > 0:
            print(f"[DepthRefiner] view not refined: {ops.STATUS_NAMES[st['status']]}")
        return {"refined_depth": depth_map, "num_correspondences": st["num_correspondences"], "scale_factor": 1.0}
