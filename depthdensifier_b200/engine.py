"""Batched device pipeline: align -> back-project + consistency vote -> voxel fusion.

This is the array-level form of the reference's ``main`` hot loops (scripts/test.py:130-333) for a
whole scene at once: structure-of-arrays inputs resident in HBM, one kernel per stage over all
views, no per-view host round trips."""

from __future__ import annotations

import math
from dataclasses import dataclass, field

import numpy as np
import torch

from . import ops
from .neighbours import default_vote_threshold


@dataclass
class DensifyConfig:
    align: ops.AlignOptions = field(default_factory=lambda: ops.AlignOptions(zero_unmasked_passthrough=True))
    filter: ops.FilterOptions = field(default_factory=ops.FilterOptions)
    vote_threshold: int | None = None  # None -> ceil(K/2); the reference default 5 assumes K = V
    voxel: float | None = 0.01  # None -> no fusion (reference behaviour: keep every point)


@dataclass
class DensifyResult:
    refined: torch.Tensor  # [V,H,W] f32
    stats: torch.Tensor  # [V,8] i32 raw ddn_view_stats
    xyz: torch.Tensor  # [n_src,Hs,Ws,3] f32 world
    votes: torch.Tensor  # [n_src,Hs,Ws] u8 (255 = no point)
    vote_threshold: int
    bbox: torch.Tensor | None = None  # [6] i32 ordered encoding
    grid: object | None = None
    voxel_keys: torch.Tensor | None = None
    voxel_xyz: torch.Tensor | None = None
    voxel_rgb: torch.Tensor | None = None
    voxel_count: torch.Tensor | None = None
    counts: torch.Tensor | None = None  # [2] i64: fused points, voxels

    def keep_mask(self) -> torch.Tensor:
        return self.votes < self.vote_threshold

    def num_points(self) -> int:
        return int((self.votes != 255).sum().item())


class DensifyEngine:
    def __init__(self, config: DensifyConfig | None = None, device: str | torch.device = "cuda"):
        if not torch.cuda.is_available():
            raise ops.DDNError("DensifyEngine needs a CUDA device: depthdensifier_b200 has no CPU fallback")
        self.cfg = config or DensifyConfig()
        self.device = torch.device(device)

    def align(self, depth, mask, cam_from_world, intr, sparse_xyz, sparse_offsets, max_sparse_per_view, out=None):
        V = depth.shape[0]
        kmat = torch.zeros((V, 3, 3), dtype=torch.float64, device=depth.device)
        kmat[:, 0, 0] = intr[:, 0]
        kmat[:, 1, 1] = intr[:, 1]
        kmat[:, 0, 2] = intr[:, 2]
        kmat[:, 1, 2] = intr[:, 3]
        kmat[:, 2, 2] = 1.0
        return ops.align_views(depth, mask, cam_from_world, kmat, sparse_xyz, sparse_offsets, max_sparse_per_view,
                               self.cfg.align, out=out)

    def run(self, depth, normal, mask, rgb, cam_from_world, intr, sparse_xyz, sparse_offsets, nbr,
            max_sparse_per_view: int | None = None, grid=None) -> DensifyResult:
        """All inputs are CUDA tensors: depth [V,H,W] f32, normal [V,H,W,3] f32, mask [V,H,W] bool,
        rgb [V,H,W,3] u8, cam_from_world [V,3,4] f64, intr [V,4] f64, sparse_xyz [S,3] f64,
        sparse_offsets [V+1] i64, nbr [V,K] i32."""
        cfg = self.cfg
        V, H, W = depth.shape
        K = nbr.shape[1]
        if max_sparse_per_view is None:
            off = sparse_offsets.cpu().numpy()
            max_sparse_per_view = int(np.max(np.diff(off))) if V > 0 else 1
        thr = cfg.vote_threshold if cfg.vote_threshold is not None else default_vote_threshold(K)
        refined, stats = self.align(depth, mask, cam_from_world, intr, sparse_xyz, sparse_offsets, max(max_sparse_per_view, 1))
        pair, src = ops.build_pair_tables(cam_from_world, intr, nbr, 0, V)
        bbox = ops.new_bbox(depth.device)
        xyz, votes = ops.backproject_filter(refined, normal, nbr, pair, src, 0, thr, cfg.filter, bbox=bbox)
        res = DensifyResult(refined=refined, stats=stats, xyz=xyz, votes=votes, vote_threshold=thr, bbox=bbox)
        if cfg.voxel is not None:
            if grid is None:
                bb = ops.decode_bbox(bbox)
                if not np.all(np.isfinite(bb)):
                    res.counts = torch.zeros(2, dtype=torch.int64, device=depth.device)
                    return res
                grid = ops.make_grid(bb[:3], bb[3:], cfg.voxel)
            s = cfg.filter.stride
            rgb_s = rgb if s == 1 else rgb[:, ::s, ::s].contiguous()
            k, x, c, n, counts = ops.voxel_fuse(xyz.view(-1, 3), rgb_s.view(-1, 3), votes.view(-1), thr, grid, row_len=xyz.shape[2])
            res.grid, res.voxel_keys, res.voxel_xyz, res.voxel_rgb, res.voxel_count, res.counts = grid, k, x, c, n, counts
        return res
