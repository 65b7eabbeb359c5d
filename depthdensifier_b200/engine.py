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
    max_grid_cells: int = 1 << 33  # capacity of the fusion session's occupancy bitmap (16 bytes per 96 cells)
    dedup_sparse: bool = False  # N5: no dense voxel where the sparse cloud already has a point (reference: plain append)
    # ShardedDensifier.run only: stage 1 of a step runs on its own stream and may start while the previous step is still
    # fusing / merging (its NVLink-bound exchange leaves the SMs idle).  The caller's inputs must then be complete when
    # run() is called - they are not ordered against work the caller queued on the current stream just before.
    overlap_align: bool = False


def clamp_vote_threshold(thr: int) -> int:
    """Votes are stored as u8, saturated at 254, with 255 = "pixel has no point": a threshold of 255 or more keeps
    every point that exists (and never a pixel without one), 0 or less keeps none."""
    return max(0, min(int(thr), 255))


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
    session: object | None = None  # ops.FuseSession that produced the voxels (device-resident grid)
    bbox_valid: torch.Tensor | None = None  # [6] i32: box of every back-projected pixel (the alignment kernel's bound)

    def keep_mask(self) -> torch.Tensor:
        return (self.votes != 255) & (self.votes < clamp_vote_threshold(self.vote_threshold))

    def host_grid(self):
        """The voxel grid as a host struct (synchronises; raises when the device could not build one)."""
        if self.grid is None and self.session is not None:
            self.grid = self.session.host_grid()
        return self.grid

    def check(self) -> int:
        """Synchronises and validates the fusion outputs; returns the number of voxels."""
        if self.counts is None:
            return 0
        if self.session is not None:
            st = self.session.grid_state()
            if st.status == 1:  # no valid point at all
                return 0
            self.host_grid()
        mv = ops.checked_voxel_count(self.counts)
        if self.voxel_keys is not None and mv > self.voxel_keys.shape[0]:
            raise ops.DDNError(f"voxel fusion: {mv} voxels exceed the output capacity {self.voxel_keys.shape[0]}")
        return mv

    def num_points(self) -> int:
        return int((self.votes != 255).sum().item())


class DensifyEngine:
    def __init__(self, config: DensifyConfig | None = None, device: str | torch.device = "cuda"):
        if not torch.cuda.is_available():
            raise ops.DDNError("DensifyEngine needs a CUDA device: depthdensifier_b200 has no CPU fallback")
        self.cfg = config or DensifyConfig()
        self.device = torch.device(device)

    def align(self, depth, mask, cam_from_world, intr, sparse_xyz, sparse_offsets, max_sparse_per_view, out=None,
              src_table=None, bbox=None):
        V = depth.shape[0]
        kmat = torch.zeros((V, 3, 3), dtype=torch.float64, device=depth.device)
        kmat[:, 0, 0] = intr[:, 0]
        kmat[:, 1, 1] = intr[:, 1]
        kmat[:, 0, 2] = intr[:, 2]
        kmat[:, 1, 2] = intr[:, 3]
        kmat[:, 2, 2] = 1.0
        return ops.align_views(depth, mask, cam_from_world, kmat, sparse_xyz, sparse_offsets, max_sparse_per_view,
                               self.cfg.align, out=out, src_table=src_table, bbox=bbox)

    def session(self) -> ops.FuseSession:
        if getattr(self, "_session", None) is None or self._session.max_cells != self.cfg.max_grid_cells:
            self._session = ops.FuseSession(self.device, self.cfg.max_grid_cells)
        return self._session

    def run(self, depth, normal, mask, rgb, cam_from_world, intr, sparse_xyz, sparse_offsets, nbr,
            max_sparse_per_view: int | None = None, grid=None, sync: bool = True) -> DensifyResult:
        """All inputs are CUDA tensors: depth [V,H,W] f32, normal [V,H,W,3] f32, mask [V,H,W] bool,
        rgb [V,H,W,3] u8, cam_from_world [V,3,4] f64, intr [V,4] f64, sparse_xyz [S,3] f64,
        sparse_offsets [V+1] i64, nbr [V,K] i32.

        No host round trip between the stages: the voxel grid is derived on the device from the bounding box the
        alignment kernel produces, the consistency kernel marks the occupancy of the points it keeps, and the
        fusion passes read the grid from device memory.  ``sync`` (default): wait at the end, validate
        (``DensifyResult.check()``) and trim the voxel arrays to their count; without it the arrays keep their
        capacity and ``counts`` [2] on the device says how many entries are valid."""
        cfg = self.cfg
        V, H, W = depth.shape
        K = nbr.shape[1]
        if max_sparse_per_view is None:
            off = sparse_offsets.cpu().numpy()
            max_sparse_per_view = int(np.max(np.diff(off))) if V > 0 else 1
        thr = clamp_vote_threshold(cfg.vote_threshold if cfg.vote_threshold is not None else default_vote_threshold(K))
        pair, src = ops.build_pair_tables(cam_from_world, intr, nbr, 0, V, H, W)
        fuse = cfg.voxel is not None
        box = ops.new_bbox(depth.device) if fuse and grid is None else None
        kmat = None
        refined = torch.empty_like(depth)
        refined, stats = self.align(depth, mask, cam_from_world, intr, sparse_xyz, sparse_offsets, max(max_sparse_per_view, 1),
                                    out=refined, src_table=src if box is not None else None, bbox=box)
        sess = None
        if fuse:
            sess = self.session()
            if grid is None:
                sess.begin([box], cfg.voxel)
            else:
                sess.begin_grid(grid)
        bbox = ops.new_bbox(depth.device)
        xyz, votes = ops.backproject_filter(refined, normal, nbr, pair, src, 0, thr, cfg.filter, bbox=bbox, mark=sess)
        res = DensifyResult(refined=refined, stats=stats, xyz=xyz, votes=votes, vote_threshold=thr, bbox=bbox)
        if fuse:
            s = cfg.filter.stride
            rgb_s = rgb if s == 1 else rgb[:, ::s, ::s].contiguous()
            if cfg.dedup_sparse and sparse_xyz.shape[0] > 0:
                sess.unmark_points(sparse_xyz.float().contiguous())
            k, x, c, n, counts = ops.fuse_finish(sess, xyz.view(-1, 3), rgb_s.view(-1, 3), votes.view(-1), thr, row_len=xyz.shape[2])
            res.session, res.grid = sess, grid
            res.voxel_keys, res.voxel_xyz, res.voxel_rgb, res.voxel_count, res.counts = k, x, c, n, counts.clone()
            if sync:
                mv = res.check()
                res.voxel_keys, res.voxel_xyz, res.voxel_rgb, res.voxel_count = k[:mv], x[:mv], c[:mv], n[:mv]
        res.bbox_valid = box
        return res
