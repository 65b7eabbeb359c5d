"""Drop-in ``FastPCHIPRefiner`` / ``FastPCHIPRefinerConfig`` (reference:
src/depthdensifier/fast_pchip_refiner.py:22-605; SURVEY.md §8(f) rank 3 - the reference's second alignment
algorithm, not wired into its pipeline).

Same constructor (config object and/or per-field overrides, the unused ``lambda1`` ... compatibility arguments
included), same ``refine_depth`` signature and result dictionary (``refined_depth``, ``scale``,
``energy_history``, ``num_iterations``, ``used_normals``).  The split of work is the reference's own: the O(C)
correspondence logic (project the sparse points, bilinear sample, MAD outliers, ``np.unique``) runs in numpy on the
host, everything per pixel - Gaussian smoothing, gradients, edge mask and dilation, the cubic-Hermite remap, the
70/30 edge blend and the 3x3 median - runs in the sm_100a kernels of ``csrc/pchip.cu``, bit-identical to
scipy / numpy / torch-CPU float32 (tests/test_gpu_pchip.py).  There is no CPU path: without a CUDA device
``refine_depth`` raises.  The depth map is processed as float32 (what MoGe produces).
"""

from __future__ import annotations

import ctypes as C
from dataclasses import dataclass
from typing import Any

import numpy as np
import torch

from . import _lib
from ._lib import DDNError


@dataclass
class FastPCHIPRefinerConfig:
    """Configuration for FastPCHIPRefiner parameters (field for field fast_pchip_refiner.py:50-66)."""

    min_correspondences: int = 100
    edge_margin: int = 20
    edge_threshold: float = 0.1
    edge_sigma: float = 2.0
    robust: bool = True
    outlier_threshold: float = 3.0
    use_image_edges: bool = False
    image_edge_threshold: float = 30.0
    scale_filter_factor: float = 2.0
    verbose: int = 1


def gaussian_taps(sigma: float, truncate: float = 4.0) -> np.ndarray:
    """Taps at distance 0..radius of scipy.ndimage's normalised Gaussian (``_gaussian_kernel1d``, order 0)."""
    radius = int(truncate * float(sigma) + 0.5)
    x = np.arange(-radius, radius + 1)
    phi = np.exp(-0.5 / (float(sigma) * float(sigma)) * x**2)
    phi = phi / phi.sum()
    return phi[radius:].copy()


class FastPCHIPRefiner:
    """Edge-aware cubic-Hermite alignment of a monocular depth map to COLMAP sparse points on a B200."""

    def __init__(self, config: FastPCHIPRefinerConfig | None = None, min_correspondences: int | None = None,
                 edge_margin: int | None = None, edge_threshold: float | None = None, edge_sigma: float | None = None,
                 robust: bool | None = None, outlier_threshold: float | None = None, scale_filter_factor: float | None = None,
                 verbose: int | None = None, use_image_edges: bool | None = None, image_edge_threshold: float | None = None,
                 # compatibility parameters of the reference (accepted, not used)
                 lambda1: float | None = None, lambda2: float | None = None, k_sigmoid: float | None = None,
                 max_iter: int | None = None, cg_max_iter: int | None = None, cg_tol: float | None = None,
                 convergence_tol: float | None = None, *, device: str | torch.device | None = None):
        config = config or FastPCHIPRefinerConfig()

        def pick(v, d):
            return v if v is not None else d

        self.min_correspondences = pick(min_correspondences, config.min_correspondences)
        self.edge_margin = pick(edge_margin, config.edge_margin)
        self.edge_threshold = pick(edge_threshold, config.edge_threshold)
        self.edge_sigma = pick(edge_sigma, config.edge_sigma)
        self.robust = pick(robust, config.robust)
        self.outlier_threshold = pick(outlier_threshold, config.outlier_threshold)
        self.scale_filter_factor = pick(scale_filter_factor, config.scale_filter_factor)
        self.verbose = pick(verbose, config.verbose)
        self.use_image_edges = pick(use_image_edges, config.use_image_edges)
        self.image_edge_threshold = pick(image_edge_threshold, config.image_edge_threshold)
        self.device = torch.device(device) if device is not None else torch.device("cuda")
        if self.verbose > 0:
            print("[FastPCHIP] Using the sm_100a kernels of libddn_b200")

    # -- device side ------------------------------------------------------------------------------------
    def _require_cuda(self):
        if not torch.cuda.is_available():
            raise DDNError("no CUDA device available: depthdensifier_b200 has no CPU fallback")

    def _workspace(self, lib, h, w):
        n = C.c_int64(0)
        _lib.check(lib.ddn_pchip_workspace_bytes(h, w, C.byref(n)))
        return torch.empty(n.value, dtype=torch.uint8, device=self.device), n.value

    def detect_edges(self, depth_d, mask_d, normal_d, rgb_image) -> torch.Tensor:
        """Edge mask [H,W] u8 on the device (fast_pchip_refiner.py:187-273)."""
        lib = _lib.load()
        h, w = depth_d.shape
        taps = gaussian_taps(self.edge_sigma)
        taps_c = (C.c_double * len(taps))(*taps.tolist())
        gray_d = None
        if self.use_image_edges and rgb_image is not None:
            rgb = np.asarray(rgb_image)
            gray = (0.299 * rgb[:, :, 0] + 0.587 * rgb[:, :, 1] + 0.114 * rgb[:, :, 2]) if rgb.ndim == 3 else rgb
            gray_d = torch.from_numpy(np.ascontiguousarray(gray, dtype=np.float64)).to(self.device)
        edge = torch.empty((h, w), dtype=torch.uint8, device=self.device)
        ws, nbytes = self._workspace(lib, h, w)
        p = lambda t: C.c_void_p(t.data_ptr()) if t is not None else C.c_void_p(0)
        with torch.cuda.device(self.device):
            _lib.check(lib.ddn_pchip_edge_mask(h, w, p(depth_d), p(mask_d), p(normal_d), p(gray_d), taps_c, len(taps) - 1,
                                               float(np.float32(self.edge_threshold)), float(self.image_edge_threshold), p(edge),
                                               p(ws), nbytes, C.c_void_p(torch.cuda.current_stream().cuda_stream)))
        return edge

    def _apply(self, depth_d, mask_d, edge_d, x, y) -> torch.Tensor:
        lib = _lib.load()
        h, w = depth_d.shape
        kx = torch.from_numpy(np.ascontiguousarray(x, dtype=np.float64)).float().to(self.device)  # torch.from_numpy(x).float()
        ky = torch.from_numpy(np.ascontiguousarray(y, dtype=np.float64)).float().to(self.device)
        out = torch.empty_like(depth_d)
        ws, nbytes = self._workspace(lib, h, w)
        p = lambda t: C.c_void_p(t.data_ptr())
        with torch.cuda.device(self.device):
            _lib.check(lib.ddn_pchip_apply(h, w, p(depth_d), p(mask_d), p(edge_d), p(kx), p(ky), len(x), p(out), p(ws), nbytes,
                                           C.c_void_p(torch.cuda.current_stream().cuda_stream)))
        return out

    # -- host side (as in the reference) ----------------------------------------------------------------
    @staticmethod
    def project_points_to_image(points3D, cam_from_world, K):
        """fast_pchip_refiner.py:142-153."""
        pts_h = np.hstack([points3D, np.ones((points3D.shape[0], 1))])
        T = np.vstack([cam_from_world, [0, 0, 0, 1]])
        cam = (T @ pts_h.T)[:3, :].T
        z = cam[:, 2].copy()
        ok = z > 1e-6
        uv = np.zeros((len(z), 2))
        uv[ok] = ((K @ (cam[ok] / z[ok, None]).T).T)[:, :2]
        return uv, z, ok

    @staticmethod
    def sample_depth_at_points(depth_map, pts2d):
        """fast_pchip_refiner.py:155-185 (weights from the clipped tap coordinates)."""
        h, w = depth_map.shape
        x, y = pts2d[:, 0], pts2d[:, 1]
        x0 = np.floor(x).astype(int)
        y0 = np.floor(y).astype(int)
        x1, y1 = np.clip(x0 + 1, 0, w - 1), np.clip(y0 + 1, 0, h - 1)
        x0, y0 = np.clip(x0, 0, w - 1), np.clip(y0, 0, h - 1)
        return ((x1 - x) * (y1 - y) * depth_map[y0, x0] + (x - x0) * (y1 - y) * depth_map[y0, x1]
                + (x1 - x) * (y - y0) * depth_map[y1, x0] + (x - x0) * (y - y0) * depth_map[y1, x1])

    def _remove_outliers(self, z_colmap, z_depth):
        """fast_pchip_refiner.py:275-298."""
        med_ratio = np.median(z_colmap / (z_depth + 1e-6))
        res = z_colmap - z_depth * med_ratio
        dev = np.abs(res - np.median(res))
        mad = np.median(dev)
        if mad < 1e-6:
            mad = np.std(res) * 0.6745
        inl = dev < self.outlier_threshold * mad
        return z_colmap[inl], z_depth[inl], int((~inl).sum())

    def refine_depth(self, depth_map: np.ndarray, normal_map: np.ndarray | None, points3D: np.ndarray, cam_from_world: np.ndarray,
                     K: np.ndarray, mask: np.ndarray | None = None, depth_uncertainty: np.ndarray | None = None,
                     rgb_image: np.ndarray | None = None, normal_uncertainty: np.ndarray | None = None) -> dict[str, Any]:
        """fast_pchip_refiner.py:386-548."""
        self._require_cuda()
        depth32 = np.ascontiguousarray(depth_map, dtype=np.float32)
        h, w = depth32.shape
        use_normals = normal_map is not None
        result = {"refined_depth": depth_map, "scale": 1.0, "energy_history": [], "num_iterations": 0, "used_normals": use_normals}
        depth_d = torch.from_numpy(depth32).to(self.device)
        mask_d = torch.from_numpy(np.ascontiguousarray(mask, dtype=np.uint8)).to(self.device) if mask is not None else None
        normal_d = None
        if normal_map is not None and normal_map.shape[-1] == 3:
            normal_d = torch.from_numpy(np.ascontiguousarray(normal_map, dtype=np.float32)).to(self.device)
        edge_d = self.detect_edges(depth_d, mask_d, normal_d, rgb_image)
        edge = edge_d.cpu().numpy().astype(bool)

        uv, z3, ok = self.project_points_to_image(np.asarray(points3D, dtype=np.float64), np.asarray(cam_from_world, dtype=np.float64),
                                                  np.asarray(K, dtype=np.float64))
        m = self.edge_margin
        inb = (uv[:, 0] >= m) & (uv[:, 0] < w - m) & (uv[:, 1] >= m) & (uv[:, 1] < h - m) & ok
        uv, z3 = uv[inb], z3[inb]
        if len(uv) == 0:
            if self.verbose > 0:
                print("[FastPCHIP] No valid correspondences found")
            return result
        samp = self.sample_depth_at_points(depth32, uv)
        u, v = np.round(uv[:, 0]).astype(int), np.round(uv[:, 1]).astype(int)
        good = (samp > 0) & (z3 > 0) & np.isfinite(samp) & ~edge[v, u]
        z_depth, z_colmap = samp[good], z3[good]
        if len(z_depth) < self.min_correspondences:
            if self.verbose > 0:
                print(f"[FastPCHIP] Too few correspondences ({len(z_depth)} < {self.min_correspondences})")
            return result
        if self.robust:
            zc, zd, n_out = self._remove_outliers(z_colmap, z_depth)
            if self.verbose > 0 and n_out > 0:
                print(f"[FastPCHIP] Removed {n_out} outliers")
        else:
            zc, zd = z_colmap, z_depth
        if len(zd) < self.min_correspondences:
            scale = float(np.median(z_colmap / z_depth)) if np.all(z_depth > 0) else 1.0
            result.update(refined_depth=depth_map * scale, scale=scale)
            return result
        ux, ui = np.unique(zd, return_index=True)
        uy = zc[ui]
        if len(ux) < 2:
            raise DDNError("FastPCHIP needs at least two distinct correspondences")
        if mask_d is None:
            mask_d = ((depth_d > 0) & torch.isfinite(depth_d)).to(torch.uint8)
        refined = self._apply(depth_d, mask_d, edge_d, ux, uy)
        scale = float(np.mean(uy / ux)) if np.all(ux > 0) else 1.0
        if self.verbose > 0:
            print(f"[FastPCHIP] Refined using {len(ux)} unique correspondences, effective scale {scale:.3f}")
        result.update(refined_depth=refined.cpu().numpy(), scale=scale, num_iterations=1)
        return result


def refine_depth_from_colmap(depth_map: np.ndarray, normal_map: np.ndarray | None, colmap_image, colmap_camera,
                             colmap_points3D: dict, **kwargs) -> dict[str, Any]:
    """Convenience wrapper on COLMAP objects (fast_pchip_refiner.py:585-605); works with ``colmap_io`` and with
    pycolmap objects (``cam_from_world`` may be a method or a property)."""
    K = colmap_camera.calibration_matrix()
    cfw = colmap_image.cam_from_world
    cam_from_world = (cfw() if callable(cfw) else cfw).matrix()
    ids = [p.point3D_id for p in colmap_image.points2D if p.has_point3D()]
    points3D = np.array([colmap_points3D[pid].xyz for pid in ids if pid in colmap_points3D]).reshape(-1, 3)
    return FastPCHIPRefiner(**kwargs).refine_depth(depth_map, normal_map, points3D, cam_from_world[:3], K)
