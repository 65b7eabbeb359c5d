"""TEST INFRASTRUCTURE ONLY - CPU restatement of the reference's alternative refiner ``FastPCHIPRefiner``
(src/depthdensifier/fast_pchip_refiner.py:69-579).  Pinned bit for bit against the
reference module itself by tests/test_oracle_vs_reference.py::test_pchip_restatement_matches_reference (the
reference imports ``pycolmap`` at module top, so it is loaded under the stand-in of oracle/standins.py).

Every function cites the reference lines it follows.  Arithmetic types are the reference's: numpy float64 for
the correspondence part, float32 for everything derived from the float32 depth map (NEP-50: Python scalars do
not promote float32 arrays), torch float32 for the Hermite evaluation.
"""

from __future__ import annotations

from dataclasses import dataclass

import numpy as np
import torch
from scipy.ndimage import binary_dilation, gaussian_filter, median_filter


@dataclass
class PchipConfig:  # fast_pchip_refiner.py:22-66
    min_correspondences: int = 100
    edge_margin: int = 20
    edge_threshold: float = 0.1
    edge_sigma: float = 2.0
    robust: bool = True
    outlier_threshold: float = 3.0
    use_image_edges: bool = False
    image_edge_threshold: float = 30.0


def project_points_to_image(points3D, cam_from_world, K):
    """:142-153 - float64; valid = depth > 1e-6; full 3x3 K on the normalised point."""
    hom = np.hstack([points3D, np.ones((points3D.shape[0], 1))])
    H = np.vstack([cam_from_world, [0, 0, 0, 1]])
    cam = (H @ hom.T)[:3, :].T
    depths = cam[:, 2].copy()
    valid = depths > 1e-6
    pts2d = np.zeros((len(depths), 2))
    pts2d[valid] = ((K @ (cam[valid] / depths[valid, None]).T).T)[:, :2]
    return pts2d, depths, valid


def sample_depth_at_points(depth_map, pts2d):
    """:155-185 - bilinear with clipped taps; note the weights use the CLIPPED x1/y1."""
    h, w = depth_map.shape
    x, y = pts2d[:, 0], pts2d[:, 1]
    x0 = np.floor(x).astype(int)
    x1 = x0 + 1
    y0 = np.floor(y).astype(int)
    y1 = y0 + 1
    x0, x1 = np.clip(x0, 0, w - 1), np.clip(x1, 0, w - 1)
    y0, y1 = np.clip(y0, 0, h - 1), np.clip(y1, 0, h - 1)
    wa, wb = (x1 - x) * (y1 - y), (x - x0) * (y1 - y)
    wc, wd = (x1 - x) * (y - y0), (x - x0) * (y - y0)
    return wa * depth_map[y0, x0] + wb * depth_map[y0, x1] + wc * depth_map[y1, x0] + wd * depth_map[y1, x1]


def detect_image_edges(rgb_image, mask, cfg: PchipConfig):
    """:187-224 - grey = .299R + .587G + .114B, Gaussian(sigma), np.gradient magnitude > threshold, & mask,
    two dilations."""
    gray = 0.299 * rgb_image[:, :, 0] + 0.587 * rgb_image[:, :, 1] + 0.114 * rgb_image[:, :, 2] if rgb_image.ndim == 3 else rgb_image
    smooth = gaussian_filter(gray, sigma=cfg.edge_sigma)
    dy, dx = np.gradient(smooth)
    edge = np.sqrt(dx**2 + dy**2) > cfg.image_edge_threshold
    if mask is not None:
        edge = edge & mask
    return binary_dilation(edge, iterations=2)


def detect_depth_edges(depth_map, mask, normal_map, cfg: PchipConfig):
    """:226-273 - normal-gradient magnitude > 0.3, OR relative gradient of the Gaussian-smoothed depth >
    edge_threshold (zeroed outside the mask), then two dilations (cross structuring element)."""
    h, w = depth_map.shape
    if mask is None:
        mask = np.ones((h, w), dtype=bool)
    edge = np.zeros((h, w), dtype=bool)
    if normal_map is not None and normal_map.shape[-1] == 3:
        gx = np.gradient(normal_map[..., 0], axis=[0, 1])
        gy = np.gradient(normal_map[..., 1], axis=[0, 1])
        gz = np.gradient(normal_map[..., 2], axis=[0, 1])
        mag = np.sqrt(gx[0]**2 + gx[1]**2 + gy[0]**2 + gy[1]**2 + gz[0]**2 + gz[1]**2)
        edge |= mag > 0.3
    smooth = gaussian_filter(depth_map, sigma=cfg.edge_sigma)
    dy, dx = np.gradient(smooth)
    gmag = np.sqrt(dx**2 + dy**2)
    with np.errstate(divide="ignore", invalid="ignore"):
        rel = gmag / (smooth + 1e-6)
        rel[~mask] = 0
    edge |= rel > cfg.edge_threshold
    return binary_dilation(edge, iterations=2)


def remove_outliers(z_colmap, z_depth, cfg: PchipConfig):
    """:275-298 - MAD test on z_colmap - z_depth * median(z_colmap / (z_depth + 1e-6)); strict <."""
    ratios = z_colmap / (z_depth + 1e-6)
    med_ratio = np.median(ratios)
    res = z_colmap - z_depth * med_ratio
    mad = np.median(np.abs(res - np.median(res)))
    if mad < 1e-6:
        mad = np.std(res) * 0.6745
    inl = np.abs(res - np.median(res)) < cfg.outlier_threshold * mad
    return z_colmap[inl], z_depth[inl], int((~inl).sum())


def hermite_eval(d, x, y):
    """:300-366 - torch float32 cubic Hermite with 0.3-scaled secant tangents.  ``slope_left[0]`` /
    ``slope_right[-1]`` in the extrapolation branches (:352-358) are the tangents of the FIRST / LAST QUERY, not
    of the first / last knot - reproduced as written."""
    xt, yt, dt = torch.from_numpy(x).float(), torch.from_numpy(y).float(), torch.from_numpy(d).float()
    slopes = (yt[1:] - yt[:-1]) / ((xt[1:] - xt[:-1]) + 1e-8)
    padded = torch.cat([slopes[:1], slopes, slopes[-1:]])
    idx = torch.clamp(torch.searchsorted(xt, dt), 1, len(xt) - 1)
    il, ir = idx - 1, idx
    xl, xr, yl, yr = xt[il], xt[ir], yt[il], yt[ir]
    sl, sr = padded[il], padded[ir]
    h = xr - xl + 1e-8
    t = (dt - xl) / h
    t2 = t * t
    t3 = t2 * t
    h00, h10, h01, h11 = 2 * t3 - 3 * t2 + 1, t3 - 2 * t2 + t, -2 * t3 + 3 * t2, t3 - t2
    res = h00 * yl + h10 * h * sl * 0.3 + h01 * yr + h11 * h * sr * 0.3
    res = torch.where(dt <= xt[0], yt[0] + sl[0] * (dt - xt[0]) * 0.3, res)
    res = torch.where(dt >= xt[-1], yt[-1] + sr[-1] * (dt - xt[-1]) * 0.3, res)
    return torch.maximum(res, torch.tensor(1e-3)).numpy()


def apply_pchip(depth_map, sel, x, y):
    """:368-385 - copy of the depth map with the selected pixels transformed."""
    q = depth_map[sel]
    if len(q) == 0:
        return depth_map
    out = depth_map.copy()
    out[sel] = hermite_eval(q, x, y)
    return out


def apply_edge_aware(depth_map, mask, edge_mask, x, y):
    """:550-579 - non-edge pixels fully transformed, edge pixels 0.7 original + 0.3 transformed, 3x3 median
    (scipy default boundary 'reflect'), zero outside the mask.  Pixels outside the mask keep their ORIGINAL depth
    until the final masking, so they take part in the median."""
    refined = np.zeros_like(depth_map)
    non_edge = mask & ~edge_mask
    if non_edge.any():
        refined = apply_pchip(depth_map, non_edge, x, y)
    edge_valid = mask & edge_mask
    if edge_valid.any():
        tr = apply_pchip(depth_map, edge_valid, x, y)
        refined[edge_valid] = 0.7 * depth_map[edge_valid] + 0.3 * tr[edge_valid]
    refined = median_filter(refined, size=3)
    refined[~mask] = 0
    return refined


def correspondences(depth_map, edge_mask, points3D, cam_from_world, K, cfg: PchipConfig):
    """:429-476 - project, margin gate, bilinear sample, drop non-positive / non-finite / edge pixels (nearest
    pixel by np.round = half to even).  Returns (z_depth, z_colmap) or None when nothing is in bounds."""
    h, w = depth_map.shape
    pts2d, depths3d, valid = project_points_to_image(points3D, cam_from_world, K)
    m = cfg.edge_margin
    inb = (pts2d[:, 0] >= m) & (pts2d[:, 0] < w - m) & (pts2d[:, 1] >= m) & (pts2d[:, 1] < h - m) & valid
    p, z3 = pts2d[inb], depths3d[inb]
    if len(p) == 0:
        return None
    samp = sample_depth_at_points(depth_map, p)
    u, v = np.round(p[:, 0]).astype(int), np.round(p[:, 1]).astype(int)
    ok = (samp > 0) & (z3 > 0) & np.isfinite(samp) & ~edge_mask[v, u]
    return samp[ok], z3[ok]


def refine_depth(depth_map, normal_map, points3D, cam_from_world, K, mask=None, rgb_image=None, cfg: PchipConfig = PchipConfig()):
    """:386-548."""
    use_normals = normal_map is not None
    out = {"energy_history": [], "num_iterations": 0, "used_normals": use_normals, "scale": 1.0, "refined_depth": depth_map}
    if cfg.use_image_edges and rgb_image is not None:
        edge_mask = detect_image_edges(rgb_image, mask, cfg)
    else:
        edge_mask = detect_depth_edges(depth_map, mask, normal_map, cfg)
    out["edge_mask"] = edge_mask
    corr = correspondences(depth_map, edge_mask, points3D, cam_from_world, K, cfg)
    if corr is None:
        return out
    z_depth, z_colmap = corr
    if len(z_depth) < cfg.min_correspondences:
        return out
    zc, zd = (remove_outliers(z_colmap, z_depth, cfg)[:2]) if cfg.robust else (z_colmap, z_depth)
    if len(zd) < cfg.min_correspondences:  # :505-515
        scale = float(np.median(z_colmap / z_depth)) if np.all(z_depth > 0) else 1.0
        out.update(refined_depth=depth_map * scale, scale=scale)
        return out
    ux, ui = np.unique(zd, return_index=True)  # :518-519
    uy = zc[ui]
    if mask is None:
        mask = (depth_map > 0) & np.isfinite(depth_map)
    out.update(refined_depth=apply_edge_aware(depth_map, mask, edge_mask, ux, uy),
               scale=float(np.mean(uy / ux)) if np.all(ux > 0) else 1.0, num_iterations=1, knots_x=ux, knots_y=uy)
    return out
