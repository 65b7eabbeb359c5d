"""Gradient masks and world-space normals next to the densification path.

``compute_depth_normal_gradient_mask`` keeps the signature of the reference function
(src/depthdensifier/initilizer.py:236-328); ``transform_normals`` is ``COLMAPVisualizer._transform_normals``
(src/depthdensifier/visualizer.py:346-376) as a free function.  Both run in sm_100a kernels (csrc/masks.cu); there
is no CPU path."""

from __future__ import annotations

import ctypes as C

import numpy as np
import torch

from . import _lib
from ._lib import DDNError


def _gaussian_taps(edge_sigma: float, dtype=torch.float32) -> torch.Tensor:
    """The reference's 1-D kernel, built with the same torch ops on the host (initilizer.py:264-273)."""
    kernel_size = int(2 * edge_sigma * 3) + 1
    if kernel_size % 2 == 0:
        kernel_size += 1
    sigma = torch.tensor(edge_sigma, dtype=dtype)
    x = torch.arange(kernel_size, dtype=dtype) - kernel_size // 2
    g = torch.exp(-0.5 * (x / sigma) ** 2)
    return g / g.sum()


def compute_depth_normal_gradient_mask(depth_map: torch.Tensor, normal_map: torch.Tensor | None = None,
                                       depth_threshold: float = 0.2, normal_threshold: float = 0.3,
                                       edge_sigma: float = 1.0) -> torch.Tensor:
    """Boolean mask [H, W], True where depth or normal gradients are high (regions to exclude)."""
    if not torch.cuda.is_available():
        raise DDNError("no CUDA device available: depthdensifier_b200 has no CPU fallback")
    lib = _lib.load()
    out_device = depth_map.device
    dev = depth_map.device if depth_map.is_cuda else torch.device("cuda")
    depth = depth_map.to(dev, torch.float32).contiguous()
    h, w = depth.shape[-2:]
    depth = depth.reshape(h, w)
    normal = None
    if normal_map is not None:
        if normal_map.dim() == 3 and normal_map.shape[0] == 3 and normal_map.shape[-1] != 3:
            normal_map = normal_map.permute(1, 2, 0)  # [3,H,W] -> [H,W,3] (initilizer.py:303-304)
        normal = normal_map.to(dev, torch.float32).contiguous()
    taps = _gaussian_taps(edge_sigma) if edge_sigma > 0 else torch.zeros(0)
    taps_c = (C.c_float * max(len(taps), 1))(*taps.tolist())
    mask = torch.empty((h, w), dtype=torch.uint8, device=dev)
    ws = torch.empty(2 * (h * w * 4 + 256) + 512, dtype=torch.uint8, device=dev)
    p = lambda t: C.c_void_p(t.data_ptr()) if t is not None else C.c_void_p(0)
    with torch.cuda.device(dev):
        _lib.check(lib.ddn_gradient_mask(h, w, p(depth), p(normal), taps_c, len(taps), float(np.float32(depth_threshold)),
                                         float(np.float32(normal_threshold)), p(mask), p(ws), ws.numel(),
                                         C.c_void_p(torch.cuda.current_stream().cuda_stream)))
    return mask.bool().to(out_device)


def transform_normals(normal_map: np.ndarray, cam_from_world: np.ndarray, mask: np.ndarray) -> np.ndarray:
    """(N, 3) float64 world-space unit normals of the masked pixels, in row-major pixel order."""
    if not torch.cuda.is_available():
        raise DDNError("no CUDA device available: depthdensifier_b200 has no CPU fallback")
    lib = _lib.load()
    valid = np.asarray(mask).reshape(-1) > 0
    n_cam = np.ascontiguousarray(np.asarray(normal_map, dtype=np.float32).reshape(-1, 3)[valid])
    pose = np.ascontiguousarray(cam_from_world, dtype=np.float64)
    if pose.shape[1] not in (3, 4) or pose.shape[0] < 3:
        raise ValueError("cam_from_world must be 3x3, 3x4 or 4x4")
    d_in = torch.from_numpy(n_cam).cuda()
    d_out = torch.empty((len(n_cam), 3), dtype=torch.float64, device="cuda")
    pose_c = (C.c_double * pose.size)(*pose.reshape(-1).tolist())
    _lib.check(lib.ddn_transform_normals(len(n_cam), C.c_void_p(d_in.data_ptr()), pose_c, int(pose.shape[1]),
                                         C.c_void_p(d_out.data_ptr()), C.c_void_p(torch.cuda.current_stream().cuda_stream)))
    return d_out.cpu().numpy()
