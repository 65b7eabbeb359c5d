"""The C-ABI library loads on a machine without a GPU and exports every symbol the header declares."""

import ctypes
import re
from pathlib import Path

import pytest

ROOT = Path(__file__).resolve().parent.parent


def test_exports_every_declared_symbol(lib_built):
    from depthdensifier_b200 import _lib

    header = (ROOT / "include" / "ddn_b200.h").read_text()
    declared = set(re.findall(r"\b(ddn_[a-z0-9_]+)\s*\(", header))
    declared -= {"ddn_view_stats"}
    assert declared, "no declarations parsed"
    lib = ctypes.CDLL(str(lib_built))
    for name in sorted(declared):
        assert hasattr(lib, name), f"{name} declared in include/ddn_b200.h but not exported"
    assert declared == set(_lib.SYMBOLS), declared ^ set(_lib.SYMBOLS)
    assert _lib.load().ddn_version() == 200


def test_argument_validation_without_gpu(lib_built):
    """Pure host-side checks: no kernel is launched for invalid arguments."""
    from depthdensifier_b200 import _lib

    lib = _lib.load()
    n = ctypes.c_int64(0)
    assert lib.ddn_align_workspace_bytes(4, 1000, ctypes.byref(n)) == 0 and n.value > 0
    assert lib.ddn_align_workspace_bytes(-1, 10, ctypes.byref(n)) == -1
    assert b"invalid argument" in lib.ddn_last_error_string()
    cfg = _lib.AlignConfig()
    lib.ddn_align_config_default(ctypes.byref(cfg))
    assert (cfg.min_correspondences, cfg.edge_margin, cfg.robust, cfg.max_pairs) == (50, 10, 1, 500)
    assert abs(cfg.outlier_threshold - 2.5) < 1e-7
    f = _lib.FilterConfig()
    lib.ddn_filter_config_default(ctypes.byref(f))
    assert abs(f.depth_threshold - 0.7) < 1e-7 and abs(f.grazing_cos - 0.087) < 1e-7 and f.stride == 1
    g = _lib.VoxelGrid()
    g.voxel = -1.0
    assert lib.ddn_voxel_fuse(ctypes.byref(g), 10, 0, None, None, None, 1, None, None, None, None, None, None, 0, None) == -1


def test_no_cpu_fallback():
    import numpy as np
    import torch

    from depthdensifier_b200 import DepthRefiner
    from depthdensifier_b200._lib import DDNError

    if torch.cuda.is_available():
        pytest.skip("GPU present")
    with pytest.raises(DDNError):
        DepthRefiner().refine_depth(np.ones((8, 8), np.float32), None, np.zeros((4, 3)), np.eye(4)[:3], np.eye(3))
