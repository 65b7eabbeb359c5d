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


def _header_struct_fields():
    """{struct name: [field names in order]} parsed from include/ddn_b200.h (plain C structs, one field per line)."""
    text = (ROOT / "include" / "ddn_b200.h").read_text()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    out = {}
    for body, name in re.findall(r"typedef struct[^{]*\{(.*?)\}\s*(ddn_\w+)\s*;", text, flags=re.S):
        out[name] = [re.match(r".*?(\w+)\s*(\[\d+\])?\s*$", f.strip()).group(1) for f in body.split(";") if f.strip()]
    return out


def test_ctypes_structs_match_the_header(tmp_path):
    """Every struct of include/ddn_b200.h, compiled as C by gcc, has the size and the field offsets of its ctypes mirror
    in depthdensifier_b200/_lib.py (a field added on one side only would shift everything behind it)."""
    import shutil
    import subprocess

    from depthdensifier_b200 import _lib

    gcc = shutil.which("gcc")
    if gcc is None:
        pytest.skip("gcc not available")
    mirrors = {"ddn_align_config": _lib.AlignConfig, "ddn_view_stats": _lib.ViewStats, "ddn_filter_config": _lib.FilterConfig,
               "ddn_voxel_grid": _lib.VoxelGrid, "ddn_grid_state": _lib.GridState, "ddn_fuse_session": _lib.FuseSession}
    fields = _header_struct_fields()
    assert set(fields) == set(mirrors), "a struct of the header has no ctypes mirror (or the other way round)"
    lines = ['#include <stddef.h>', '#include <stdio.h>', f'#include "{ROOT / "include" / "ddn_b200.h"}"', "int main(void) {"]
    for s, names in fields.items():
        lines.append(f'  printf("{s} %zu\\n", sizeof({s}));')
        lines += [f'  printf("{s}.{n} %zu\\n", offsetof({s}, {n}));' for n in names]
    lines += ["  return 0;", "}"]
    src = tmp_path / "layout.c"
    src.write_text("\n".join(lines))
    exe = tmp_path / "layout"
    subprocess.run([gcc, "-std=c11", "-Wall", "-Werror", "-o", str(exe), str(src)], check=True)  # the header is plain C
    got = dict(line.split() for line in subprocess.run([str(exe)], check=True, capture_output=True, text=True).stdout.splitlines())
    for s, cls in mirrors.items():
        assert int(got[s]) == ctypes.sizeof(cls), s
        assert [n for n, *_ in cls._fields_] == fields[s], s
        for n in fields[s]:
            assert int(got[f"{s}.{n}"]) == getattr(cls, n).offset, (s, n)


def test_integration_stub_lists_every_align_config_field():
    """The ctypes stub shown to the reference's maintainers declares ddn_align_config field for field."""
    from depthdensifier_b200 import _lib

    text = (ROOT / "INTEGRATION.md").read_text()
    block = text[text.index("class AlignConfig(C.Structure)"):]
    block = block[:block.index("lib.ddn_align_views.restype")]
    names = re.findall(r'\("(\w+)", C\.c_\w+\)', block)
    assert names == [n for n, *_ in _lib.AlignConfig._fields_]
