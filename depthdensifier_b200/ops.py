"""Device-tensor wrappers around the C ABI.  torch is plumbing only (device memory + streams); every
op below runs a hand-written sm_100a kernel from libddn_b200.so and raises when no GPU is present."""

from __future__ import annotations

import ctypes as C
import math
from dataclasses import dataclass

import numpy as np
import torch

from . import _lib
from ._lib import DDNError

STATUS_REFINED = 0
STATUS_NAMES = {
    0: "refined",
    1: "no_points_in_bounds",
    2: "no_positive_samples",
    3: "too_few_correspondences",
    4: "degenerate_fit",
    5: "no_sparse_points",
}


def _require_cuda(*tensors) -> torch.device:
    if not torch.cuda.is_available():
        raise DDNError("no CUDA device available: depthdensifier_b200 has no CPU fallback")
    dev = None
    for t in tensors:
        if t is None:
            continue
        if not t.is_cuda:
            raise DDNError("expected CUDA tensors")
        if not t.is_contiguous():
            raise DDNError("expected contiguous tensors")
        dev = t.device if dev is None else dev
        if t.device != dev:
            raise DDNError("tensors on different devices")
    return dev


def _p(t):
    return C.c_void_p(t.data_ptr()) if t is not None else C.c_void_p(0)


def _stream():
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


@dataclass
class AlignOptions:
    """RefinerConfig fields (depth_refiner.py:16-31) + the north-star mode switch."""

    min_correspondences: int = 50
    edge_margin: int = 10
    robust: bool = True
    outlier_threshold: float = 2.5
    skip_smoothing: bool = False
    adaptive_correspondences: bool = True
    max_pairs: int = 500
    align_mode: str = "pwl"
    subsample_seed: int = 0
    zero_unmasked_passthrough: bool = False
    use_tma: bool = False  # K3 depth tile via one TMA tensor copy (W % 4 == 0, W >= 132, H >= 34): same bits, measured 6 % slower

    def to_c(self, mask_packed: bool = False) -> _lib.AlignConfig:
        if self.align_mode not in ("pwl", "affine"):
            raise ValueError("align_mode must be 'pwl' or 'affine'")
        return _lib.AlignConfig(
            int(self.min_correspondences),
            int(self.edge_margin),
            int(bool(self.robust)),
            float(self.outlier_threshold),
            int(bool(self.skip_smoothing)),
            int(bool(self.adaptive_correspondences)),
            int(self.max_pairs),
            0 if self.align_mode == "pwl" else 1,
            int(self.subsample_seed) & 0xFFFFFFFF,
            int(bool(self.zero_unmasked_passthrough)),
            int(bool(mask_packed)),
            int(bool(self.use_tma)),
        )


@dataclass
class FilterOptions:
    depth_threshold: float = 0.7
    grazing_cos: float = 0.087
    sample_mode: str = "nearest"
    two_sided_tau: float = 0.0
    stride: int = 1
    normals_in_world: bool = False
    pixel_layout: int = 0  # which 4 pixels a K4 thread owns: 0 = 32 apart, 1 = adjacent (same results)

    def to_c(self) -> _lib.FilterConfig:
        if self.sample_mode not in ("nearest", "bilinear"):
            raise ValueError("sample_mode must be 'nearest' or 'bilinear'")
        return _lib.FilterConfig(
            float(self.depth_threshold),
            float(self.grazing_cos),
            0 if self.sample_mode == "nearest" else 1,
            float(self.two_sided_tau),
            int(self.stride),
            int(bool(self.normals_in_world)),
            int(self.pixel_layout),
        )


def align_views(depth, mask, cam_from_world, kmat, sparse_xyz, sparse_offsets, max_sparse_per_view: int,
                opts: AlignOptions, out=None, src_table=None, bbox=None):
    """Stage 1 for V views.  Returns (refined [V,H,W] f32, stats [V,8] int32 raw ddn_view_stats).
    ``src_table`` [V,16] (from build_pair_tables) + ``bbox`` [6] (new_bbox): the kernel also extends the box to
    enclose the back-projection of every refined pixel, so the voxel grid is known before stages 2+3."""
    lib = _lib.load()
    dev = _require_cuda(depth, mask, cam_from_world, kmat, sparse_xyz, sparse_offsets, out, src_table, bbox)
    assert (src_table is None) == (bbox is None)
    assert src_table is None or (src_table.dtype == torch.float32 and tuple(src_table.shape) == (depth.shape[0], 16))
    V, H, W = depth.shape
    assert depth.dtype == torch.float32 and cam_from_world.dtype == torch.float64 and kmat.dtype == torch.float64
    assert sparse_xyz.dtype == torch.float64 and sparse_offsets.dtype == torch.int64
    # a 2-D uint8 mask [V, ceil(H*W/8)] is the bit-packed form (pack_mask)
    packed = mask is not None and mask.dtype == torch.uint8 and mask.dim() == 2
    assert mask is None or (packed and tuple(mask.shape) == (V, (H * W + 7) // 8)) or (
        mask.dtype in (torch.bool, torch.uint8) and mask.shape == depth.shape)
    assert tuple(cam_from_world.shape) == (V, 3, 4) and tuple(kmat.shape) == (V, 3, 3)
    refined = out if out is not None else torch.empty_like(depth)
    stats = torch.zeros((V, 8), dtype=torch.int32, device=dev)
    nbytes = C.c_int64(0)
    _lib.check(lib.ddn_align_workspace_bytes(V, max_sparse_per_view, C.byref(nbytes)))
    ws = torch.empty(nbytes.value, dtype=torch.uint8, device=dev)
    cfg = opts.to_c(mask_packed=packed)
    with torch.cuda.device(dev):
        _lib.check(
            lib.ddn_align_views(
                C.byref(cfg), V, H, W, _p(depth), _p(mask), _p(cam_from_world), _p(kmat), _p(sparse_xyz),
                _p(sparse_offsets), int(max_sparse_per_view), _p(refined), _p(stats), _p(ws), nbytes.value, _p(src_table),
                _p(bbox), _stream(),
            )
        )
    return refined, stats


def pack_mask(mask: torch.Tensor) -> torch.Tensor:
    """[V,H,W] bool mask -> [V, ceil(H*W/8)] uint8, one bit per pixel (bit g & 7 of byte g >> 3): the form the alignment
    kernel also accepts, an eighth of the bytes to upload.  Host tensors (numpy packbits)."""
    m = np.ascontiguousarray(mask.cpu().numpy()).reshape(mask.shape[0], -1)
    return torch.from_numpy(np.packbits(m, axis=1, bitorder="little"))


def decode_stats(stats: torch.Tensor) -> list[dict]:
    """Host dicts from the raw [V,8] int32 stats tensor (synchronises)."""
    s = stats.cpu().numpy()
    f = s.view(np.float32)
    out = []
    for v in range(s.shape[0]):
        out.append(
            {
                "status": int(s[v, 0]),
                "num_correspondences": int(s[v, 1]),
                "outliers_removed": int(s[v, 2]),
                "num_table": int(s[v, 3]),
                "scale_factor": float(f[v, 4]),
                "affine_scale": float(f[v, 5]),
                "affine_shift": float(f[v, 6]),
            }
        )
    return out


def build_pair_tables(cam_from_world, intr, nbr, src_begin: int, n_src: int, height: int, width: int):
    """Per-(source view, neighbour) float32 tables from the float64 poses; ``height`` / ``width``: the size of the
    depth maps the tables will be used with (K4's gather offsets are precomputed per entry)."""
    lib = _lib.load()
    dev = _require_cuda(cam_from_world, intr, nbr)
    V = cam_from_world.shape[0]
    K = nbr.shape[1]
    assert cam_from_world.dtype == torch.float64 and intr.dtype == torch.float64 and nbr.dtype == torch.int32
    assert tuple(intr.shape) == (V, 4) and nbr.shape[0] == V
    pair = torch.empty((n_src, K, _lib.PAIR_TABLE_FLOATS), dtype=torch.float32, device=dev)
    src = torch.empty((n_src, 16), dtype=torch.float32, device=dev)
    with torch.cuda.device(dev):
        _lib.check(lib.ddn_build_pair_tables(V, src_begin, n_src, K, int(height), int(width), _p(cam_from_world), _p(intr), _p(nbr),
                                             _p(pair), _p(src), _stream()))
    return pair, src


def init_bbox(bbox: torch.Tensor) -> torch.Tensor:
    """(+inf, -inf) in the order-preserving int encoding the kernels fold points into."""
    lib = _lib.load()
    with torch.cuda.device(bbox.device):
        _lib.check(lib.ddn_bbox_init(_p(bbox), _stream()))
    return bbox


def new_bbox(dev) -> torch.Tensor:
    return init_bbox(torch.empty(6, dtype=torch.int32, device=dev))


def decode_bbox(bbox: torch.Tensor) -> np.ndarray:
    """[6] float32 (min xyz, max xyz) from the order-preserving int encoding (synchronises)."""
    i = bbox.cpu().numpy().astype(np.int32)
    i = np.where(i >= 0, i, i ^ np.int32(0x7FFFFFFF))
    return i.view(np.float32)


def backproject_filter(refined_all, normal, nbr, pair_table, src_table, src_begin: int, vote_threshold: int,
                       opts: FilterOptions, bbox=None, xyz_out=None, votes_out=None, mark=None):
    """Stages 2+3 for the source views src_begin..src_begin+n_src (n_src = normal.shape[0]).  ``normal`` may
    be a PINNED HOST tensor: the kernel then reads the normals of its vote candidates in place over PCIe
    (unified addressing) and the 12 B/pixel normal map never moves to the device.  ``mark``: an open
    ``FuseSession``; the kernel then also sets the occupancy bits of the kept points (stage 4's mark pass)."""
    lib = _lib.load()
    normal_on_host = not normal.is_cuda
    if normal_on_host and not (normal.is_pinned() and normal.is_contiguous()):
        raise DDNError("a host normal map must be pinned and contiguous")
    dev = _require_cuda(refined_all, None if normal_on_host else normal, nbr, pair_table, src_table, bbox, xyz_out, votes_out)
    V, H, W = refined_all.shape
    n_src = normal.shape[0]
    K = nbr.shape[1]
    assert refined_all.dtype == torch.float32 and normal.dtype == torch.float32
    assert tuple(normal.shape) == (n_src, H, W, 3)
    assert tuple(pair_table.shape) == (n_src, K, _lib.PAIR_TABLE_FLOATS) and tuple(src_table.shape) == (n_src, 16)
    s = int(opts.stride)
    Hs, Ws = (H + s - 1) // s, (W + s - 1) // s
    xyz = xyz_out if xyz_out is not None else torch.empty((n_src, Hs, Ws, 3), dtype=torch.float32, device=dev)
    votes = votes_out if votes_out is not None else torch.empty((n_src, Hs, Ws), dtype=torch.uint8, device=dev)
    cfg = opts.to_c()
    with torch.cuda.device(dev):
        _lib.check(
            lib.ddn_backproject_filter(
                C.byref(cfg), V, src_begin, n_src, H, W, K, _p(refined_all), _p(normal), _p(nbr), _p(pair_table),
                _p(src_table), int(vote_threshold), _p(xyz), _p(votes), _p(bbox),
                C.byref(mark.c) if mark is not None else None, _stream(),
            )
        )
    return xyz, votes


def make_grid(bbox_min, bbox_max, voxel: float):
    """Voxel grid from a bounding box: origin = floor(min/voxel)*voxel in float32 and the
    number of significant bits per axis."""
    v = np.float32(voxel)
    lo = np.asarray(bbox_min, dtype=np.float32)
    hi = np.asarray(bbox_max, dtype=np.float32)
    origin = (np.floor(lo / v) * v).astype(np.float32)
    cells = np.floor((hi - origin) / v).astype(np.int64) + 2
    bits = [max(1, int(math.ceil(math.log2(max(int(c), 2))))) for c in cells]
    if max(bits) > 21:
        raise DDNError(f"voxel grid needs {bits} bits per axis; the 3x21-bit key allows at most 21")
    g = _lib.VoxelGrid()
    g.voxel = float(v)
    for i in range(3):
        g.origin[i] = float(origin[i])
        g.bits[i] = bits[i]
        g.dims[i] = int(cells[i])
    return g


def checked_voxel_count(counts: torch.Tensor) -> int:
    """Number of voxels from the device counts [2] (synchronises); raises on the colour-sum overflow flag."""
    n_pts, mv = (int(x) for x in counts.cpu().tolist())
    if n_pts < 0:
        raise DDNError("voxel fusion: a voxel collected 2^24 or more points (32-bit colour sums); use a smaller voxel")
    return mv


GRID_STATUS = {0: "ok", 1: "empty (no finite bounding box)", 2: "grid larger than the session's capacity",
               3: "an axis needs more than 21 bits"}


class FuseSession:
    """Device buffers of one fusion session (include/ddn_b200.h: ddn_fuse_session): the grid lives on the device,
    derived there from bounding boxes, so a step runs without the host looking at an intermediate result.
    ``alloc(name, nbytes)`` may supply a buffer from elsewhere (NVLink-visible symmetric memory for ``units`` and
    ``tile_prefix`` in the multi-GPU path); every buffer is otherwise a plain device tensor."""

    def __init__(self, device, max_cells: int = 1 << 33, tile_prefix: bool = False, dirty: bool = True, alloc=None):
        lib = _lib.load()
        if not torch.cuda.is_available():
            raise DDNError("no CUDA device available: depthdensifier_b200 has no CPU fallback")
        self.device = torch.device(device)
        sizes = [C.c_int64(0) for _ in range(5)]
        _lib.check(lib.ddn_fuse_session_sizes(int(max_cells), *[C.byref(x) for x in sizes]))
        cap_units, units_b, dirty_b, sums_b, prefix_b = (x.value for x in sizes)
        self.max_cells, self.cap_units = int(max_cells), cap_units
        self.n_own_cap = cap_units // 256 + 2

        def get(name, nbytes):
            t = alloc(name, nbytes) if alloc is not None else None
            if t is None:
                t = torch.empty(nbytes, dtype=torch.uint8, device=self.device)
            assert t.is_cuda and t.is_contiguous() and t.numel() * t.element_size() >= nbytes and t.data_ptr() % 16 == 0
            return t

        self.units = get("units", units_b)
        self.dirty = get("dirty", dirty_b) if dirty else None
        self.tile_sums = get("tile_sums", sums_b)
        self.tile_prefix = get("tile_prefix", prefix_b) if tile_prefix else None
        self.grid = torch.zeros(16, dtype=torch.int32, device=self.device)
        self.counts = torch.zeros(2, dtype=torch.int64, device=self.device)
        self.c = _lib.FuseSession(self.grid.data_ptr(), self.units.data_ptr(), cap_units,
                                  self.dirty.data_ptr() if self.dirty is not None else None, self.tile_sums.data_ptr(),
                                  self.tile_prefix.data_ptr() if self.tile_prefix is not None else None, self.counts.data_ptr())
        self._accum = None
        with torch.cuda.device(self.device):
            _lib.check(lib.ddn_fuse_session_reset(C.byref(self.c), _stream()))

    def begin(self, boxes, voxel: float) -> None:
        """Opens a step on the union of ``boxes``: device tensors [6] (new_bbox encoding) or raw device addresses
        (peer memory)."""
        lib = _lib.load()
        ptrs = [b if isinstance(b, int) else b.data_ptr() for b in boxes]
        arr = (C.c_void_p * len(ptrs))(*ptrs)
        with torch.cuda.device(self.device):
            _lib.check(lib.ddn_fuse_begin(C.byref(self.c), arr, len(ptrs), float(np.float32(voxel)), _stream()))

    def begin_grid(self, grid: _lib.VoxelGrid) -> None:
        lib = _lib.load()
        with torch.cuda.device(self.device):
            _lib.check(lib.ddn_fuse_begin_grid(C.byref(self.c), C.byref(grid), _stream()))

    def grid_state(self) -> _lib.GridState:
        """Host copy of the device grid (synchronises)."""
        raw = self.grid.cpu().numpy().tobytes()
        return _lib.GridState.from_buffer_copy(raw)

    def host_grid(self) -> _lib.VoxelGrid:
        """The device grid as the host struct the host-grid entry points and the tests take (synchronises);
        raises when the session could not build one."""
        st = self.grid_state()
        if st.status != _lib.GRID_OK:
            raise DDNError(f"fusion grid: {GRID_STATUS.get(st.status, st.status)} (dims {list(st.dims)}, "
                           f"capacity {self.max_cells} cells); raise max_grid_cells or use a larger voxel")
        g = _lib.VoxelGrid()
        g.voxel = st.voxel
        for i in range(3):
            g.origin[i], g.bits[i], g.dims[i] = st.origin[i], st.bits[i], st.dims[i]
        return g

    def merge_scratch(self, world: int, cap_out: int) -> torch.Tensor:
        """Scratch of fuse_merge_peers for ``world`` ranks and ``cap_out`` records (allocated once per session)."""
        if getattr(self, "_merge_scratch", None) is None or self._merge_scratch[0] != (world, cap_out):
            nbytes = C.c_int64(0)
            _lib.check(_lib.load().ddn_fuse_merge_scratch_bytes(self.cap_units, int(world), int(cap_out), C.byref(nbytes)))
            self._merge_scratch = ((world, cap_out), torch.empty(nbytes.value, dtype=torch.uint8, device=self.device))
        return self._merge_scratch[1]

    def accum(self, cap_out: int) -> torch.Tensor:
        nbytes = cap_out * 40 + 16
        if self._accum is None or self._accum.numel() < nbytes:
            self._accum = torch.empty(nbytes, dtype=torch.uint8, device=self.device)
        return self._accum

    def mark_points(self, xyz, votes, vote_threshold: int) -> None:
        lib = _lib.load()
        _require_cuda(xyz, votes)
        with torch.cuda.device(self.device):
            _lib.check(lib.ddn_fuse_mark_points(C.byref(self.c), xyz.shape[0], _p(xyz), _p(votes), int(vote_threshold), _stream()))


    def unmark_points(self, xyz) -> None:
        """N5: the cells of these points (xyz [N,3] f32, e.g. the sparse cloud) leave the occupancy, so no dense
        voxel is created there.  Between the mark and fuse_finish*."""
        lib = _lib.load()
        _require_cuda(xyz)
        assert xyz.dtype == torch.float32 and xyz.dim() == 2 and xyz.shape[1] == 3
        with torch.cuda.device(self.device):
            _lib.check(lib.ddn_fuse_unmark_points(C.byref(self.c), xyz.shape[0], _p(xyz), _stream()))


def new_voxel_outputs(cap_out: int, dev):
    return (torch.empty(cap_out, dtype=torch.int64, device=dev), torch.empty((cap_out, 3), dtype=torch.float32, device=dev),
            torch.empty((cap_out, 3), dtype=torch.uint8, device=dev), torch.empty(cap_out, dtype=torch.int32, device=dev))


def fuse_finish(sess: FuseSession, xyz, rgb, votes, vote_threshold: int, row_len: int = 0, cap_out: int | None = None, out=None):
    """Rank + accumulate + finalise of the points marked in ``sess``.  Returns keys, xyz, rgb, count (capacity
    ``cap_out``, default = the number of points) and the session's device counts [2] = (points, voxels)."""
    lib = _lib.load()
    dev = _require_cuda(xyz, rgb, votes)
    N = xyz.shape[0]
    assert xyz.dtype == torch.float32 and rgb.dtype == torch.uint8 and tuple(rgb.shape) == (N, 3)
    cap = max(int(cap_out if cap_out is not None else N), 1)
    k, x, c, n = out if out is not None else new_voxel_outputs(cap, dev)
    acc = sess.accum(cap)
    with torch.cuda.device(dev):
        _lib.check(lib.ddn_fuse_finish(C.byref(sess.c), N, int(row_len), _p(xyz), _p(rgb), _p(votes), int(vote_threshold), _p(k), _p(x),
                                       _p(c), _p(n), cap, _p(acc), acc.numel(), _stream()))
    return k, x, c, n, sess.counts


def fuse_finish_partial(sess: FuseSession, xyz, rgb, votes, vote_threshold: int, records, row_len: int = 0):
    """Rank + accumulate into partial ``records`` [cap, 6] i64 (and the session's tile prefix)."""
    lib = _lib.load()
    dev = _require_cuda(xyz, rgb, votes, records)
    N = xyz.shape[0]
    assert records.dtype == torch.int64 and records.shape[1] == _lib.RECORD_WORDS
    with torch.cuda.device(dev):
        _lib.check(lib.ddn_fuse_finish_partial(C.byref(sess.c), N, int(row_len), _p(xyz), _p(rgb), _p(votes), int(vote_threshold),
                                               _p(records), records.shape[0], _stream()))
    return sess.counts


def fuse_merge_peers(sess: FuseSession, rank: int, world: int, peer_records, peer_tile_prefix, plan, cap_out: int, out=None,
                     drop_xyz=None):
    """Owner-side exchange + merge over peer memory.  ``peer_*``: per rank, the device address of that rank's
    records / tile prefix as mapped into this process.  ``drop_xyz`` [n,3] f32: N5, the sparse points of
    ALL ranks whose cells are removed from the merged occupancy.  Returns keys, xyz, rgb, count, counts."""
    lib = _lib.load()
    dev = sess.device
    k, x, c, n = out if out is not None else new_voxel_outputs(cap_out, dev)
    acc = sess.accum(cap_out)
    scratch = sess.merge_scratch(world, cap_out)
    arr = lambda ptrs: (C.c_void_p * world)(*[int(v) for v in ptrs])
    with torch.cuda.device(dev):
        _lib.check(lib.ddn_fuse_merge_peers(C.byref(sess.c), int(rank), int(world), arr(peer_records), arr(peer_tile_prefix),
                                            _p(plan), _p(scratch), scratch.numel(), _p(drop_xyz),
                                            0 if drop_xyz is None else drop_xyz.shape[0], _p(k), _p(x), _p(c), _p(n), int(cap_out),
                                            _p(acc), acc.numel(), _stream()))
    return k, x, c, n, sess.counts


def voxel_fuse(xyz, rgb, votes, vote_threshold: int, grid: _lib.VoxelGrid, trim: bool = True, row_len: int = 0):
    """Stage 4.  xyz [N,3] f32, rgb [N,3] u8, votes [N] u8 or None.  Returns keys, xyz, rgb, count
    (trimmed to the voxel count when ``trim``; that reads the counts back and synchronises) and the
    device counts tensor [2] = (participating points, voxels).  ``row_len``: image width when the points
    are the pixels of row-major depth maps (locality hint only)."""
    lib = _lib.load()
    dev = _require_cuda(xyz, rgb, votes)
    N = xyz.shape[0]
    assert xyz.dtype == torch.float32 and rgb.dtype == torch.uint8 and tuple(rgb.shape) == (N, 3)
    assert votes is None or (votes.dtype == torch.uint8 and votes.numel() == N)
    nbytes = C.c_int64(0)
    _lib.check(lib.ddn_fuse_workspace_bytes(C.byref(grid), N, C.byref(nbytes)))
    ws = torch.empty(nbytes.value, dtype=torch.uint8, device=dev)
    out_keys = torch.empty(N, dtype=torch.int64, device=dev)
    out_xyz = torch.empty((N, 3), dtype=torch.float32, device=dev)
    out_rgb = torch.empty((N, 3), dtype=torch.uint8, device=dev)
    out_cnt = torch.empty(N, dtype=torch.int32, device=dev)
    counts = torch.zeros(2, dtype=torch.int64, device=dev)
    with torch.cuda.device(dev):
        _lib.check(
            lib.ddn_voxel_fuse(
                C.byref(grid), N, int(row_len), _p(xyz), _p(rgb), _p(votes), int(vote_threshold), _p(out_keys), _p(out_xyz),
                _p(out_rgb), _p(out_cnt), _p(counts), _p(ws), nbytes.value, _stream(),
            )
        )
    if trim:
        mv = checked_voxel_count(counts)
        return out_keys[:mv], out_xyz[:mv], out_rgb[:mv], out_cnt[:mv], counts
    return out_keys, out_xyz, out_rgb, out_cnt, counts


def fuse_tile_info(grid: _lib.VoxelGrid) -> tuple[int, int]:
    """(number of ownership tiles, cells per tile); 0 tiles = grid too large for the dense path."""
    lib = _lib.load()
    nt, cpt = C.c_int64(0), C.c_int64(0)
    _lib.check(lib.ddn_fuse_tile_info(C.byref(grid), C.byref(nt), C.byref(cpt)))
    return nt.value, cpt.value


def voxel_fuse_partial(xyz, rgb, votes, vote_threshold: int, grid: _lib.VoxelGrid, row_len: int = 0, tile_prefix=None,
                       out=None):
    """Rank-local stage 4: per-voxel partial RECORDS, keys ascending (layout: include/ddn_b200.h,
    ``unpack_records`` below).  Returns records [N, 6] i64 (sized for the worst case N) and the device
    counts [2] = (participating points, local voxels).  ``tile_prefix``: optional int32 [n_tiles + 1] output,
    the index of each tile's first record."""
    lib = _lib.load()
    dev = _require_cuda(xyz, rgb, votes)
    N = xyz.shape[0]
    assert xyz.dtype == torch.float32 and rgb.dtype == torch.uint8 and tuple(rgb.shape) == (N, 3)
    nbytes = C.c_int64(0)
    _lib.check(lib.ddn_fuse_workspace_bytes(C.byref(grid), N, C.byref(nbytes)))
    ws = torch.empty(nbytes.value, dtype=torch.uint8, device=dev)
    if out is not None:  # caller-provided record buffer (e.g. NVLink-visible symmetric memory)
        assert out.dtype == torch.int64 and out.is_contiguous() and out.shape[0] >= N and out.shape[1] == _lib.RECORD_WORDS
        rec = out
    else:
        rec = torch.empty((max(N, 1), _lib.RECORD_WORDS), dtype=torch.int64, device=dev)
    counts = torch.zeros(2, dtype=torch.int64, device=dev)
    with torch.cuda.device(dev):
        _lib.check(
            lib.ddn_voxel_partials(
                C.byref(grid), N, int(row_len), _p(xyz), _p(rgb), _p(votes), int(vote_threshold), _p(rec), _p(tile_prefix),
                _p(counts), _p(ws), nbytes.value, _stream(),
            )
        )
    return rec, counts


def unpack_records(rec):
    """(keys, sums [n,3], colour sums [n,3], count) from records [n, 6] i64 (torch or numpy)."""
    lo = 0xFFFFFFFF
    return rec[:, 0], rec[:, 1:4], (rec[:, 4] >> 32) & lo, rec[:, 4] & lo, (rec[:, 5] >> 32) & lo, rec[:, 5] & lo


def voxel_merge_partials(records, grid: _lib.VoxelGrid, trim: bool = False, tile_range: tuple[int, int] = (0, 0)):
    """Owner-side merge of partial records [n, 6] i64 (any order) into final voxels.  ``tile_range``: the
    tiles this rank owns ((0, 0) = all); records of other tiles are ignored."""
    lib = _lib.load()
    dev = _require_cuda(records)
    n = records.shape[0]
    assert records.dtype == torch.int64 and (n == 0 or tuple(records.shape) == (n, _lib.RECORD_WORDS))
    nbytes = C.c_int64(0)
    _lib.check(lib.ddn_fuse_workspace_bytes(C.byref(grid), max(n, 1), C.byref(nbytes)))
    ws = torch.empty(nbytes.value, dtype=torch.uint8, device=dev)
    m = max(n, 1)
    out_keys = torch.empty(m, dtype=torch.int64, device=dev)
    out_xyz = torch.empty((m, 3), dtype=torch.float32, device=dev)
    out_rgb = torch.empty((m, 3), dtype=torch.uint8, device=dev)
    out_cnt = torch.empty(m, dtype=torch.int32, device=dev)
    counts = torch.zeros(2, dtype=torch.int64, device=dev)
    with torch.cuda.device(dev):
        _lib.check(
            lib.ddn_voxel_merge(
                C.byref(grid), n, _p(records), int(tile_range[0]), int(tile_range[1]), _p(out_keys), _p(out_xyz), _p(out_rgb),
                _p(out_cnt), _p(counts), _p(ws), nbytes.value, _stream(),
            )
        )
    if trim:
        mv = checked_voxel_count(counts)
        return out_keys[:mv], out_xyz[:mv], out_rgb[:mv], out_cnt[:mv], counts
    return out_keys, out_xyz, out_rgb, out_cnt, counts


def voxel_keys(xyz, voxel: float, origin) -> torch.Tensor:
    """Canonical 3x21-bit keys (int64 view of the uint64 key) for xyz [N,3] f32."""
    lib = _lib.load()
    dev = _require_cuda(xyz)
    g = _lib.VoxelGrid()
    g.voxel = float(np.float32(voxel))
    for i in range(3):
        g.origin[i] = float(np.float32(origin[i]))
        g.bits[i] = 21
        g.dims[i] = 0
    keys = torch.empty(xyz.shape[0], dtype=torch.int64, device=dev)
    with torch.cuda.device(dev):
        _lib.check(lib.ddn_voxel_keys(C.byref(g), xyz.shape[0], _p(xyz), _p(keys), _stream()))
    return keys


def project_points_device(points3d, cam_from_world34, kmat33):
    """float64 device form of the reference's project_points (scripts/test.py:58-76)."""
    lib = _lib.load()
    dev = _require_cuda(points3d, cam_from_world34, kmat33)
    n = points3d.shape[0]
    assert points3d.dtype == torch.float64 and cam_from_world34.dtype == torch.float64 and kmat33.dtype == torch.float64
    uv = torch.empty((n, 2), dtype=torch.float64, device=dev)
    z = torch.empty(n, dtype=torch.float64, device=dev)
    with torch.cuda.device(dev):
        _lib.check(lib.ddn_project_points(n, _p(points3d), _p(cam_from_world34), _p(kmat33), _p(uv), _p(z), _stream()))
    return uv, z


def unproject_points_device(points2d, depth, params4):
    """float64 device form of the reference's unproject_points (scripts/test.py:79-90); depth is float32."""
    lib = _lib.load()
    dev = _require_cuda(points2d, depth)
    n = points2d.shape[0]
    assert points2d.dtype == torch.float64 and depth.dtype == torch.float32 and len(params4) == 4
    out = torch.empty((n, 3), dtype=torch.float64, device=dev)
    par = (C.c_double * 4)(*[float(v) for v in params4])
    with torch.cuda.device(dev):
        _lib.check(lib.ddn_unproject_points(n, _p(points2d), _p(depth), par, _p(out), _stream()))
    return out
