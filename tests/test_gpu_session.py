"""GPU tests of the device-resident fusion session (include/ddn_b200.h: ddn_fuse_*): the sync-free form of
stage 4, K4's fused occupancy marking, the two K4 pixel layouts, and the owner-side merge over peer memory
(ddn_fuse_merge_peers) driven with VIRTUAL ranks on one GPU - R sessions in one process, whose buffers stand
in for the peers' NVLink-mapped ones, so the exchange logic is under bit-exact parity on a one-GPU box."""

import numpy as np
import pytest
import torch

from depthdensifier_b200.neighbours import default_vote_threshold, nearest_views_table
from depthdensifier_b200.synthetic import SceneConfig, make_scene

pytestmark = pytest.mark.gpu


def _cuda(x, dtype=None):
    t = torch.from_numpy(np.ascontiguousarray(x))
    if dtype is not None:
        t = t.to(dtype)
    return t.cuda().contiguous()


def _cloud(n, seed, spread=0.7):
    rng = np.random.default_rng(seed)
    xyz = rng.normal(0, spread, (n, 3)).astype(np.float32)
    rgb = rng.integers(0, 256, (n, 3)).astype(np.uint8)
    votes = rng.integers(0, 4, n).astype(np.uint8)
    votes[rng.random(n) < 0.05] = 255
    return xyz, rgb, votes


@pytest.mark.parametrize("dirty", [True, False])
def test_session_equals_host_grid_path(lib_built, dirty):
    """ddn_fuse_begin (grid from a box, on the device) + mark + finish == ddn_voxel_fuse on the same grid, and the
    session can be reused (its units are clean again after every step)."""
    from depthdensifier_b200 import ops

    sess = ops.FuseSession("cuda", max_cells=1 << 28, dirty=dirty)
    for step, (n, seed, voxel, thr) in enumerate([(120000, 1, 0.02, 2), (50000, 2, 0.05, 3), (120000, 1, 0.02, 2)]):
        xyz, rgb, votes = _cloud(n, seed)
        d_xyz, d_rgb, d_votes = _cuda(xyz), _cuda(rgb), _cuda(votes)
        box = ops.new_bbox("cuda")
        lo, hi = xyz.min(0) - 0.013, xyz.max(0) + 0.021  # any box enclosing the points
        enc = lambda f: np.where(f.view(np.int32) >= 0, f.view(np.int32), f.view(np.int32) ^ np.int32(0x7FFFFFFF))
        box.copy_(torch.from_numpy(np.concatenate([enc(lo.astype(np.float32)), enc(hi.astype(np.float32))])))
        sess.begin([box], voxel)
        sess.mark_points(d_xyz, d_votes, thr)
        k, x, c, m, counts = ops.fuse_finish(sess, d_xyz, d_rgb, d_votes, thr, row_len=400)
        grid = sess.host_grid()
        assert all(o <= l for o, l in zip(grid.origin, lo))
        k2, x2, c2, m2, counts2 = ops.voxel_fuse(d_xyz, d_rgb, d_votes, thr, grid, row_len=400)
        n_pts, mv = counts.cpu().tolist()
        assert [n_pts, mv] == counts2.cpu().tolist() and n_pts == int((votes < thr).sum()) and mv == len(k2)
        assert torch.equal(k[:mv], k2) and torch.equal(x[:mv], x2) and torch.equal(c[:mv], c2) and torch.equal(m[:mv], m2)


def test_session_capacity_and_empty(lib_built):
    from depthdensifier_b200 import _lib, ops

    sess = ops.FuseSession("cuda", max_cells=1 << 20)
    xyz, rgb, votes = _cloud(1000, 3, spread=5.0)
    d_xyz = _cuda(xyz)
    box = ops.new_bbox("cuda")
    sess.begin([box], 0.01)  # untouched box: +inf / -inf
    assert sess.grid_state().status == _lib.GRID_EMPTY
    k, x, c, m, counts = ops.fuse_finish(sess, d_xyz, _cuda(rgb), None, 1)
    assert counts.cpu().tolist() == [0, 0]
    enc = lambda f: np.where(f.view(np.int32) >= 0, f.view(np.int32), f.view(np.int32) ^ np.int32(0x7FFFFFFF))
    box.copy_(torch.from_numpy(np.concatenate([enc(xyz.min(0)), enc(xyz.max(0))])))
    sess.begin([box], 0.01)  # ~ (4000)^3 cells >> 2^20
    assert sess.grid_state().status == _lib.GRID_TOO_LARGE
    with pytest.raises(ops.DDNError):
        sess.host_grid()
    sess.mark_points(d_xyz, None, 1)
    assert ops.fuse_finish(sess, d_xyz, _cuda(rgb), None, 1)[4].cpu().tolist() == [0, 0]


def _scene_inputs(V=6, W=203, H=131, K=4, seed=9):
    sc = make_scene(SceneConfig(n_views=V, width=W, height=H, n_sparse=700, seed=seed))
    nbr = nearest_views_table(sc.cam_from_world.numpy(), K)
    return sc, nbr, default_vote_threshold(K)


@pytest.mark.parametrize("stride", [1, 2])
@pytest.mark.parametrize("sample_mode", ["nearest", "bilinear"])
def test_k4_layouts_and_fused_mark(lib_built, stride, sample_mode):
    """The two pixel layouts of K4 give identical bits (odd width: unaligned rows, row wraps inside a thread's
    four pixels), and the occupancy K4 marks equals the stand-alone mark pass."""
    from depthdensifier_b200 import ops

    sc, nbr, thr = _scene_inputs()
    V, H, W = sc.mono_depth.shape
    refined = _cuda((sc.mono_depth * sc.mask).numpy())
    d_nbr = _cuda(nbr.astype(np.int32))
    pair, src = ops.build_pair_tables(sc.cam_from_world.cuda(), sc.intrinsics.cuda(), d_nbr, 0, V, H, W)
    outs = []
    for layout in (0, 1):
        opts = ops.FilterOptions(stride=stride, sample_mode=sample_mode, pixel_layout=layout)
        bbox = ops.new_bbox("cuda")
        xyz, votes = ops.backproject_filter(refined, sc.normal.cuda(), d_nbr, pair, src, 0, thr, opts, bbox=bbox)
        outs.append((xyz, votes, bbox))
    assert torch.equal(outs[0][0], outs[1][0]) and torch.equal(outs[0][1], outs[1][1]) and torch.equal(outs[0][2], outs[1][2])
    xyz, votes, bbox = outs[0]
    # fused mark (both layouts) vs the stand-alone pass, through the fused cloud they produce
    rgb = sc.rgb.cuda()[:, ::stride, ::stride].contiguous()
    flat = (xyz.view(-1, 3), rgb.view(-1, 3), votes.view(-1))
    ref = None
    for layout in (None, 0, 1):
        sess = ops.FuseSession("cuda", max_cells=1 << 28)
        sess.begin([bbox], 0.05)
        assert sess.grid_state().status == 0
        if layout is None:
            sess.mark_points(flat[0], flat[2], thr)
        else:
            opts = ops.FilterOptions(stride=stride, sample_mode=sample_mode, pixel_layout=layout)
            xyz_l, votes_l = ops.backproject_filter(refined, sc.normal.cuda(), d_nbr, pair, src, 0, thr, opts, mark=sess)
            assert torch.equal(xyz_l, xyz) and torch.equal(votes_l, votes)
        k, x, c, m, counts = ops.fuse_finish(sess, *flat, thr, row_len=xyz.shape[2])
        n_pts, mv = counts.cpu().tolist()
        got = (n_pts, mv, k[:mv].clone(), x[:mv].clone(), c[:mv].clone(), m[:mv].clone())
        if ref is None:
            ref = got
            assert n_pts == int((votes < thr).sum()) and mv > 0
        else:
            assert got[:2] == ref[:2] and all(torch.equal(a, b) for a, b in zip(got[2:], ref[2:]))


def test_vote_threshold_clamp(lib_built):
    """A threshold of 255 or more keeps every point that exists and never a pixel without one (votes are u8,
    saturated at 254; 255 = no point)."""
    from depthdensifier_b200 import ops
    from depthdensifier_b200.engine import DensifyConfig, DensifyEngine

    sc, nbr, _ = _scene_inputs()
    args = (sc.mono_depth.cuda(), sc.normal.cuda(), sc.mask.cuda(), sc.rgb.cuda(), sc.cam_from_world.cuda(), sc.intrinsics.cuda(),
            sc.sparse_xyz.cuda(), sc.sparse_offsets.cuda(), _cuda(nbr.astype(np.int32)))
    res = DensifyEngine(DensifyConfig(voxel=0.02, vote_threshold=1000)).run(*args)
    n_valid = int((res.votes != 255).sum())
    assert int(res.counts[0]) == n_valid and int(res.keep_mask().sum()) == n_valid and 0 < n_valid < res.votes.numel()
    assert int(res.voxel_count.sum()) == n_valid
    res0 = DensifyEngine(DensifyConfig(voxel=0.02, vote_threshold=0)).run(*args)
    assert int(res0.counts[0]) == 0 and int(res0.keep_mask().sum()) == 0


@pytest.mark.parametrize("R", [2, 3, 8])
def test_merge_peers_virtual_ranks(lib_built, R):
    """ddn_fuse_merge_peers with R virtual ranks on one GPU: the points are dealt to R sessions (contiguous blocks,
    like views to ranks), each makes its partial records, then every "rank" runs the owner-side merge with the R
    buffer addresses.  The rank-ordered concatenation must equal the one-rank fusion bit for bit."""
    from depthdensifier_b200 import ops

    n, voxel, thr = 400000, 0.02, 2
    xyz, rgb, votes = _cloud(n, 11)
    xyz[: n // 4] *= 0.05  # a dense clump: some tiles are hit by every rank
    d_xyz, d_rgb, d_votes = _cuda(xyz), _cuda(rgb), _cuda(votes)
    grid = ops.make_grid(xyz.min(0), xyz.max(0), voxel)
    one = ops.voxel_fuse(d_xyz, d_rgb, d_votes, thr, grid, row_len=500)
    bounds = np.linspace(0, n, R + 1).astype(int)
    bounds[1] = max(bounds[1] // 7, 1)  # uneven shares; with R = 8 several ranks still overlap everywhere
    sessions, records = [], []
    for r in range(R):
        a, b = int(bounds[r]), int(bounds[r + 1])
        sess = ops.FuseSession("cuda", max_cells=1 << 28, tile_prefix=True)
        sess.begin_grid(grid)
        sess.mark_points(d_xyz[a:b], d_votes[a:b], thr)
        rec = torch.empty((max(b - a, 1), 6), dtype=torch.int64, device="cuda")
        ops.fuse_finish_partial(sess, d_xyz[a:b], d_rgb[a:b], d_votes[a:b], thr, rec, row_len=500)
        sessions.append(sess)
        records.append(rec)
    torch.cuda.synchronize()
    n_local = [int(s.counts[1]) for s in sessions]
    assert sum(int(s.counts[0]) for s in sessions) == int(one[4][0])
    pp = [s.tile_prefix.data_ptr() for s in sessions]
    pr = [t.data_ptr() for t in records]
    cap = n
    outs, plans = [], []
    for r in range(R):
        plan = torch.zeros(64, dtype=torch.int64, device="cuda")
        k, x, c, m, counts = ops.fuse_merge_peers(sessions[r], r, R, pr, pp, plan, cap)
        mv = int(counts[1])
        outs.append((k[:mv].clone(), x[:mv].clone(), c[:mv].clone(), m[:mv].clone()))
        plans.append(plan.cpu().tolist())
    # the ranks agree on the cuts, the ranges are contiguous and cover every tile that holds a record (the empty
    # head and tail of the grid belong to nobody), the shares add up to all records
    for r in range(R):
        assert plans[r][1] >= plans[r][0] and (r == 0 or plans[r][0] == plans[r - 1][1])
    assert 0 <= plans[0][0] and plans[-1][1] <= plans[-1][3]
    assert sum(p[2] for p in plans) == sum(n_local)
    tot = sum(n_local)
    for r in range(R):  # balanced up to one tile's worth of records per rank
        assert abs(plans[r][2] - tot / R) <= 24576 * R + tot * 0.02
    cat = [torch.cat([o[i] for o in outs]) for i in range(4)]
    assert torch.equal(cat[0], one[0]) and torch.equal(cat[1], one[1]) and torch.equal(cat[2], one[2]) and torch.equal(cat[3], one[3])
    assert bool((cat[0][1:] > cat[0][:-1]).all())


def test_sparse_dedup_n5(lib_built):
    """N5: with dedup_sparse the fused cloud is the oracle's voxel fusion minus the voxels whose key holds a sparse
    point (oracle/restatement.py:merge_sparse, dedup=True); everything that remains is bit-identical."""
    from depthdensifier_b200.engine import DensifyConfig, DensifyEngine
    from oracle import restatement as R

    sc, nbr, thr = _scene_inputs(V=6, W=160, H=120)
    args = (sc.mono_depth.cuda(), sc.normal.cuda(), sc.mask.cuda(), sc.rgb.cuda(), sc.cam_from_world.cuda(), sc.intrinsics.cuda(),
            sc.sparse_xyz.cuda(), sc.sparse_offsets.cuda(), _cuda(nbr.astype(np.int32)))
    voxel = 0.03
    plain = DensifyEngine(DensifyConfig(voxel=voxel)).run(*args)
    dedup = DensifyEngine(DensifyConfig(voxel=voxel, dedup_sparse=True)).run(*args)
    grid = plain.host_grid()
    assert list(grid.origin) == list(dedup.host_grid().origin)
    origin = np.array(list(grid.origin), np.float32)
    keys = plain.voxel_keys.cpu().numpy().view(np.uint64)
    dense_xyz = plain.voxel_xyz.cpu().numpy()
    merged_xyz, merged_rgb = R.merge_sparse(sc.sparse_xyz.numpy(), np.zeros((len(sc.sparse_xyz), 3), np.uint8), dense_xyz,
                                            plain.voxel_rgb.cpu().numpy(), dense_keys=keys, voxel=voxel, origin=origin, dedup=True)
    n_sparse = len(sc.sparse_xyz)
    s32 = sc.sparse_xyz.numpy().astype(np.float32)
    k3 = np.floor((s32 - origin) / np.float32(voxel)).astype(np.int64)
    sk = (k3[:, 0] | (k3[:, 1] << 21) | (k3[:, 2] << 42)).astype(np.uint64)[((k3 >= 0) & (k3 < (1 << 21))).all(1)]
    keep = ~np.isin(keys, sk)
    assert 0 < keep.sum() < len(keys)  # the sparse points sit on the surface: some dense voxels must go
    assert np.array_equal(dedup.voxel_keys.cpu().numpy().view(np.uint64), keys[keep])
    assert np.array_equal(dedup.voxel_xyz.cpu().numpy(), dense_xyz[keep])
    assert np.array_equal(dedup.voxel_rgb.cpu().numpy(), plain.voxel_rgb.cpu().numpy()[keep])
    assert np.array_equal(dedup.voxel_count.cpu().numpy(), plain.voxel_count.cpu().numpy()[keep])
    assert np.array_equal(merged_xyz[n_sparse:], dense_xyz[keep].astype(np.float64)) and len(merged_rgb) == len(merged_xyz)
