"""Parity at BASELINE.json's sizes.

* configs[0] (20 views, 512x384, K=4, 1 cm voxels - the case the reference's CPU path can run): the whole
  device pipeline against the oracle chain, stage by stage, tie-aware votes, bit-exact voxel keys / counts.
* configs[1] (185 views, 1297x840, K=8): too large for the oracle, so size-independent properties of the
  domain: conservation of points, sortedness and uniqueness of keys, permutation invariance (integer sums),
  idempotence of fusion, sub-scene consistency of votes, bounding-box containment."""

import numpy as np
import pytest
import torch

from depthdensifier_b200.hashperm import hash_perm
from depthdensifier_b200.neighbours import default_vote_threshold, nearest_views_table
from depthdensifier_b200.synthetic import SceneConfig, make_scene
from oracle import restatement as R

import parity

pytestmark = pytest.mark.gpu
RTOL = 1e-5


def test_cfg1_pipeline_vs_oracle(lib_built):
    from depthdensifier_b200 import ops
    from depthdensifier_b200.engine import DensifyConfig, DensifyEngine

    V, W, H, K, voxel = 20, 512, 384, 4, 0.01
    sc = make_scene(SceneConfig(n_views=V, width=W, height=H, n_sparse=4096, seed=0))
    poses, intr = sc.cam_from_world.numpy(), sc.intrinsics.numpy()
    nbr = nearest_views_table(poses, K)
    thr = default_vote_threshold(K)
    eng = DensifyEngine(DensifyConfig(voxel=voxel))
    res = eng.run(sc.mono_depth.cuda(), sc.normal.cuda(), sc.mask.cuda(), sc.rgb.cuda(), sc.cam_from_world.cuda(),
                  sc.intrinsics.cuda(), sc.sparse_xyz.cuda(), sc.sparse_offsets.cuda(),
                  torch.from_numpy(nbr.astype(np.int32)).cuda())
    torch.cuda.synchronize()
    # stage 1 against the oracle (hash permutation shared)
    out = R.densify(sc.mono_depth.numpy(), sc.normal.numpy(), sc.mask.numpy(), sc.rgb.numpy(), sc.sparse_xyz.numpy(),
                    sc.sparse_offsets.numpy(), poses, intr, nbr, thr, randperm=lambda n: hash_perm(n, 0), voxel=None)
    refined = res.refined.cpu().numpy()
    np.testing.assert_allclose(refined, out["refined"], rtol=RTOL, atol=0)
    valid = out["refined"] > 0
    votes = res.votes.cpu().numpy()
    assert np.array_equal(votes != 255, valid) and valid.sum() > 3_000_000
    xyz = res.xyz.cpu().numpy()[valid]
    assert np.abs(xyz - out["points"]).max() <= RTOL * max(1.0, np.abs(out["points"]).max())
    # stage 3, exact: oracle votes computed from the GPU's OWN refined maps and points (identical inputs)
    vv, yy, xx = np.nonzero(valid)
    pts64 = np.concatenate([R.backproject_view(refined[v], intr[v], poses[v])[0] for v in range(V)])
    nrm = sc.normal.numpy()[valid]
    ref_votes, nties = parity.votes_with_ties(pts64, nrm, vv, refined, poses, intr, nbr)
    frac = parity.assert_votes_match(votes[valid], ref_votes, nties)
    assert frac < 0.01
    # stage 4 on the GPU's kept points: keys / counts / colours bit-exact, positions within tolerance
    keep = votes[valid] < thr
    kept_xyz, kept_rgb = xyz[keep], sc.rgb.numpy()[valid][keep]
    origin = np.array(list(res.host_grid().origin), np.float32)  # derived on the device from the alignment kernel's box
    assert (origin <= kept_xyz.min(0)).all()
    k_ref, m_ref, c_ref, n_ref = R.voxel_fuse(kept_xyz, kept_rgb, voxel, origin)
    mv = int(res.counts[1])
    assert res.counts.cpu().tolist() == [int(keep.sum()), len(k_ref)]
    assert np.array_equal(res.voxel_keys[:mv].cpu().numpy().view(np.uint64), k_ref)
    assert np.array_equal(res.voxel_count[:mv].cpu().numpy(), n_ref)
    assert np.array_equal(res.voxel_rgb[:mv].cpu().numpy(), c_ref)
    assert np.abs(res.voxel_xyz[:mv].cpu().numpy() - m_ref).max() <= RTOL * max(1.0, np.abs(m_ref).max())


@pytest.mark.parametrize("V,W,H,K", [(185, 1297, 840, 8), (6, 3840, 2160, 4), (48, 1920, 1080, 8), (24, 1600, 1200, 10)],
                         ids=["cfg2", "4k_slice_of_cfg5", "cfg3_shape_48_views", "cfg4_shape_24_views_k10"])
def test_fullsize_properties(lib_built, V, W, H, K):
    from depthdensifier_b200 import ops
    from depthdensifier_b200.engine import DensifyConfig, DensifyEngine

    voxel = 0.01
    dev = torch.device("cuda", 0)
    sc = make_scene(SceneConfig(n_views=V, width=W, height=H, n_sparse=4096, seed=0), device=dev)
    poses = sc.cam_from_world.cpu().numpy()
    nbr_np = nearest_views_table(poses, K)
    nbr = torch.from_numpy(nbr_np.astype(np.int32)).to(dev)
    thr = default_vote_threshold(K)
    eng = DensifyEngine(DensifyConfig(voxel=voxel))
    res = eng.run(sc.mono_depth, sc.normal, sc.mask, sc.rgb, sc.cam_from_world, sc.intrinsics, sc.sparse_xyz, sc.sparse_offsets, nbr)
    n_pts, mv = (int(v) for v in res.counts.cpu().tolist())
    keep = res.keep_mask()
    # conservation: every kept point is in exactly one voxel
    assert n_pts == int(keep.sum()) and int(res.voxel_count[:mv].sum()) == n_pts
    # keys strictly ascending (sorted + unique), every axis inside the grid
    keys = res.voxel_keys[:mv]
    assert bool((keys[1:] > keys[:-1]).all())
    grid = res.host_grid()
    for ax in range(3):
        assert int(((keys >> (21 * ax)) & 0x1FFFFF).max()) < grid.dims[ax]
    # fused positions lie inside their voxel (up to float32 rounding of the centre)
    org = torch.tensor(list(grid.origin), device=dev)
    cell = torch.stack([(keys >> (21 * ax)) & 0x1FFFFF for ax in range(3)], 1).float()
    rel = (res.voxel_xyz[:mv] - org) / voxel - cell
    assert float(rel.min()) > -1e-3 and float(rel.max()) < 1 + 1e-3
    # bounding box of the kept points == what K4 reduced
    bb = ops.decode_bbox(res.bbox)
    kept = res.xyz[keep]
    assert np.array_equal(bb[:3], kept.min(0).values.cpu().numpy()) and np.array_equal(bb[3:], kept.max(0).values.cpu().numpy())
    # the alignment kernel's tile-wise box encloses EVERY back-projected pixel, and not by much
    bv = ops.decode_bbox(res.bbox_valid)
    allp = res.xyz[res.votes != 255]
    lo, hi = allp.min(0).values.cpu().numpy(), allp.max(0).values.cpu().numpy()
    assert (bv[:3] <= lo).all() and (bv[3:] >= hi).all()
    assert np.prod(bv[3:] - bv[:3]) < 2.0 * np.prod(hi - lo)
    # permutation invariance (integer sums): fusing the same kept points in a random order gives the same bits
    perm = torch.randperm(kept.shape[0], device=dev, generator=torch.Generator(device=dev).manual_seed(1))
    rgb_kept = sc.rgb[keep]
    k2, x2, c2, n2, cnt2 = ops.voxel_fuse(kept[perm].contiguous(), rgb_kept[perm].contiguous(), None, 1, grid)
    assert torch.equal(k2, keys) and torch.equal(n2, res.voxel_count[:mv]) and torch.equal(c2, res.voxel_rgb[:mv])
    assert torch.equal(x2, res.voxel_xyz[:mv])
    # idempotence: fusing the fused cloud reproduces it - except for the few means that float32 rounding puts
    # exactly on a voxel face (they may fall into the neighbour cell)
    k3, x3, c3, n3, cnt3 = ops.voxel_fuse(res.voxel_xyz[:mv].contiguous(), res.voxel_rgb[:mv].contiguous(), None, 1, grid)
    assert int(n3.sum()) == mv and mv - len(k3) <= 3e-4 * mv
    same = torch.isin(k3, keys)
    assert float(same.float().mean()) > 1 - 3e-4
    single = n3 == 1
    pos = torch.searchsorted(keys, k3[single & same])
    assert float((x3[single & same] - res.voxel_xyz[:mv][pos]).abs().max()) <= 1e-6
    # sub-scene consistency: the votes of 3 source views recomputed alone (same refined maps) are identical
    sub = [0, V // 2, V - 1]
    pair, src = ops.build_pair_tables(sc.cam_from_world, sc.intrinsics, nbr, 0, V, H, W)
    for s in sub:
        xyz_s, votes_s = ops.backproject_filter(res.refined, sc.normal[s:s + 1].contiguous(), nbr, pair[s:s + 1].contiguous(),
                                                src[s:s + 1].contiguous(), s, thr, eng.cfg.filter)
        assert torch.equal(votes_s[0], res.votes[s]) and torch.equal(xyz_s[0], res.xyz[s])
    # and against the float64 oracle on a 16-row strip of THREE views - first, middle, last (tie-aware)
    intr_np = sc.intrinsics.cpu().numpy()
    refined_host = {}
    for s in sub:
        refined_s = res.refined[s].cpu().numpy()
        rows = slice(H // 2, H // 2 + 16) if s != sub[-1] else slice(H - 16, H)  # the last view: the bottom rows (tile tails)
        strip = np.zeros_like(refined_s)
        strip[rows] = refined_s[rows]
        pts64, pyv, pxv = R.backproject_view(strip, intr_np[s], poses[s])
        nrm = sc.normal[s].cpu().numpy()[pyv, pxv]
        ref_votes = np.zeros(len(pts64), np.int64)
        nties = np.zeros(len(pts64), np.int64)
        for t in sorted(set(int(t) for t in nbr_np[s] if t >= 0)):
            if t not in refined_host:
                refined_host[t] = res.refined[t].cpu().numpy()
            v_t, tie_t = parity.pair_ties(pts64, nrm, np.full(len(pts64), s), t, refined_host[t], poses[t], intr_np[t])
            ref_votes += v_t
            nties += tie_t
        parity.assert_votes_match(res.votes[s].cpu().numpy()[pyv, pxv], ref_votes, nties, max_tie_fraction=0.05)
