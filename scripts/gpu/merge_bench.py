#!/usr/bin/env python
"""Owner-side merge (ddn_fuse_merge_peers) at full size with VIRTUAL ranks on one GPU: the views of a workload are
dealt to R sessions in contiguous blocks exactly as ShardedDensifier deals them to ranks, every "rank" makes its
partial records, then the merge of some ranks is timed (CUDA events) and checked bit for bit against the one-rank
fusion.  All buffers are local, so this measures the kernels without the NVLink transfer; run it under
`ncu -k regex:merge_` for the per-kernel list.
  python scripts/gpu/merge_bench.py [workload] [R ...]"""
import json
import sys
from pathlib import Path

import numpy as np
import torch

ROOT = Path(__file__).resolve().parent.parent.parent
sys.path.insert(0, str(ROOT))
from bench import WORKLOADS, VOXEL  # noqa: E402
from depthdensifier_b200 import ops  # noqa: E402
from depthdensifier_b200.distributed import shard_bounds  # noqa: E402
from depthdensifier_b200.neighbours import default_vote_threshold, nearest_views_table  # noqa: E402
from depthdensifier_b200.synthetic import SceneConfig, make_scene  # noqa: E402


def main():
    wl = sys.argv[1] if len(sys.argv) > 1 else "cfg2"
    Rs = [int(a) for a in sys.argv[2:]] or [2, 8]
    V, W, H, K, C, _ = WORKLOADS[wl]
    dev = torch.device("cuda", 0)
    sc = make_scene(SceneConfig(n_views=V, width=W, height=H, n_sparse=C, seed=0), device=dev)
    nbr = torch.from_numpy(nearest_views_table(sc.cam_from_world.cpu().numpy(), K).astype(np.int32)).to(dev)
    thr = default_vote_threshold(K)
    kmat = torch.zeros((V, 3, 3), dtype=torch.float64, device=dev)
    kmat[:, 0, 0], kmat[:, 1, 1], kmat[:, 0, 2], kmat[:, 1, 2], kmat[:, 2, 2] = (sc.intrinsics[:, 0], sc.intrinsics[:, 1],
                                                                                 sc.intrinsics[:, 2], sc.intrinsics[:, 3], 1.0)
    pair, src = ops.build_pair_tables(sc.cam_from_world, sc.intrinsics, nbr, 0, V, H, W)
    box = ops.new_bbox(dev)
    refined, _ = ops.align_views(sc.mono_depth, sc.mask, sc.cam_from_world, kmat, sc.sparse_xyz, sc.sparse_offsets, C,
                                 ops.AlignOptions(zero_unmasked_passthrough=True), src_table=src, bbox=box)
    xyz, votes = ops.backproject_filter(refined, sc.normal, nbr, pair, src, 0, thr, ops.FilterOptions())
    del refined, pair
    flat = (xyz.view(-1, 3), sc.rgb.view(-1, 3), votes.view(-1))
    one_sess = ops.FuseSession(dev, 1 << 33)
    one_sess.begin([box], VOXEL)
    one_sess.mark_points(flat[0], flat[2], thr)
    one = ops.fuse_finish(one_sess, *flat, thr, row_len=W)
    mv_one = int(one[4][1])
    one = [t[:mv_one].clone() for t in one[:4]]
    del one_sess
    torch.cuda.empty_cache()
    report = {"workload": wl, "voxels": mv_one}
    P = H * W
    for R in Rs:
        bounds = shard_bounds(V, R)
        sessions, records = [], []
        for r in range(R):
            a, b = bounds[r][0] * P, bounds[r][1] * P
            s = ops.FuseSession(dev, 1 << 33, tile_prefix=True)
            s.begin([box], VOXEL)
            s.mark_points(flat[0][a:b], flat[2][a:b], thr)
            rec = torch.empty((max(b - a, 1), 6), dtype=torch.int64, device=dev)
            ops.fuse_finish_partial(s, flat[0][a:b], flat[1][a:b], flat[2][a:b], thr, rec, row_len=W)
            sessions.append(s)
            records.append(rec)
        torch.cuda.synchronize()
        n_local = [int(s.counts[1]) for s in sessions]
        pp = [s.tile_prefix.data_ptr() for s in sessions]
        pr = [t.data_ptr() for t in records]
        cap = max(b - a for a, b in bounds) * P + 24576 * R + 1024
        outs, times, recv = [], [], []
        for r in range(R):  # every rank once (the merge rewrites the rank's own units inside its range only)
            plan = torch.zeros(64, dtype=torch.int64, device=dev)
            out = ops.new_voxel_outputs(cap, dev)
            ev = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
            ev[0].record()
            k, x, c, m, counts = ops.fuse_merge_peers(sessions[r], r, R, pr, pp, plan, cap, out=out)
            ev[1].record()
            torch.cuda.synchronize()
            times.append(ev[0].elapsed_time(ev[1]))
            mv = int(counts[1])
            recv.append(int(plan[2]))
            outs.append([t[:mv].clone() for t in (k, x, c, m)])
            del out
        cat = [torch.cat([o[i] for o in outs]) for i in range(4)]
        same = all(torch.equal(cat[i], one[i]) for i in range(4))
        report[f"R{R}"] = {"merge_ms_per_rank": [round(t, 3) for t in times], "records_received": recv, "local_voxels": n_local,
                           "voxels_owned": [len(o[0]) for o in outs], "equals_one_rank": same}
        del sessions, records, outs, cat
        torch.cuda.empty_cache()
    print(json.dumps(report))


if __name__ == "__main__":
    main()
