#!/usr/bin/env python
"""Stage-level micro-benchmark on one GPU (CUDA events, median of N): alignment with / without the bounding-box
epilogue, K4 with / without bounding box and occupancy marking in both pixel layouts, the fusion passes.
  python scripts/gpu/kbench.py [workload] [reps]      (DDN_LIB_PATH selects a variant build)"""
import json
import sys
from pathlib import Path

import numpy as np
import torch

ROOT = Path(__file__).resolve().parent.parent.parent
sys.path.insert(0, str(ROOT))
from bench import WORKLOADS, VOXEL  # noqa: E402
from depthdensifier_b200 import ops  # noqa: E402
from depthdensifier_b200.neighbours import default_vote_threshold, nearest_views_table  # noqa: E402
from depthdensifier_b200.synthetic import SceneConfig, make_scene  # noqa: E402


def timeit(fn, reps):
    for _ in range(2):
        fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(reps):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        fn()
        b.record()
        torch.cuda.synchronize()
        ts.append(a.elapsed_time(b))
    return float(np.median(ts))


def main():
    wl = sys.argv[1] if len(sys.argv) > 1 else "cfg2"
    reps = int(sys.argv[2]) if len(sys.argv) > 2 else 7
    V, W, H, K, C, _ = WORKLOADS[wl]
    dev = torch.device("cuda", 0)
    sc = make_scene(SceneConfig(n_views=V, width=W, height=H, n_sparse=C, seed=0), device=dev)
    nbr = torch.from_numpy(nearest_views_table(sc.cam_from_world.cpu().numpy(), K).astype(np.int32)).to(dev)
    thr = default_vote_threshold(K)
    kmat = torch.zeros((V, 3, 3), dtype=torch.float64, device=dev)
    kmat[:, 0, 0], kmat[:, 1, 1], kmat[:, 0, 2], kmat[:, 1, 2], kmat[:, 2, 2] = (sc.intrinsics[:, 0], sc.intrinsics[:, 1],
                                                                                 sc.intrinsics[:, 2], sc.intrinsics[:, 3], 1.0)
    pair, src = ops.build_pair_tables(sc.cam_from_world, sc.intrinsics, nbr, 0, V, H, W)
    refined = torch.empty_like(sc.mono_depth)
    aopts = ops.AlignOptions(zero_unmasked_passthrough=True)
    box = ops.new_bbox(dev)
    out = {"workload": wl}
    align = lambda **kw: ops.align_views(sc.mono_depth, sc.mask, sc.cam_from_world, kmat, sc.sparse_xyz, sc.sparse_offsets, C, aopts,
                                         out=refined, **kw)
    out["align"] = timeit(lambda: align(), reps)
    out["align_bbox"] = timeit(lambda: align(src_table=src, bbox=ops.init_bbox(box)), reps)
    xyz = torch.empty((V, H, W, 3), dtype=torch.float32, device=dev)
    votes = torch.empty((V, H, W), dtype=torch.uint8, device=dev)
    sess = ops.FuseSession(dev, 1 << 33)
    bb = ops.new_bbox(dev)

    def k4(layout=0, bbox=None, mark=None):
        ops.backproject_filter(refined, sc.normal, nbr, pair, src, 0, thr, ops.FilterOptions(pixel_layout=layout), bbox=bbox, xyz_out=xyz,
                               votes_out=votes, mark=mark)

    for layout in (0, 1):
        out[f"k4_l{layout}_plain"] = timeit(lambda: k4(layout), reps)
        out[f"k4_l{layout}_bbox"] = timeit(lambda: k4(layout, bbox=ops.init_bbox(bb)), reps)

        def marked():
            sess.begin([box], VOXEL)
            k4(layout, mark=sess)

        out[f"begin+k4_l{layout}_mark"] = timeit(marked, reps)
    sess_nd = ops.FuseSession(dev, 1 << 33, dirty=False)

    def marked_nd():
        sess_nd.begin([box], VOXEL)
        k4(0, mark=sess_nd)

    out["begin+k4_l0_mark_nodirty"] = timeit(marked_nd, reps)
    out["begin_nodirty"] = timeit(lambda: sess_nd.begin([box], VOXEL), reps)
    out["begin"] = timeit(lambda: sess.begin([box], VOXEL), reps)
    flat = (xyz.view(-1, 3), sc.rgb.view(-1, 3), votes.view(-1))

    def mark_alone():
        sess.begin([box], VOXEL)
        sess.mark_points(flat[0], flat[2], thr)

    out["begin+mark_points"] = timeit(mark_alone, reps)
    outs = ops.new_voxel_outputs(flat[0].shape[0], dev)

    def full():
        sess.begin([box], VOXEL)
        k4(0, mark=sess)
        ops.fuse_finish(sess, *flat, thr, row_len=W, out=outs)

    out["begin+k4_mark+finish"] = timeit(full, reps)
    sess.begin([box], VOXEL)
    k4(0, mark=sess)
    out["finish_only(rank+accumulate+finalize)"] = timeit(lambda: ops.fuse_finish(sess, *flat, thr, row_len=W, out=outs), reps)
    sess_nd.begin([box], VOXEL)
    k4(0, mark=sess_nd)
    out["finish_only_nodirty"] = timeit(lambda: ops.fuse_finish(sess_nd, *flat, thr, row_len=W, out=outs), reps)
    out["counts"] = sess.counts.cpu().tolist()
    print(json.dumps({k: (round(v, 4) if isinstance(v, float) else v) for k, v in out.items()}))


if __name__ == "__main__":
    main()
