#!/usr/bin/env python
"""K3 with its depth tile loaded by one TMA tensor copy (ddn_align_config.use_tma) against the default LDG path:
bit-identity on small scenes with partial tiles, then timing at 1920x1080 (the TMA form needs W % 4 == 0, so not cfg 2).
The measurement of record was taken with the no-Python harness k3_tma_check.cu (profiles/r02_k3_tma_check.log)."""
import json
import sys
from pathlib import Path

import numpy as np
import torch

ROOT = Path(__file__).resolve().parent.parent.parent
sys.path.insert(0, str(ROOT))
from depthdensifier_b200 import ops  # noqa: E402
from depthdensifier_b200.synthetic import SceneConfig, make_scene  # noqa: E402


def align(sc, C, use_tma, reps=1):
    V = sc.mono_depth.shape[0]
    kmat = torch.zeros((V, 3, 3), dtype=torch.float64, device=sc.mono_depth.device)
    kmat[:, 0, 0], kmat[:, 1, 1], kmat[:, 0, 2], kmat[:, 1, 2], kmat[:, 2, 2] = (sc.intrinsics[:, 0], sc.intrinsics[:, 1],
                                                                                 sc.intrinsics[:, 2], sc.intrinsics[:, 3], 1.0)
    opts = ops.AlignOptions(zero_unmasked_passthrough=True, use_tma=use_tma)
    out = torch.empty_like(sc.mono_depth)
    ts = []
    for _ in range(reps):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        refined, stats = ops.align_views(sc.mono_depth, sc.mask, sc.cam_from_world, kmat, sc.sparse_xyz, sc.sparse_offsets, C, opts, out=out)
        b.record()
        torch.cuda.synchronize()
        ts.append(a.elapsed_time(b))
    return refined.clone(), stats.clone(), float(np.median(ts[1:])) if reps > 2 else ts[-1]


def main():
    dev = torch.device("cuda", 0)
    res = {"identical": {}}
    for (V, W, H) in ((5, 160, 120), (4, 200, 152), (3, 512, 384)):
        sc = make_scene(SceneConfig(n_views=V, width=W, height=H, n_sparse=600, seed=3), device=dev)
        r0, s0, _ = align(sc, 600, False)
        r1, s1, _ = align(sc, 600, True)
        res["identical"][f"{V}x{W}x{H}"] = bool(torch.equal(r0, r1) and torch.equal(s0, s1))
    sc = make_scene(SceneConfig(n_views=48, width=1920, height=1080, n_sparse=4096, seed=0), device=dev)
    r0, _, t0 = align(sc, 4096, False, reps=7)
    r1, _, t1 = align(sc, 4096, True, reps=7)
    res["identical"]["48x1920x1080"] = bool(torch.equal(r0, r1))
    res["align_ms_48_views_1920x1080"] = {"ldg (default)": round(t0, 4), "tma": round(t1, 4)}
    print(json.dumps(res))


if __name__ == "__main__":
    main()
