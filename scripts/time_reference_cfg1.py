#!/usr/bin/env python
"""BASELINE configs[0] on the host CPU, in the build container (needs /root/reference): the reference's UNMODIFIED
``scripts/test.py:main`` under the stand-in pycolmap / moge modules (oracle/run_reference.py; every view tests every
view, K = V = 20, as the reference does) and the CPU port with the same all-views table and with the K = 4 table of the
config.  Prints one JSON line for BASELINE.md; test infrastructure, not part of the product."""
import json
import os
import sys
import time
from pathlib import Path

import numpy as np
import torch

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
from depthdensifier_b200.hashperm import hash_perm  # noqa: E402
from depthdensifier_b200.neighbours import all_views_table, default_vote_threshold, nearest_views_table  # noqa: E402
from depthdensifier_b200.synthetic import SceneConfig, make_scene  # noqa: E402
from oracle import restatement as R  # noqa: E402
from oracle.run_reference import reference_available, run_reference_main  # noqa: E402


def main():
    V, W, H, K = 20, 512, 384, 4
    sc = make_scene(SceneConfig(n_views=V, width=W, height=H, n_sparse=4096, seed=0))
    out = {"config": "cfg1: 20 views 512x384", "cores": os.cpu_count(), "torch_threads": torch.get_num_threads()}
    if reference_available():
        t0 = time.perf_counter()
        ref = run_reference_main(sc, downsample_density=1, vote_threshold=5, randperm=lambda n: torch.from_numpy(hash_perm(n, 0)))
        dt = time.perf_counter() - t0
        n = int(len(ref["points"]))
        out["reference_main"] = {"seconds": round(dt, 2), "valid_pixels": n, "pixels_per_s": round(n / dt), "k": "all 20 views (K = V)",
                                 "note": "unmodified scripts/test.py:main incl. its PNG reads and per-point add_point3D loop"}
    poses, intr = sc.cam_from_world.numpy(), sc.intrinsics.numpy()
    for name, nbr, thr in (("port_all_views", all_views_table(V), 5), ("port_k4", nearest_views_table(poses, K), default_vote_threshold(K))):
        t0 = time.perf_counter()
        res = R.densify(sc.mono_depth.numpy(), sc.normal.numpy(), sc.mask.numpy(), sc.rgb.numpy(), sc.sparse_xyz.numpy(),
                        sc.sparse_offsets.numpy(), poses, intr, nbr, thr, randperm=lambda m: hash_perm(m, 0), voxel=0.01)
        dt = time.perf_counter() - t0
        n = int(len(res["points"]))
        out[name] = {"seconds": round(dt, 2), "valid_pixels": n, "pixels_per_s": round(n / dt)}
    print(json.dumps(out))


if __name__ == "__main__":
    main()
