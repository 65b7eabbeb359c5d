#!/usr/bin/env python
"""End-to-end host path on one GPU, cfg 2: run_host back to back vs the pipelined submit/collect form, normals read in
place vs copied, mask as bytes vs bits - each variant twice, to separate real effects from run-to-run noise."""
import json
import sys
import time
from pathlib import Path

import numpy as np
import torch

ROOT = Path(__file__).resolve().parent.parent.parent
sys.path.insert(0, str(ROOT))
from bench import WORKLOADS, VOXEL  # noqa: E402
from depthdensifier_b200.distributed import ShardedDensifier  # noqa: E402
from depthdensifier_b200.engine import DensifyConfig  # noqa: E402
from depthdensifier_b200.neighbours import default_vote_threshold, nearest_views_table  # noqa: E402
from depthdensifier_b200.synthetic import SceneConfig, make_scene  # noqa: E402


def main():
    wl = sys.argv[1] if len(sys.argv) > 1 else "cfg2"
    V, W, H, K, C, _ = WORKLOADS[wl]
    dev = torch.device("cuda", 0)
    sc = make_scene(SceneConfig(n_views=V, width=W, height=H, n_sparse=C, seed=0), device=dev)
    nbr = nearest_views_table(sc.cam_from_world.cpu().numpy(), K)
    sd = ShardedDensifier(DensifyConfig(voxel=VOXEL, vote_threshold=default_vote_threshold(K)), dev, 0, 1, V, 0, V, sc.cam_from_world,
                          sc.intrinsics, nbr, H, W)
    host_in = [t.cpu() for t in (sc.mono_depth, sc.normal, sc.mask, sc.rgb, sc.sparse_xyz, sc.sparse_offsets)]
    del sc
    torch.cuda.empty_cache()
    res = {}
    for packed in (False, True):
        host = sd.pin_host_inputs(*host_in, pack_mask=packed)
        for in_place in (True, False):
            for pipelined in (False, True):
                for rep in range(2):
                    for _ in range(2):
                        sd.run_host(*host, normals_in_place=in_place)
                    torch.cuda.synchronize()
                    n = 6
                    t0 = time.perf_counter()
                    if pipelined:
                        prev = None
                        for _ in range(n):
                            tk = sd.submit_host(*host, normals_in_place=in_place)
                            if prev is not None:
                                sd.collect_host(prev)
                            prev = tk
                        sd.collect_host(prev)
                    else:
                        for _ in range(n):
                            sd.run_host(*host, normals_in_place=in_place)
                    torch.cuda.synchronize()
                    res.setdefault(f"mask_{'bits' if packed else 'bytes'}|normals_{'in_place' if in_place else 'copied'}|"
                                   f"{'pipelined' if pipelined else 'one_at_a_time'}", []).append(round((time.perf_counter() - t0) * 1e3 / n, 2))
    print(json.dumps(res, indent=1))


if __name__ == "__main__":
    main()
