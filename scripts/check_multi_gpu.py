#!/usr/bin/env python
"""Multi-GPU equivalence check (run under torchrun on N GPUs of one node):
the N-rank sharded pipeline must reproduce the 1-rank pipeline - refined depth, votes, voxel keys,
counts, colours and positions bit for bit.

  python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 \
      scripts/check_multi_gpu.py
"""

import os
import sys
from pathlib import Path

import numpy as np
import torch
import torch.distributed as dist

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))

from depthdensifier_b200.distributed import ShardedDensifier, shard_bounds  # noqa: E402
from depthdensifier_b200.engine import DensifyConfig  # noqa: E402
from depthdensifier_b200.neighbours import nearest_views_table  # noqa: E402
from depthdensifier_b200.synthetic import SceneConfig, make_scene  # noqa: E402


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    V, W, H, K, voxel = 24, 320, 240, 4, 0.02
    sc = make_scene(SceneConfig(n_views=V, width=W, height=H, n_sparse=1500, seed=4))  # host, identical on all ranks
    nbr = nearest_views_table(sc.cam_from_world.numpy(), K)
    off = sc.sparse_offsets.numpy()

    sharded = {}

    def run(r, w, lo, hi, group):
        sd = ShardedDensifier(DensifyConfig(voxel=voxel), dev, r, w, V, lo, hi, sc.cam_from_world, sc.intrinsics, nbr, H, W)
        sharded[w] = sd
        res = sd.run(sc.mono_depth[lo:hi].to(dev), sc.normal[lo:hi].to(dev), sc.mask[lo:hi].to(dev), sc.rgb[lo:hi].to(dev),
                     sc.sparse_xyz[off[lo]:off[hi]].to(dev), (sc.sparse_offsets[lo:hi + 1] - off[lo]).to(dev))
        mv = int(res.counts[1])
        return {"keys": res.voxel_keys[:mv].cpu().numpy(), "xyz": res.voxel_xyz[:mv].cpu().numpy(),
                "rgb": res.voxel_rgb[:mv].cpu().numpy(), "count": res.voxel_count[:mv].cpu().numpy(),
                "votes": res.votes.cpu().numpy(), "refined": res.refined.cpu().numpy(), "n": int(res.counts[0])}

    lo, hi = shard_bounds(V, world)[rank]
    mine = run(rank, world, lo, hi, None)
    gathered = [None] * world
    dist.gather_object(mine, gathered if rank == 0 else None, dst=0)
    ok = True
    if rank == 0:
        one = run(0, 1, 0, V, None)
        cat = {k: np.concatenate([g[k] for g in gathered]) for k in ("keys", "xyz", "rgb", "count", "votes", "refined")}
        for k in ("refined", "votes", "keys", "count", "rgb", "xyz"):
            same = np.array_equal(cat[k], one[k])
            ok &= same
            print(f"{k:8s} {'identical' if same else 'DIFFERENT'}  {cat[k].shape}")
        ok &= sum(g["n"] for g in gathered) == one["n"]
        print("voxels per rank:", [len(g["keys"]) for g in gathered], "points:", one["n"])
    # the end-to-end host entry point (pinned host arrays, normals read in place, peer-memory exchanges) must give the same cloud
    sd = sharded[world]
    host = sd.pin_host_inputs(sc.mono_depth[lo:hi], sc.normal[lo:hi], sc.mask[lo:hi], sc.rgb[lo:hi], sc.sparse_xyz[off[lo]:off[hi]],
                              sc.sparse_offsets[lo:hi + 1] - off[lo])
    for _ in range(2):
        out = sd.run_host(*host, chunk_views=5)
    mine_h = {k: out[k].numpy().copy() for k in ("keys", "xyz", "rgb", "count")} if "keys" in out else None
    gathered_h = [None] * world
    dist.gather_object(mine_h, gathered_h if rank == 0 else None, dst=0)
    if rank == 0:
        for k in ("keys", "count", "rgb", "xyz"):
            same = np.array_equal(np.concatenate([g[k] for g in gathered_h]), one[k])
            ok &= same
            print(f"run_host {k:6s} {'identical' if same else 'DIFFERENT'}")
        print("MULTI-GPU CHECK", "PASSED" if ok else "FAILED")
    dist.barrier()
    dist.destroy_process_group()
    sys.exit(0 if ok else 1)


if __name__ == "__main__":
    main()
