"""world_size-2 (and 3) tests of the sharded driver on CPU with the gloo backend.

The exchange logic under test is the product's (depthdensifier_b200/distributed.py); the per-rank
compute is supplied by the oracle-backed stand-in in tests/cpu_backend.py.  Criterion:
the R-rank result equals the 1-rank result - keys, counts, colours AND positions bit for bit, because
voxel sums are integer fixed point."""

import os
import socket
import sys
from pathlib import Path

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = Path(__file__).resolve().parent.parent
for p in (str(ROOT), str(ROOT / "tests")):
    if p not in sys.path:
        sys.path.insert(0, p)

from depthdensifier_b200.distributed import ShardedDensifier, make_halo_plan, needed_views, shard_bounds  # noqa: E402
from depthdensifier_b200.engine import DensifyConfig  # noqa: E402
from depthdensifier_b200.neighbours import nearest_views_table  # noqa: E402
from depthdensifier_b200.synthetic import SceneConfig, make_scene, ring_poses  # noqa: E402

V, W, H, K, VOXEL = 8, 80, 60, 3, 0.05


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _scene():
    return make_scene(SceneConfig(n_views=V, width=W, height=H, n_sparse=300, seed=2))


def _run_rank(rank, world, port, out_dir):
    from cpu_backend import OracleBackend

    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    if world > 1:
        dist.init_process_group("gloo", rank=rank, world_size=world)
    sc = _scene()
    nbr = nearest_views_table(sc.cam_from_world.numpy(), K)
    lo, hi = shard_bounds(V, world)[rank] if world > 1 else (0, V)
    off = sc.sparse_offsets.numpy()
    sd = ShardedDensifier(DensifyConfig(voxel=VOXEL), "cpu", rank, world, V, lo, hi, sc.cam_from_world, sc.intrinsics, nbr, H, W,
                          backend=OracleBackend())
    res = sd.run(sc.mono_depth[lo:hi].contiguous(), sc.normal[lo:hi].contiguous(), sc.mask[lo:hi].contiguous(),
                 sc.rgb[lo:hi].contiguous(), sc.sparse_xyz[off[lo]:off[hi]].contiguous(),
                 (sc.sparse_offsets[lo:hi + 1] - off[lo]).contiguous())
    mv = int(res.counts[1])
    np.savez(os.path.join(out_dir, f"w{world}_r{rank}.npz"), keys=res.voxel_keys[:mv].numpy(), xyz=res.voxel_xyz[:mv].numpy(),
             rgb=res.voxel_rgb[:mv].numpy(), count=res.voxel_count[:mv].numpy(), votes=res.votes.numpy(),
             refined=res.refined.numpy(), n_points=int(res.counts[0]))
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


def _gather(out_dir, world):
    parts = [np.load(os.path.join(out_dir, f"w{world}_r{r}.npz")) for r in range(world)]
    cat = {k: np.concatenate([p[k] for p in parts]) for k in ("keys", "xyz", "rgb", "count", "votes", "refined")}
    cat["n_points"] = sum(int(p["n_points"]) for p in parts)
    cat["per_rank_voxels"] = [len(p["keys"]) for p in parts]
    cat["parts"] = parts
    return cat


@pytest.mark.parametrize("world", [2, 3])
def test_sharded_equals_single_rank(tmp_path, world):
    out = str(tmp_path)
    _run_rank(0, 1, _free_port(), out)
    mp.spawn(_run_rank, args=(world, _free_port(), out), nprocs=world, join=True)
    one, many = _gather(out, 1), _gather(out, world)
    assert np.array_equal(one["refined"], many["refined"]) and np.array_equal(one["votes"], many["votes"])
    assert one["n_points"] == many["n_points"]
    # every rank owns a disjoint, ascending key range; rank-ordered concatenation is globally sorted
    assert (np.diff(many["keys"]) > 0).all()
    assert np.array_equal(one["keys"], many["keys"])
    assert np.array_equal(one["count"], many["count"])
    assert np.array_equal(one["rgb"], many["rgb"])
    assert np.array_equal(one["xyz"], many["xyz"])  # integer fixed-point sums: bit-identical
    assert min(many["per_rank_voxels"]) > 0.5 * len(one["keys"]) / world  # sampled splitters balance the ranks


def test_halo_plan_is_consistent():
    poses = ring_poses(SceneConfig(n_views=23))
    nbr = nearest_views_table(poses, 6)
    for world in (1, 2, 4, 5):
        bounds = shard_bounds(23, world)
        assert bounds[0][0] == 0 and bounds[-1][1] == 23
        plans = [make_halo_plan(nbr, bounds, r) for r in range(world)]
        for r, p in enumerate(plans):
            lo, hi = bounds[r]
            assert list(p.slots[: hi - lo]) == list(range(lo, hi))
            # every neighbour of every own view is reachable through a slot
            for i in range(hi - lo):
                assert [int(p.slots[s]) for s in p.nbr_slots[i]] == [int(t) for t in nbr[lo + i]]
            # what r receives from q is exactly what q sends to r, in the same order
            pos = hi - lo
            for q in range(world):
                recv = [int(v) for v in p.slots[pos: pos + p.recv_counts[q]]]
                pos += p.recv_counts[q]
                sent = [int(v) + bounds[q][0] for v in plans[q].send_views[r]]
                assert recv == sent
            assert set(int(v) for v in p.slots) == set(range(lo, hi)) | set(int(v) for v in needed_views(nbr, lo, hi))
