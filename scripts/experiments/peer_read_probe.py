#!/usr/bin/env python
"""What do SM-issued loads get out of NVLink peer memory?  Every rank reads a 256 MB symmetric buffer of rank+1 with
ld.global.cv / plain ld.global / ld.global.nc, 1..8 16-byte loads in flight per thread, and for comparison pulls the
same bytes with a copy-engine transfer.  torchrun --nproc-per-node N scripts/experiments/peer_read_probe.py"""
import ctypes as C
import json
import os
import sys
from pathlib import Path

import torch
import torch.distributed as dist
import torch.distributed._symmetric_memory as symm

sys.path.insert(0, str(Path(__file__).resolve().parent.parent.parent))
from depthdensifier_b200 import _lib  # noqa: E402


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")
    dist.init_process_group("nccl", device_id=dev)
    lib = _lib.load()
    nbytes = 256 << 20
    buf = symm.empty(nbytes, dtype=torch.uint8, device=dev)
    hdl = symm.rendezvous(buf, dist.group.WORLD)
    buf.random_(0, 255)
    peer = hdl.get_buffer((rank + 1) % world, (nbytes,), torch.uint8)
    out = torch.zeros(4, dtype=torch.int32, device=dev)
    dst = torch.empty(nbytes, dtype=torch.uint8, device=dev)
    res = {}

    def timed(fn):
        hdl.barrier()
        torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        best = 1e9
        for _ in range(4):
            hdl.barrier()
            a.record()
            fn()
            b.record()
            torch.cuda.synchronize()
            best = min(best, a.elapsed_time(b))
        t = torch.tensor([best], device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return nbytes / (float(t.item()) * 1e-3) / 1e9

    st = C.c_void_p(torch.cuda.current_stream().cuda_stream)
    for src_name, src in (("peer", peer), ("local", buf)):
        for mode, mname in ((0, "cv"), (1, "plain"), (2, "nc")):
            for per_thread in (1, 3, 8):
                for ctas in (148 * 8,):
                    fn = lambda: _lib.check(lib.ddn_debug_peer_read(C.c_void_p(src.data_ptr()), nbytes, mode, per_thread, ctas,
                                                                    C.c_void_p(out.data_ptr()), st))
                    res[f"{src_name}_{mname}_x{per_thread}"] = round(timed(fn), 1)
    res["peer_copy_engine"] = round(timed(lambda: dst.copy_(peer, non_blocking=True)), 1)
    if rank == 0:
        print(json.dumps({"world": world, "GBps_per_rank_min_over_ranks": res}))
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
