"""Probe: NCCL all_to_all_single vs symmetric-memory peer copies for the voxel record exchange."""
import os, time, torch, torch.distributed as dist
rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local); dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
n = 7_400_000 * 6  # int64 words per rank (355 MB)
src = torch.arange(n, dtype=torch.int64, device=dev) + rank
dst = torch.empty_like(src)
per = n // world
def timeit(fn, reps=10):
    for _ in range(3): fn()
    torch.cuda.synchronize(); dist.barrier(); torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(reps): fn()
    b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b) / reps
t = timeit(lambda: dist.all_to_all_single(dst[: per * world], src[: per * world]))
if rank == 0: print(f"nccl all_to_all_single {per*8*(world-1)/1e6:.0f} MB out per rank: {t:.3f} ms -> {per*8*(world-1)/t/1e6:.0f} GB/s")
try:
    import torch.distributed._symmetric_memory as symm
    buf = symm.empty(n, dtype=torch.int64, device=dev)
    hdl = symm.rendezvous(buf, dist.group.WORLD)
    buf.copy_(src)
    peers = [hdl.get_buffer(q, (n,), torch.int64) for q in range(world)]
    def pull():
        hdl.barrier()
        for k in range(1, world):
            q = (rank + k) % world
            dst[q * per:(q + 1) * per].copy_(peers[q][rank * per:(rank + 1) * per], non_blocking=True)
        hdl.barrier()
    t2 = timeit(pull)
    if rank == 0: print(f"symmetric-memory pull: {t2:.3f} ms -> {per*8*(world-1)/t2/1e6:.0f} GB/s")
    ok = all(bool((dst[q * per:(q + 1) * per] == torch.arange(rank * per, (rank + 1) * per, device=dev) + q).all()) for q in range(world) if q != rank)
    if rank == 0: print("pull data ok:", ok)
except Exception as e:
    if rank == 0: print("symmetric memory unavailable:", repr(e)[:300])
dist.barrier(); dist.destroy_process_group()
