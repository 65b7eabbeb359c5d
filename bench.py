#!/usr/bin/env python
"""Benchmark of the densification hot path (BASELINE.json metric: depth pixels fused per second and
HBM GB/s as a fraction of the roofline).

  python bench.py [--gpus N] [--steps K] [--warmup W] [--workload cfg2] [--impl reference]

A step is one pass of the whole device pipeline (align -> [neighbour-depth exchange] -> back-project +
consistency vote -> voxel fusion [-> voxel all-to-all]) over one synthetic scene.  Rank 0 prints ONE
JSON line.  `value` has inputs resident in HBM; `e2e` goes through the public engine call with host
(pinned) buffers, H2D and D2H copies inside the timed region.  `--impl reference` times the CPU port
of the reference (oracle/restatement.py, pinned bit-for-bit to the reference's own code) on the
host cores, on a bounded sample of the same workload.
"""

from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time
from pathlib import Path

import numpy as np
import torch

ROOT = Path(__file__).resolve().parent
if str(ROOT) not in sys.path:
    sys.path.insert(0, str(ROOT))

WORKLOADS = {
    # name: (views per GPU, width, height, K, sparse points per view, description)
    "cfg1": (20, 512, 384, 4, 4096, "BASELINE configs[0]: synthetic 20-view COLMAP scene, 512x384, K=4, 1 cm voxels"),
    "cfg2": (185, 1297, 840, 8, 4096, "BASELINE configs[1]: garden-shaped synthetic, 185 views at 1297x840, K=8, 1 cm voxels"),
    "cfg3": (200, 1920, 1080, 8, 4096, "BASELINE configs[2]: 200 views at 1920x1080, K=8, 1 cm voxels"),
    "cfg4": (1000, 1600, 1200, 10, 4096, "BASELINE configs[3]: 1000 views at 1600x1200, K=10, 1 cm voxels (use --scaling strong on 8 GPUs)"),
    "cfg5": (300, 3840, 2160, 8, 4096, "BASELINE configs[4]: 300 views at 3840x2160 (4K), K=8 assumed, 1 cm voxels (--scaling strong, 8 GPUs)"),
    "cfg5_slice": (24, 3840, 2160, 8, 4096, "24 of the 300 4K views of BASELINE configs[4] on one GPU (shape coverage, not a BASELINE config)"),
    "small": (12, 320, 240, 4, 1024, "smoke-sized scene (not a BASELINE config)"),
}
VOXEL = 0.01
METRIC = "depth_pixels_fused_per_s"
UNIT = "pixels/s"


def env_int(name, default):
    return int(os.environ.get(name, default))


# ----------------------------------------------------------------------------------------------
# clocks
# ----------------------------------------------------------------------------------------------
class ClockSampler:
    FIELDS = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
              "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
              "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index: int):
        self.gpu = gpu_index
        self.lines = []
        self.proc = None

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--query-gpu={self.FIELDS}", "--format=csv,noheader,nounits", "-lms", "20", "-i", str(self.gpu)],
                stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except OSError:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append((time.time(), line.strip()))

    def stop(self, t0=None, t1=None):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except subprocess.TimeoutExpired:
            self.proc.kill()
        sm, mx, reasons = [], None, set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ts, line in self.lines:
            if t0 is not None and not (t0 <= ts <= t1 + 0.15):
                continue
            parts = [p.strip() for p in line.split(",")]
            if len(parts) < 8:
                continue
            try:
                sm.append(float(parts[1]))
                mx = float(parts[2])
            except ValueError:
                continue
            for n, val in zip(names, parts[4:8]):
                if val.lower().startswith("active"):
                    reasons.add(n)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": mx, "reasons": sorted(reasons), "samples": len(sm)}


# ----------------------------------------------------------------------------------------------
# reference arm / cpu baseline (the ONLY place bench.py touches oracle/)
# ----------------------------------------------------------------------------------------------
def cpu_port_sample(workload: str, n_sample_views: int = 9, repeats: int = 1):
    """Time the CPU port of the reference (stages 1-4) on `n_sample_views` consecutive views of the
    workload at full resolution, every view testing against all the others (K = n-1)."""
    from depthdensifier_b200.hashperm import hash_perm
    from depthdensifier_b200.neighbours import default_vote_threshold, nearest_views_table
    from depthdensifier_b200.synthetic import SceneConfig, make_scene
    from oracle import restatement as R

    V, W, H, K, C, _ = WORKLOADS[workload]
    n = min(n_sample_views, V)
    k = min(K, n - 1)
    sc = make_scene(SceneConfig(n_views=V, width=W, height=H, n_sparse=C, seed=0), device="cpu", views=range(n))
    poses = sc.cam_from_world.numpy()[:n]
    intr = sc.intrinsics.numpy()[:n]
    nbr = nearest_views_table(poses, k)
    thr = default_vote_threshold(k)
    args = (sc.mono_depth.numpy(), sc.normal.numpy(), sc.mask.numpy(), sc.rgb.numpy(), sc.sparse_xyz.numpy(),
            sc.sparse_offsets.numpy(), poses, intr, nbr, thr)
    times, px = [], 0
    for _ in range(repeats):
        t0 = time.perf_counter()
        out = R.densify(*args, randperm=lambda m: hash_perm(m, 0), voxel=VOXEL)
        times.append(time.perf_counter() - t0)
        px = int(len(out["points"]))
    sample = (f"{n} consecutive views of {workload} at {W}x{H}, K={k} nearest of those views, stages 1-4 "
              f"(align, back-project, consistency vote, voxel fusion), {px} valid pixels")
    return px, times, sample, k


def run_reference_arm(args):
    rank = env_int("RANK", 0)
    if rank != 0:
        return 0
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    V, W, H, K, C, desc = WORKLOADS[args.workload]
    # bounded sample: one step costs ~1.2 s per sampled cfg-2 view on these cores; keep the whole run to minutes
    reps = args.warmup + args.steps
    n_views = args.cpu_views if reps <= 8 else min(args.cpu_views, 6) if reps <= 20 else min(args.cpu_views, 4)
    px, times, sample, k = cpu_port_sample(args.workload, n_views, repeats=reps)
    timed = times[args.warmup:]
    ms = 1e3 * float(np.mean(timed))
    value = px / (ms / 1e3)
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f64", "data": "synthetic",
        "config": {"workload": args.workload, "description": desc, "views": V, "width": W, "height": H, "k_neighbours": K,
                   "voxel": VOXEL, "sparse_per_view": C},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample,
                         "torch_threads": torch.get_num_threads(),
                         "note": "numpy stages 2-4 are effectively single-threaded, torch CPU ops of stage 1 use all threads"},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)
    return 0


# ----------------------------------------------------------------------------------------------
# our arm
# ----------------------------------------------------------------------------------------------
def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="cfg2", choices=sorted(WORKLOADS))
    ap.add_argument("--scaling", default="weak", choices=["weak", "strong"])
    ap.add_argument("--cpu-views", type=int, default=9)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    args = ap.parse_args()
    if args.warmup < 3:
        args.warmup = max(args.warmup, 1) if args.workload == "small" else 3
    if args.impl == "reference":
        return run_reference_arm(args)

    from depthdensifier_b200 import _lib
    from depthdensifier_b200 import build as ddn_build

    rank, world, local = env_int("RANK", 0), env_int("WORLD_SIZE", 1), env_int("LOCAL_RANK", 0)
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: there is no CPU fallback (use --impl reference for the CPU arm)")
    if world == 1 or rank == 0:
        ddn_build.build()
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    numa_cpus = None
    if world > 1:  # keep each rank's pinned host buffers on the socket next to its GPU
        from depthdensifier_b200.hostmem import bind_to_gpu_numa

        numa_cpus = bind_to_gpu_numa(local)
    dist = None
    if world > 1:
        import torch.distributed as dist

        # rank 0 prints ONE JSON line on stdout: keep NCCL's own messages (its version banner goes to stdout
        # at any NCCL_DEBUG level) on stderr
        if os.environ.get("NCCL_DEBUG", "").upper() == "VERSION":
            os.environ.pop("NCCL_DEBUG")
        os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")

        dist.init_process_group("nccl", device_id=dev)
        dist.barrier()
    _lib.load()

    from depthdensifier_b200.distributed import ShardedDensifier
    from depthdensifier_b200.engine import DensifyConfig
    from depthdensifier_b200.neighbours import default_vote_threshold, nearest_views_table
    from depthdensifier_b200.synthetic import SceneConfig, make_scene

    Vper, W, H, K, C, desc = WORKLOADS[args.workload]
    if args.scaling == "weak":
        V_total = Vper * world
        lo, hi = rank * Vper, (rank + 1) * Vper
    else:
        V_total = Vper
        per = (V_total + world - 1) // world
        lo, hi = min(rank * per, V_total), min((rank + 1) * per, V_total)
    scfg = SceneConfig(n_views=V_total, width=W, height=H, n_sparse=C, seed=0)
    sc = make_scene(scfg, device=dev, views=range(lo, hi))
    poses_np = sc.cam_from_world.cpu().numpy()
    nbr_np = nearest_views_table(poses_np, K)
    thr = default_vote_threshold(K)
    torch.cuda.synchronize()

    sharded = ShardedDensifier(DensifyConfig(voxel=VOXEL, vote_threshold=thr), dev, rank, world, V_total, lo, hi,
                               sc.cam_from_world, sc.intrinsics, nbr_np, H, W)
    dev_inputs = (sc.mono_depth, sc.normal, sc.mask, sc.rgb, sc.sparse_xyz, sc.sparse_offsets)

    def barrier():
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- device-resident timing ----
    for _ in range(args.warmup):
        res = sharded.run(*dev_inputs)
    barrier()
    n_valid_local = int((res.votes != 255).sum().item())
    n_kept_local, n_vox_local = [int(x) for x in res.counts.cpu().tolist()]
    sampler = ClockSampler(local)
    sampler.start()
    time.sleep(0.3)
    launches0 = _lib.launch_count()
    stage_ms = {}
    barrier()
    t_wall0 = time.time()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ev0.record()
    stage_events = []
    for _ in range(args.steps):
        res = sharded.run(*dev_inputs, record_events=True)
        stage_events.append(res.events)
    ev1.record()
    barrier()
    t_wall1 = time.time()
    launches = _lib.launch_count() - launches0
    clocks = sampler.stop(t_wall0, t_wall1)
    ms_total = ev0.elapsed_time(ev1)
    ms_step = ms_total / args.steps
    for evs in stage_events:
        for name, (a, b) in evs.items():
            stage_ms.setdefault(name, []).append(a.elapsed_time(b))
    stage_ms = {k: float(np.mean(v)) for k, v in stage_ms.items()}
    # multi-GPU runs launch K4 twice (views that need no halo first, the shard's boundary views after the halo)
    k4_local_ms = stage_ms["backproject_filter"] + stage_ms.get("backproject_filter_boundary", 0.0)

    def allreduce(x, op):
        if dist is None:
            return x
        t = torch.tensor([x], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=op)
        return float(t.item())

    ms_step = allreduce(ms_step, dist.ReduceOp.MAX if dist else None)
    n_valid = int(allreduce(float(n_valid_local), dist.ReduceOp.SUM if dist else None))
    n_kept = int(allreduce(float(n_kept_local), dist.ReduceOp.SUM if dist else None))
    n_vox = int(allreduce(float(n_vox_local), dist.ReduceOp.SUM if dist else None))
    k4_ms = allreduce(k4_local_ms, dist.ReduceOp.MAX if dist else None)
    value = n_valid / (ms_step / 1e3)

    # ---- roofline of the judged kernel (K4 back-project + consistency), algorithmic bytes ----
    peaks_file = ROOT / "MEASURED_PEAKS.json"
    if peaks_file.exists():
        peak = float(json.loads(peaks_file.read_text())["hbm_gbs"])
        peak_src = "measured (MEASURED_PEAKS.json hbm_gbs)"
    else:
        peak, peak_src = 6650.0, "fallback (B200_PROFILING.md)"
    bytes_per_px = 29 + 4 * K
    k4_bytes = bytes_per_px * n_valid_local
    achieved = k4_bytes / (k4_local_ms * 1e-3) / 1e9
    roofline = {"bound": "hbm", "kernel": "backproject_filter_kernel", "achieved": achieved, "peak": peak, "unit": "GB/s",
                "frac": achieved / peak, "traffic": None, "peak_source": peak_src,
                "algorithmic_bytes_per_pixel": bytes_per_px, "pixels_per_launch": n_valid_local,
                "kernel_ms": k4_local_ms}
    traffic_file = ROOT / "profiles" / "k4_traffic.json"
    if traffic_file.exists():
        try:
            roofline["traffic"] = json.loads(traffic_file.read_text()).get(args.workload)
        except (ValueError, OSError):
            pass

    # ---- end to end through the public engine call with host buffers ----
    e2e = None
    if not args.no_e2e:
        del res
        torch.cuda.empty_cache()
        host = sharded.pin_host_inputs(*[t.cpu() for t in dev_inputs])

        def time_e2e(**kw):
            for _ in range(2):
                out_host = sharded.run_host(*host, **kw)
            barrier()
            t0 = time.perf_counter()
            e_steps = max(2, min(args.steps, 5))
            for _ in range(e_steps):
                out_host = sharded.run_host(*host, **kw)
            barrier()
            e_ms = (time.perf_counter() - t0) * 1e3 / e_steps
            e_ms = allreduce(e_ms, dist.ReduceOp.MAX if dist else None)
            return e_ms, int(out_host["h2d_bytes"]), int(out_host["d2h_bytes"])

        e_ms, h2d, d2h = time_e2e()
        f_ms, f_h2d, _ = time_e2e(normals_in_place=False)
        tot = lambda x: int(allreduce(float(x), dist.ReduceOp.SUM if dist else None))
        e2e = {"value": n_valid / (e_ms / 1e3), "unit": UNIT, "ms_per_step": e_ms,
               "h2d_bytes_per_step": tot(h2d), "d2h_bytes_per_step": tot(d2h),
               "api": "ShardedDensifier.run_host (pinned host arrays in, fused cloud out through pinned buffers)",
               "note": "depth, mask, colours and sparse points are copied to the device every step; the normal maps "
                       "stay in pinned host memory and the consistency kernel reads only the normals of its vote "
                       "candidates over PCIe (not counted in h2d_bytes_per_step)",
               "all_inputs_copied": {"value": n_valid / (f_ms / 1e3), "ms_per_step": f_ms, "h2d_bytes_per_step": tot(f_h2d)}}

    # ---- CPU baseline on this box's host cores (rank 0, N=1 only) ----
    cpu_baseline = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        cores = os.cpu_count() or 1
        torch.set_num_threads(cores)
        px, times, sample, _k = cpu_port_sample(args.workload, args.cpu_views, repeats=1)
        cpu_baseline = {"value": px / times[0], "unit": UNIT, "cores": cores, "kind": "port", "sample": sample,
                        "seconds": times[0]}

    if rank == 0:
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms_step, "higher_is_better": True, "scaling": args.scaling, "vs_baseline": None,
            "dtype": "f32", "data": "synthetic",
            "config": {"workload": args.workload, "description": desc, "views_total": V_total, "views_per_gpu": hi - lo,
                       "width": W, "height": H, "k_neighbours": K, "vote_threshold": thr, "voxel": VOXEL,
                       "sparse_per_view": C, "align_mode": "pwl", "sample_mode": "nearest",
                       "l2": "inputs per step (>= 3.9 GB at cfg2) are far larger than the 126 MB L2; no explicit flush",
                       "valid_pixels": n_valid, "kept_points": n_kept, "voxels": n_vox},
            "roofline": roofline, "cpu_baseline": cpu_baseline, "e2e": e2e, "gpu_launches": int(launches),
            "host_cpus_rank0": (f"{numa_cpus[0]}-{numa_cpus[-1]} ({len(numa_cpus)})" if numa_cpus else None),
            "clocks": clocks, "stages_ms": stage_ms,
            "stage_GBps_algorithmic": {
                "align_remap": 9 * (hi - lo) * H * W / (stage_ms.get("align", float("nan")) * 1e-3) / 1e9,
                "backproject_filter": achieved,
                # compulsory bytes of the fusion: 16 B per kept point in, 28 B per voxel out
                "voxel_fuse": (16 * n_kept_local + 28 * n_vox_local) / (stage_ms.get("voxel_fuse", float("nan")) * 1e-3) / 1e9,
            },
        }
        print(json.dumps(line), flush=True)
    if dist is not None:
        dist.barrier()
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
