#!/bin/bash
# 8 GPUs: the driver's default command, then BASELINE configs[3] and configs[4] at full size (strong: the views are
# split over the ranks)
mkdir -p gpurun_out
run() { timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29541 bench.py --gpus 8 "$@"; }
run --steps 20 --warmup 5 > gpurun_out/r02_bench_g8.json 2> gpurun_out/r02_bench_g8.err; echo "bench g8 rc=$?"
run --workload cfg4 --scaling strong --steps 5 --warmup 3 --no-e2e --no-check > gpurun_out/r02_cfg4_g8.json 2> gpurun_out/r02_cfg4_g8.err; echo "cfg4 rc=$?"; tail -2 gpurun_out/r02_cfg4_g8.err | cut -c1-300
run --workload cfg5 --scaling strong --steps 5 --warmup 3 --no-e2e --no-check --dedup-sparse > gpurun_out/r02_cfg5_g8.json 2> gpurun_out/r02_cfg5_g8.err; echo "cfg5 rc=$?"; tail -2 gpurun_out/r02_cfg5_g8.err | cut -c1-300
python - <<'PY'
import json
for f in ("r02_bench_g8", "r02_cfg4_g8", "r02_cfg5_g8"):
    try:
        d=json.loads([l for l in open(f"gpurun_out/{f}.json") if l.startswith("{")][-1])
        print(f, round(d["ms_per_step"],3), d["path"], d["workload_stats"], {k:round(v,3) for k,v in d["stages_ms"].items()})
        if d.get("strong"): print("  strong", round(d["strong"]["ms_per_step"],3), d["multi_gpu_check"]["passed"])
        if d.get("e2e"): print("  e2e", round(d["e2e"]["ms_per_step"],2), round(d["e2e"]["ms_per_step_all_copied"],2))
    except Exception as e: print(f, "ERR", e)
PY
