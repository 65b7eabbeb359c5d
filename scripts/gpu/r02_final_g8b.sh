#!/bin/bash
mkdir -p gpurun_out
run() { timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29541 bench.py --gpus 8 "$@"; }
run --steps 20 --warmup 5 > gpurun_out/r02_bench_g8.json 2> gpurun_out/r02_bench_g8.err; echo "bench g8 rc=$?"
python - <<'PY'
import json
for f in ("r02_bench_g8", "r02_bench_g8_nooverlap"):
    try:
        d=json.loads([l for l in open(f"gpurun_out/{f}.json") if l.startswith("{")][-1])
        print(f, round(d["ms_per_step"],3), d["path"], d.get("multi_gpu_check") and d["multi_gpu_check"]["passed"], {k:round(v,3) for k,v in d["stages_ms"].items()})
        print("  strong", round(d["strong"]["ms_per_step"],3), {k:round(v,3) for k,v in d["strong"]["stages_ms"].items()})
        if d.get("e2e"): e=d["e2e"]; print("  e2e", round(e["ms_per_step"],2), round(e["ms_per_step_one_call_at_a_time"],2), round(e["ms_per_step_all_copied"],2))
    except Exception as e: print(f, "ERR", e)
PY
