#!/bin/bash
# 2 GPUs: the per-rank footprints of cfg4 / cfg5 on 8 GPUs (de-risks the 8-GPU call), then the default bench
mkdir -p gpurun_out
run() { timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus 2 "$@"; }
run --workload cfg5_quarter --scaling strong --steps 3 --warmup 3 --no-e2e --no-check --dedup-sparse > gpurun_out/r02_cfg5q_g2.json 2> gpurun_out/r02_cfg5q_g2.err; echo "cfg5q rc=$?"; tail -2 gpurun_out/r02_cfg5q_g2.err | cut -c1-300
run --workload cfg4_quarter --scaling strong --steps 3 --warmup 3 --no-e2e --no-check > gpurun_out/r02_cfg4q_g2.json 2> gpurun_out/r02_cfg4q_g2.err; echo "cfg4q rc=$?"; tail -2 gpurun_out/r02_cfg4q_g2.err | cut -c1-300
run --steps 20 --warmup 5 > gpurun_out/r02_bench_g2.json 2> gpurun_out/r02_bench_g2.err; echo "bench g2 rc=$?"
timeout 600 python -m pytest tests/test_gpu_multi.py -m gpu -q > gpurun_out/r02_pytest_multi_g2.log 2>&1; echo "pytest multi rc=$?"
python - <<'PY'
import json
for f in ("r02_cfg5q_g2", "r02_cfg4q_g2", "r02_bench_g2"):
    try:
        d=json.loads([l for l in open(f"gpurun_out/{f}.json") if l.startswith("{")][-1])
        print(f, round(d["ms_per_step"],3), d["path"], d["workload_stats"], {k:round(v,3) for k,v in d["stages_ms"].items()})
        if d.get("strong"): print("  strong", round(d["strong"]["ms_per_step"],3))
        if d.get("e2e"): print("  e2e", round(d["e2e"]["ms_per_step"],2), round(d["e2e"]["ms_per_step_all_copied"],2))
    except Exception as e: print(f, "ERR", e)
PY
