#!/bin/bash
# 2 GPUs: two-rank hardware test + bench (barrier-free step start)
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_multi.py -m gpu -q > gpurun_out/pytest_r02l.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/pytest_r02l.log
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus 2 --steps 10 --warmup 3 --no-e2e > gpurun_out/bench_r02l_g2.json 2> gpurun_out/bench_r02l_g2.err; echo "bench g2 rc=$?"
python - <<'PY'
import json
try:
    d=json.loads([l for l in open("gpurun_out/bench_r02l_g2.json") if l.startswith("{")][-1])
    print("g2", round(d["ms_per_step"],3), d["path"], "host", d["host_enqueue_ms_per_step"], {k:round(v,3) for k,v in d["stages_ms"].items()})
    print("check", d["multi_gpu_check"] and (d["multi_gpu_check"]["passed"], d["multi_gpu_check"]["path"]))
    s=d["strong"]; print("strong", round(s["ms_per_step"],3), "host", s["host_enqueue_ms_per_step"], {k:round(v,3) for k,v in s["stages_ms"].items()})
    print(s.get("stage4_kernels_ms_per_rank"))
except Exception as e: print("ERR", e)
PY
