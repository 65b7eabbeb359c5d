#!/bin/bash
# 8 GPUs: default bench (weak cfg2 + multi-GPU check + strong cfg3 block with the stage-4 kernel list)
mkdir -p gpurun_out
TAG=${1:-r02g}
timeout 900 python -m pytest tests/test_gpu_session.py -m gpu -q -k merge > gpurun_out/pytest_${TAG}.log 2>&1; echo "pytest rc=$?"; tail -2 gpurun_out/pytest_${TAG}.log
timeout 1200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29541 bench.py --gpus 8 --steps 10 --warmup 3 --no-e2e > gpurun_out/bench_${TAG}_g8.json 2> gpurun_out/bench_${TAG}_g8.err; echo "bench g8 rc=$?"
tail -3 gpurun_out/bench_${TAG}_g8.err | cut -c1-300
python - <<PY
import json
try:
    d=json.loads([l for l in open("gpurun_out/bench_${TAG}_g8.json") if l.startswith("{")][-1])
    print("g8 weak", round(d["ms_per_step"],3), d["path"], {k:round(v,3) for k,v in d["stages_ms"].items()})
    print("check", d["multi_gpu_check"] and (d["multi_gpu_check"]["passed"], d["multi_gpu_check"]["path"]))
    s=d["strong"]; print("strong", round(s["ms_per_step"],3), {k:round(v,3) for k,v in s["stages_ms"].items()})
    print("kernels", s.get("stage4_kernels_ms")); print("per rank", s.get("stage4_kernels_ms_per_rank")); print("host", s.get("host_enqueue_ms_per_step"), d.get("host_enqueue_ms_per_step"))
except Exception as e: print("ERR", e)
PY
