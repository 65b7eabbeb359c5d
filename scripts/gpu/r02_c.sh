#!/bin/bash
# round 2, GPU call C (2 GPUs): failed tests again + two-rank hardware test + 2-GPU bench with the strong block
mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_gpu_session.py tests/test_gpu_parity.py tests/test_gpu_multi.py -m gpu -q > gpurun_out/pytest_r02c.log 2>&1; echo "pytest rc=$?" | tee -a gpurun_out/pytest_r02c.log
tail -8 gpurun_out/pytest_r02c.log
timeout 300 python scripts/gpu/kbench.py cfg2 5 > gpurun_out/kbench_r02c.json 2> gpurun_out/kbench_r02c.err; cat gpurun_out/kbench_r02c.json
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus 2 --steps 10 --warmup 3 > gpurun_out/bench_r02c_g2.json 2> gpurun_out/bench_r02c_g2.err; echo "bench g2 rc=$?"
tail -5 gpurun_out/bench_r02c_g2.err
python - <<'PY'
import json
try:
    d=json.loads([l for l in open("gpurun_out/bench_r02c_g2.json") if l.startswith("{")][-1])
    print("g2", round(d["ms_per_step"],3), d["path"], {k:round(v,3) for k,v in d["stages_ms"].items()})
    print("check", d["multi_gpu_check"] and (d["multi_gpu_check"]["passed"], d["multi_gpu_check"]["path"]))
    s=d["strong"]; print("strong", round(s["ms_per_step"],3), {k:round(v,3) for k,v in s["stages_ms"].items()})
    print("e2e", d["e2e"] and round(d["e2e"]["ms_per_step"],2))
except Exception as e: print("ERR", e)
PY
